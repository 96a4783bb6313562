# GPU job x (2 GPUs): the copy-out of the one-shot call through a small ring of page-locked pieces (non-temporal stores) against the
# whole-group staging slots, two ranks on one socket
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -p no:cacheprovider -k "oneshot or c2_thousand or one_shot" > gpurun_out/r2x_tests.log 2>&1; tail -4 gpurun_out/r2x_tests.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 5 --warmup 3 --extra 0 --no-cpu-baseline 2>gpurun_out/r2x_$2.err | tail -1 > gpurun_out/r2x_$2.json; python - <<PY
import json
d = json.loads(open('gpurun_out/r2x_$2.json').read())
print('$2', 'N=2 kernel-only ms', round(d['ms_per_step'], 4), 'e2e pageable ms', round(d['e2e']['ms_per_step'], 1), d['e2e'].get('ms_each'), 'pinned out', round(d['e2e_pinned_out']['ms_per_step'], 1), 'warm', round(d['e2e_warm']['ms_per_step'], 1))
PY
}
run 29521 ring_2mb_x8
WAE_STAGE_PIECE_KB=1024 run 29522 ring_1mb_x8
WAE_STAGE_PIECE_KB=4096 WAE_STAGE_SLOTS=6 run 29523 ring_4mb_x6
WAE_STAGE_RING=0 run 29524 group_slots
timeout 300 python bench.py --steps 3 --warmup 2 --extra 0 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1 ring: e2e pageable ms', round(d['e2e']['ms_per_step'],1), 'pinned out', round(d['e2e_pinned_out']['ms_per_step'],1))"
