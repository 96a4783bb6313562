# GPU job y (2 GPUs): copy-out of the whole-group staging slots with non-temporal stores against memcpy, one and two ranks
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -p no:cacheprovider -k "oneshot or c2_thousand or one_shot" > gpurun_out/r2y_tests.log 2>&1; tail -2 gpurun_out/r2y_tests.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 5 --warmup 3 --extra 0 --no-cpu-baseline 2>gpurun_out/r2y_$2.err | tail -1 > gpurun_out/r2y_$2.json; python - <<PY
import json
d = json.loads(open('gpurun_out/r2y_$2.json').read())
print('$2', 'N=2 e2e pageable ms', round(d['e2e']['ms_per_step'], 1), d['e2e'].get('ms_each'), 'pinned out', round(d['e2e_pinned_out']['ms_per_step'], 1))
PY
}
run1() { timeout 300 python bench.py --steps 4 --warmup 2 --extra 0 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 N=1: e2e pageable ms', round(d['e2e']['ms_per_step'],1), d['e2e'].get('ms_each'), 'pinned out', round(d['e2e_pinned_out']['ms_per_step'],1))"; }
run 29531 nt_stores
WAE_STAGE_NT=0 run 29532 memcpy
run1 nt_stores
WAE_STAGE_NT=0 run1 memcpy
