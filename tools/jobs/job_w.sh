# GPU job w (2 GPUs, end of round 2): the multi-rank bench legs — C2 sharded, C4 "512 graphs over N GPUs" with the NCCL gather inside the step, and the
# N = 8 leg (C5 + gather) at a reduced size for a path check
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err; tail -c 1200 gpurun_out/r2w_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench_n2.json').read().strip().splitlines()[-1])
print({k:(d[k] if k not in ('other_workloads','roofline','config','cpu_baseline','clocks') else '...') for k in d})
for w in d.get('other_workloads',[]): print(json.dumps({k:w[k] for k in w if k not in ('note','time_batched_model')}))
PY
WAE_BENCH_EXTRA=c5_small timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --graphs 200 --seconds 2 > gpurun_out/r2w_bench_n2_c5.json 2> gpurun_out/r2w_bench_n2_c5.err; tail -c 800 gpurun_out/r2w_bench_n2_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench_n2_c5.json').read().strip().splitlines()[-1])
for w in d.get('other_workloads',[]): print(json.dumps({k:w[k] for k in w if k not in ('note','time_batched_model')}))
print('e2e', d['e2e'], d['e2e_pinned_out'], d['e2e_warm'])
PY
