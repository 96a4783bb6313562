# GPU job f: all GPU tests (incl. dynamic layouts, fuzz, full-size), C2 kernel timing + ncu, complete bench line
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_fullsize.py > gpurun_out/r2f_tests.log 2>&1; tail -15 gpurun_out/r2f_tests.log | cut -c1-220
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider > gpurun_out/r2f_tests_full.log 2>&1; tail -15 gpurun_out/r2f_tests_full.log | cut -c1-220
run() { echo "== $3 TMA=$1 WAVES=$2"; WAE_CHAIN_TMA=$1 WAE_CHAIN_WAVES=$2 timeout 300 python bench.py --extra 0 --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2f_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], 'e2e', d['e2e']['ms_per_step'], 'pinned', d['e2e_pinned_out']['ms_per_step'], 'warm', d['e2e_warm']['ms_per_step'])"; tail -2 gpurun_out/r2f_bench.err; }
run 0 20 v6; run 0 0 v6; run 0 40 v6
WAE_CHAIN_TMA=0 WAE_CHAIN_WAVES=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 2 -c 1 -o gpurun_out/r2f_chain_v6 python bench.py --extra 0 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2f_ncu.log 2>&1; tail -2 gpurun_out/r2f_ncu.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_full.json 2> gpurun_out/r2f_bench_full.err; tail -c 1500 gpurun_out/r2f_bench_full.err; python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench_full.json').read().strip().splitlines()[-1])
print({k:(d[k] if k not in ('other_workloads','roofline','config','cpu_baseline','clocks') else '...') for k in d})
for w in d.get('other_workloads',[]): print(w['workload'], w['ms_per_step'], w['value'], w.get('e2e_ms_per_step'), w['stages_ms_per_step'], w.get('kernel_rooflines'), w.get('cpu_port'))
print(d['cpu_baseline'])
"
