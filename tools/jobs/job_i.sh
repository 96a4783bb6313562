# GPU job i: k_chain occupancy experiment (7 CTAs / SM), then the reference's own benchmark scenarios (examples/benchmarks.rs) as batches of 64
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
run() { echo "== $3 TMA=$1 WAVES=$2: $(WAE_CHAIN_TMA=$1 WAE_CHAIN_WAVES=$2 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2i_bench.err | tail -1 | cut -c1-150)"; }
run 0 16 v7_minb6
WAE_NVCC_DEFS="-DWAE_CH_MINB=7 -DWAE_CH_STAGES=3" python __graft_entry__.py --force > /dev/null 2>&1
run 0 16 minb7_st3; run 0 24 minb7_st3
WAE_NVCC_DEFS="-DWAE_CH_MINB=8 -DWAE_CH_STAGES=2" python __graft_entry__.py --force > /dev/null 2>&1
run 0 16 minb8_st2
python __graft_entry__.py --force > /dev/null 2>&1
timeout 1500 python tools/reference_benchmarks.py --seconds 120 --graphs 64 --steps 2 --out gpurun_out/r2_i_reference_benchmarks_64graphs_120s.json 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.rstrip()[:200]); continue
    if 'error' in r: print(r['scenario'], 'ERROR', r['error'][:150])
    else: print('%-55s gpu %9.2f ms  prep %8.1f ms  x_rt %10.0f  cpu1 %8.0f  cpuall %s  diff %.1e' % (r['scenario'][:55], r['gpu_ms_per_batch'], r['prepare_ms'], r['gpu_x_realtime'], r['cpu_1core_x_realtime'], str(round(r.get('cpu_allcores_x_realtime', 0))), r['max_abs_diff']))
"
