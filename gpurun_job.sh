mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1; tail -15 gpurun_out/r2c_tests.log
run() { echo "== $3 TMA=$1 WAVES=$2"; WAE_CHAIN_TMA=$1 WAE_CHAIN_WAVES=$2 timeout 300 python bench.py --extra 0 --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2c_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['e2e']['ms_per_step'], d['prepare_ms_once'])"; tail -2 gpurun_out/r2c_bench.err; }
run 1 20 st4; run 0 20 st4; run 0 0 st4; run 1 0 st4
timeout 300 python tools/oneshot_time.py 1000 10 4 2>&1 | tail -12
WAE_NUMA=1 timeout 300 python tools/oneshot_time.py 1000 10 3 2>&1 | tail -8
WAE_NVCC_DEFS="-DWAE_CH_STAGES=3" python __graft_entry__.py --force > /dev/null 2>&1
run 1 20 st3; run 0 20 st3; run 0 40 st3
WAE_NVCC_DEFS="-DWAE_CH_STAGES=2" python __graft_entry__.py --force > /dev/null 2>&1
run 1 20 st2; run 0 20 st2; run 0 0 st2; run 0 40 st2
