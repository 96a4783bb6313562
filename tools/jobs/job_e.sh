# GPU job e: dynamic-layout tests + fuzz, then k_chain thread-count experiments
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_fuzz.py --deselect tests/test_gpu_dynamic_layout.py > gpurun_out/r2e_tests_base.log 2>&1; tail -6 gpurun_out/r2e_tests_base.log
timeout 900 python -m pytest tests/test_gpu_dynamic_layout.py -m gpu -q -p no:cacheprovider > gpurun_out/r2e_tests_dyn.log 2>&1; tail -40 gpurun_out/r2e_tests_dyn.log | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -p no:cacheprovider > gpurun_out/r2e_tests_fuzz.log 2>&1; tail -30 gpurun_out/r2e_tests_fuzz.log | cut -c1-220
run() { echo "== $3 TMA=$1 WAVES=$2"; WAE_CHAIN_TMA=$1 WAE_CHAIN_WAVES=$2 timeout 300 python bench.py --extra 0 --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2e_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], 'e2e', d['e2e']['ms_per_step'], 'pinned', d['e2e_pinned_out']['ms_per_step'], 'warm', d['e2e_warm']['ms_per_step'])"; tail -2 gpurun_out/r2e_bench.err; }
run 0 20 t128_st4; run 0 0 t128_st4
WAE_NVCC_DEFS="-DWAE_CH_THREADS=64" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 t64_st4; run 0 40 t64_st4
WAE_NVCC_DEFS="-DWAE_CH_THREADS=32" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 t32_st4; run 0 40 t32_st4; run 0 0 t32_st4
WAE_NVCC_DEFS="-DWAE_CH_THREADS=32 -DWAE_CH_STAGES=3" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 t32_st3; run 1 20 t32_st3
WAE_NVCC_DEFS="-DWAE_CH_THREADS=32 -DWAE_CH_STAGES=2" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 t32_st2
WAE_NVCC_DEFS="-DWAE_CH_THREADS=64 -DWAE_CH_STAGES=2" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 t64_st2
