# GPU job g: k_chain tuning (kernel-only runs), then the C3 full-size test
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
run() { echo "== $3 TMA=$1 WAVES=$2: $(WAE_CHAIN_TMA=$1 WAE_CHAIN_WAVES=$2 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2g_bench.err | tail -1 | cut -c1-160)"; }
run 0 20 v7; run 0 20 v7; run 0 0 v7; run 0 12 v7; run 0 30 v7
WAE_NVCC_DEFS="-DWAE_CH_STAGES=2" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 v7_st2; run 0 12 v7_st2
WAE_NVCC_DEFS="-DWAE_CH_STAGES=3" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 v7_st3
WAE_NVCC_DEFS="-DWAE_CHAIN_NOCHECK" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 v7_nocheck; run 0 0 v7_nocheck
python __graft_entry__.py --force > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -k c3 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dynamic_layout.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
