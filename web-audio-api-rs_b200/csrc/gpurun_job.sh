mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; tail -12 gpurun_out/r2d_tests.log
run() { echo "== $3 TMA=$1 WAVES=$2"; WAE_CHAIN_TMA=$1 WAE_CHAIN_WAVES=$2 timeout 300 python bench.py --extra 0 --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2d_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], 'e2e', d['e2e']['ms_per_step'], 'pinned', d['e2e_pinned_out']['ms_per_step'], 'warm', d['e2e_warm']['ms_per_step'])"; tail -2 gpurun_out/r2d_bench.err; }
run 0 20 default
WAE_CHAIN_NO_CARVEOUT=1 run 0 20 nocarve
for t in 0 1; do
WAE_CHAIN_TMA=$t WAE_CHAIN_WAVES=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 2 -c 1 -o gpurun_out/r2d_chain_tma$t python bench.py --extra 0 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r2d_ncu$t.log 2>&1; tail -2 gpurun_out/r2d_ncu$t.log
done
WAE_NVCC_DEFS="-DWAE_CHAIN_NOFLUSH" python __graft_entry__.py --force > /dev/null 2>&1
run 0 20 noflush; run 0 0 noflush; run 1 20 noflush
WAE_NVCC_DEFS="-DWAE_CHAIN_NOFLUSH -DWAE_CH_STAGES=2" python __graft_entry__.py --force > /dev/null 2>&1
run 0 0 noflush_st2; run 0 20 noflush_st2
