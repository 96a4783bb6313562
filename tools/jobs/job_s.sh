# GPU job s: oscillator chains trimmed (host flag, hoisted predicates, direct 16-byte stores, high-word window flags), k_mix direct path for
# small ports, head prefetch of k_voice_sum (opt-in); full suite
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2s_tests.log 2>&1; tail -8 gpurun_out/r2s_tests.log
echo "== kernel-only C2: $(timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2s_bench.err | tail -1 | cut -c1-200)"
echo "== $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== fused(opt-in) $(WAE_VOICE_SUM=1 timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C3 1 1 2>&1 | tail -1)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_chain" -s 2 -c 1 -o gpurun_out/r2s_chain_ns python tools/profile_workload.py north_star 2 2 > gpurun_out/r2s_ncu.log 2>&1; tail -1 gpurun_out/r2s_ncu.log
