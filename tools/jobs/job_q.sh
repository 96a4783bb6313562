# GPU job q: k_voice_sum (voices + ordered sum in one kernel), k_mix with staged pointers + alternating batches, oscillator chains
# (pair table, compacted polyBLEP, 96 registers), lighter looping gather (the C2 kernel must be back at 1.34 ms)
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 300 python -m pytest tests/test_gpu_voice_sum.py -q -p no:cacheprovider -x > gpurun_out/r2q_vsum_tests.log 2>&1; tail -15 gpurun_out/r2q_vsum_tests.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider --deselect tests/test_gpu_voice_sum.py > gpurun_out/r2q_tests.log 2>&1; tail -8 gpurun_out/r2q_tests.log
echo "== kernel-only C2: $(timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2q_bench.err | tail -1 | cut -c1-200)"
echo "== fused $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== unfused $(WAE_VOICE_SUM=0 timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C3 1 1 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_voice_sum" -s 1 -c 1 -o gpurun_out/r2q_vsum python tools/profile_workload.py north_star 8 10 > gpurun_out/r2q_ncu.log 2>&1; tail -1 gpurun_out/r2q_ncu.log
WAE_VOICE_SUM=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_chain" -s 2 -c 1 -o gpurun_out/r2q_chain_ns python tools/profile_workload.py north_star 2 2 > gpurun_out/r2q_ncu2.log 2>&1; tail -1 gpurun_out/r2q_ncu2.log
WAE_NVCC_DEFS="-DWAE_CH_MINB_OSC=6 -DWAE_VS_MINB=4" python __graft_entry__.py --force > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
echo "== MINB_OSC=6 VS_MINB=4 fused $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== MINB_OSC=6 unfused $(WAE_VOICE_SUM=0 timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
WAE_NVCC_DEFS="-DWAE_CH_MINB_OSC=4 -DWAE_VS_MINB=6" python __graft_entry__.py --force > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
echo "== MINB_OSC=4 VS_MINB=6 fused $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== MINB_OSC=4 unfused $(WAE_VOICE_SUM=0 timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
