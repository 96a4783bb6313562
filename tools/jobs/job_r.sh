# GPU job r: k_voice_sum with the hand-off off the barrier path (relaxed polls, early state loads, publisher in warp 1), k_mix small-port path
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 300 python -m pytest tests/test_gpu_voice_sum.py -q -p no:cacheprovider -x > gpurun_out/r2r_vsum_tests.log 2>&1; tail -5 gpurun_out/r2r_vsum_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_dynamic_layout.py -q -p no:cacheprovider -x > gpurun_out/r2r_tests.log 2>&1; tail -3 gpurun_out/r2r_tests.log
echo "== fused $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== unfused $(WAE_VOICE_SUM=0 timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C3 1 1 2>&1 | tail -1)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_voice_sum" -s 1 -c 1 -o gpurun_out/r2r_vsum python tools/profile_workload.py north_star 8 10 > gpurun_out/r2r_ncu.log 2>&1; tail -1 gpurun_out/r2r_ncu.log
WAE_NVCC_DEFS="-DWAE_VS_MINB=4" python __graft_entry__.py --force > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
echo "== VS_MINB=4 fused $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
