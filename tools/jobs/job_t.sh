# GPU job t: static HRTF panners lowered to the convolver kernels; chain kernels back at the r2_r state
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2t_tests.log 2>&1; tail -12 gpurun_out/r2t_tests.log
echo "== kernel-only C2: $(timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2t_bench.err | tail -1 | cut -c1-200)"
echo "== fft  $(timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
echo "== fir  $(WAE_HRTF_FFT=0 timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C3 1 1 2>&1 | tail -1)"
