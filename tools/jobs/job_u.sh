# GPU job u: static HRTF panners on the convolver kernels with the mono down-mix in the forward transform and the one-partition product
# inside the inverse transform (also taken by every ConvolverNode whose response fits one 8192-frame partition)
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2u_tests.log 2>&1; tail -12 gpurun_out/r2u_tests.log
echo "== fft  $(timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
echo "== fir  $(WAE_HRTF_FFT=0 timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
echo "== $(timeout 300 python tools/profile_workload.py C4 128 10 2>&1 | tail -1)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_conv_ifft" -s 3 -c 1 -o gpurun_out/r2u_ifft python tools/profile_workload.py C5 64 5 > gpurun_out/r2u_ncu.log 2>&1; tail -1 gpurun_out/r2u_ncu.log
