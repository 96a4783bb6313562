# GPU job l: full suite (IIR lowering only without suspend points; convolver MAC walked by diagonals; convolver reads the source buffer in
# place / writes the destination), C4 + north_star timing, ncu of k_conv_mac
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1; tail -15 gpurun_out/r2l_tests.log
echo "== C4 128 x 10 s: $(timeout 300 python tools/profile_workload.py C4 128 10 2>&1 | tail -1)"
echo "== north_star 8 x 10 s: $(timeout 300 python tools/profile_workload.py north_star 8 10 2>&1 | tail -1)"
echo "== C5 256 x 5 s: $(timeout 300 python tools/profile_workload.py C5 256 5 2>&1 | tail -1)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_conv_(fft_in|mac|ifft)" -s 6 -c 3 -o gpurun_out/r2l_conv_c4 python tools/profile_workload.py C4 64 10 > gpurun_out/r2l_ncu.log 2>&1; tail -2 gpurun_out/r2l_ncu.log
