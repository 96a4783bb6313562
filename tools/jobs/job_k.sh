# GPU job k: position-order convolver transforms (DIF forward / DIT inverse) — parity, timing of two unroll variants, ncu of the conv kernels;
# localise the two fuzz graphs that regressed in job j; per-stage times of the slow benchmark scenarios
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "conv or Conv or c4 or C4 or c5 or C5 or north or fullsize or scenario" > gpurun_out/r2k_tests.log 2>&1; tail -6 gpurun_out/r2k_tests.log
for t in "" WAE_DEBUG_NO_STEREO4=1 WAE_DEBUG_NO_IIR_CHAIN=1; do echo "== fuzz taps [$t]"; env $t FUZZ_TAPS=1 timeout 300 python tools/debug_fuzz.py 31002 33004 2>&1 | tail -45; done
for v in cvu2 cvu4; do cp build_variants/libwae_$v.so web-audio-api-rs_b200/libwae_b200.so; echo "== C4 128 graphs x 10 s [$v]"; timeout 300 python tools/profile_workload.py C4 128 10 2>&1 | tail -1; done
cp build_variants/libwae_cvu2.so web-audio-api-rs_b200/libwae_b200.so
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_conv_(fft_in|mac|ifft)" -s 6 -c 3 -o gpurun_out/r2k_conv_c4 python tools/profile_workload.py C4 64 10 > gpurun_out/r2k_ncu.log 2>&1; tail -3 gpurun_out/r2k_ncu.log
for sc in "Substractive" "Envelope" "Sawtooth with automation" "Granular" "Biquad filter" "Stereo panning with automation"; do timeout 300 python tools/stage_times.py --scenario "$sc" --graphs 64 --seconds 120 2>&1 | tail -12; done
