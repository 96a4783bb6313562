# GPU job m: speculative AudioParam kernel (default) vs the two serial-walk kernels bit for bit; warp-staged a-rate biquad; MAC group modes
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1; tail -15 gpurun_out/r2m_tests.log
for v in mac3 mac2; do cp build_variants/libwae_$v.so web-audio-api-rs_b200/libwae_b200.so; echo "== C4 128 x 10 s [$v]: $(timeout 300 python tools/profile_workload.py C4 128 10 2>&1 | tail -1)"; done
cp build_variants/libwae_mac3.so web-audio-api-rs_b200/libwae_b200.so
for sc in "Substractive" "Envelope" "Sawtooth with automation" "Stereo panning with automation" "Granular"; do timeout 300 python tools/stage_times.py --scenario "$sc" --graphs 64 --seconds 120 2>&1 | tail -7; done
