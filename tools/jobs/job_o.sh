# GPU job o: k_chain<PRE> (few long renders), a-rate biquad as a scan over affine maps; the C2 kernel must not have moved
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2o_tests.log 2>&1; tail -12 gpurun_out/r2o_tests.log
echo "== kernel-only C2: $(timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2o_bench.err | tail -1 | cut -c1-200)"
for sc in "Biquad filter" "IIR filter" "Substractive" "Simple source test without resampling (Stereo)"; do timeout 300 python tools/stage_times.py --scenario "$sc" --graphs 64 --seconds 120 2>&1 | tail -6; done
echo "== prepass off"; for sc in "Biquad filter"; do WAE_CHAIN_PREPASS=0 timeout 300 python tools/stage_times.py --scenario "$sc" --graphs 64 --seconds 120 2>&1 | tail -3; done
echo "== 8 graphs"; timeout 300 python tools/stage_times.py --scenario "Biquad filter" --graphs 8 --seconds 120 2>&1 | tail -3
