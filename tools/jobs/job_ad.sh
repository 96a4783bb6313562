# GPU job ad (the last seconds of the round's budget): the suite on the tree with deferred table uploads, most relevant files first
mkdir -p gpurun_out
timeout 30 python -u -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmark_scenarios.py tests/test_gpu_dynamic_layout.py tests/test_gpu_reference_cases.py tests/test_gpu_criterion_and_setters.py tests/test_gpu_voice_sum.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -x -v -m gpu -p no:cacheprovider > gpurun_out/r2ad_tests.log 2>&1; tail -4 gpurun_out/r2ad_tests.log | cut -c1-200
timeout 12 python tools/oneshot_workloads.py 2 > gpurun_out/r2ad_oneshot.log 2>&1; tail -4 gpurun_out/r2ad_oneshot.log
