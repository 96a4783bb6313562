# GPU job ac: the suite on the tree with shared AudioBuffer assets and the split planner, then the one-shot call on north_star / Granular synthesis
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 300 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2ac_tests.log 2>&1; tail -5 gpurun_out/r2ac_tests.log
timeout 60 python tools/oneshot_workloads.py 3 > gpurun_out/r2ac_oneshot.log 2>&1; cat gpurun_out/r2ac_oneshot.log | tail -8
WAE_PLAN_SPLIT=0 timeout 40 python tools/oneshot_workloads.py 2 > gpurun_out/r2ac_oneshot_nosplit.log 2>&1; grep north gpurun_out/r2ac_oneshot_nosplit.log | tail -2
