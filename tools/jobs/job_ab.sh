# GPU job ab: HEAD on a fresh box — the suite and smoke(), as the driver runs them
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 600 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2ab_tests.log 2>&1; tail -3 gpurun_out/r2ab_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
