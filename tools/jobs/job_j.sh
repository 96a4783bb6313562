# GPU job j: full suite after the IIR->biquad lowering, a-rate coefficient kernel, stereo mix fast path, host-memory API; scenarios again
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2j_tests.log 2>&1; tail -15 gpurun_out/r2j_tests.log
echo "== kernel-only C2: $(timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2j_bench.err | tail -1 | cut -c1-200)"
timeout 1500 python tools/reference_benchmarks.py --seconds 120 --graphs 64 --steps 2 --out gpurun_out/r2_j_reference_benchmarks_64graphs_120s.json 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.rstrip()[:200]); continue
    if 'error' in r: print(r['scenario'], 'ERROR', r['error'][:150])
    else: print('%-55s gpu %9.2f ms  prep %8.1f ms  x_rt %10.0f  cpu1 %8.0f  cpuall %s  diff %.1e' % (r['scenario'][:55], r['gpu_ms_per_batch'], r['prepare_ms'], r['gpu_x_realtime'], r['cpu_1core_x_realtime'], str(round(r.get('cpu_allcores_x_realtime', 0))), r['max_abs_diff']))
"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_j_bench_full.json 2> gpurun_out/r2j_bench_full.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_j_bench_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'], d['roofline'])
for w in d.get('workloads', []): print(w)
PY
