# GPU job z: the tree as it is handed over — full suite, smoke, bench line
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2z_tests.log 2>&1; tail -4 gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_z_bench_full.json 2> gpurun_out/r2z_bench_full.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_z_bench_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e ms', d['e2e']['ms_per_step'], 'pinned', d['e2e_pinned_out']['ms_per_step'], 'warm', d['e2e_warm']['ms_per_step'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])
for w in d.get('other_workloads', []): print(w['workload'], round(w['ms_per_step'], 3), w.get('stages_ms_per_step'), 'e2e', round(w.get('e2e_ms_per_step', 0), 1), 'diff', w.get('cpu_port', {}).get('max_abs_diff_vs_gpu'), w.get('cpu_port', {}).get('max_abs_diff_vs_timed_batch'))
PY
