# GPU job v: state of the tree at the end of round 2 — full suite, bench line (C2 + C3 / C4 / north_star / C5), the reference arm, the ncu
# launch list of the bench command, one --set full capture of the C2 kernel (DRAM traffic), the reference's benchmark scenarios
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2v_tests.log 2>&1; tail -6 gpurun_out/r2v_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_v_bench_full.json 2> gpurun_out/r2v_bench_full.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_v_bench_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline'].get('traffic'))
for w in d.get('other_workloads', []): print(w['workload'], round(w['ms_per_step'], 3), w.get('stages_ms_per_step'), 'e2e', round(w.get('e2e_ms_per_step', 0), 1), 'diff', w.get('cpu_port', {}).get('max_abs_diff_vs_gpu'), w.get('cpu_port', {}).get('max_abs_diff_vs_timed_batch'), [(k.get('kernels'), round(k.get('frac', 0), 3)) for k in w.get('kernel_rooflines', [])])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_v_bench_reference_arm.json 2> gpurun_out/r2v_ref.err; tail -c 600 gpurun_out/r2_v_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_v_launches_bench.csv python bench.py --steps 2 --warmup 1 --extra 0 > gpurun_out/r2v_ncu_bench.log 2>&1; grep -c k_chain gpurun_out/r2_v_launches_bench.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_chain" -s 3 -c 1 -o gpurun_out/r2v_chain_c2 python bench.py --kernel-only --steps 2 --warmup 1 > gpurun_out/r2v_ncu_c2.log 2>&1; tail -1 gpurun_out/r2v_ncu_c2.log
timeout 1500 python tools/reference_benchmarks.py --seconds 120 --graphs 64 --steps 2 --out gpurun_out/r2_v_reference_benchmarks_64graphs_120s.json 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.rstrip()[:200]); continue
    if 'error' in r: print(r['scenario'], 'ERROR', r['error'][:150])
    else: print('%-55s gpu %9.2f ms  prep %8.1f ms  x_rt %10.0f  cpu1 %8.0f  cpuall %s  diff %.1e' % (r['scenario'][:55], r['gpu_ms_per_batch'], r['prepare_ms'], r['gpu_x_realtime'], r['cpu_1core_x_realtime'], str(round(r.get('cpu_allcores_x_realtime', 0))), r['max_abs_diff']))
"
