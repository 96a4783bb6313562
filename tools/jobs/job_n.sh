# GPU job n: a-rate biquad recurrence in groups of four; which of k_biquad_coefs / k_biquad_arate takes the time; k_conv_mac at 2 CTAs / SM;
# the reference's benchmark scenarios again; a full bench line
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2n_tests.log 2>&1; tail -8 gpurun_out/r2n_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_biquad|k_param" --csv --log-file gpurun_out/r2n_substractive_launches.csv python tools/stage_times.py --scenario Substractive --graphs 64 --seconds 20 > gpurun_out/r2n_sub.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2n_substractive_launches.csv')) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split('(')[0]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(',', ''))
for k, (n, t) in agg.items(): print('ncu launch list (64 graphs x 20 s, all runs):', k, n, 'launches', round(t / 1e6, 2), 'ms' )
PY
for v in default mac2; do cp build_variants/libwae_$v.so web-audio-api-rs_b200/libwae_b200.so; echo "== C4 128 x 10 s [$v]: $(timeout 300 python tools/profile_workload.py C4 128 10 2>&1 | tail -1)"; done
cp build_variants/libwae_default.so web-audio-api-rs_b200/libwae_b200.so
timeout 1500 python tools/reference_benchmarks.py --seconds 120 --graphs 64 --steps 2 --out gpurun_out/r2_n_reference_benchmarks_64graphs_120s.json 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.rstrip()[:200]); continue
    if 'error' in r: print(r['scenario'], 'ERROR', r['error'][:150])
    else: print('%-55s gpu %9.2f ms  prep %8.1f ms  x_rt %10.0f  cpu1 %8.0f  cpuall %s  diff %.1e' % (r['scenario'][:55], r['gpu_ms_per_batch'], r['prepare_ms'], r['gpu_x_realtime'], r['cpu_1core_x_realtime'], str(round(r.get('cpu_allcores_x_realtime', 0))), r['max_abs_diff']))
"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_n_bench_full.json 2> gpurun_out/r2n_bench_full.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_n_bench_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'])
for w in d.get('other_workloads', []): print(w['workload'], w['ms_per_step'], w.get('stages_ms_per_step'), w.get('e2e_ms_per_step'), w.get('kernel_rooflines'))
PY
