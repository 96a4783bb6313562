# GPU job p: zero-padded IR spectra (no range tests in k_conv_mac), one modulo per thread for looping sources; suite, C4,
# ncu of the conv kernels, a full bench line
mkdir -p gpurun_out
python __graft_entry__.py > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/r2p_tests.log 2>&1; tail -8 gpurun_out/r2p_tests.log
echo "== kernel-only C2: $(timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>gpurun_out/r2p_bench.err | tail -1 | cut -c1-200)"
echo "== C4 128 x 10 s: $(timeout 300 python tools/profile_workload.py C4 128 10 2>&1 | tail -1)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_conv_(fft_in|mac|ifft)" -s 6 -c 3 -o gpurun_out/r2p_conv_c4 python tools/profile_workload.py C4 64 10 > gpurun_out/r2p_ncu.log 2>&1; tail -2 gpurun_out/r2p_ncu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_chain" -s 2 -c 1 -o gpurun_out/r2p_chain_ns python tools/profile_workload.py north_star 2 2 > gpurun_out/r2p_ncu2.log 2>&1; tail -2 gpurun_out/r2p_ncu2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_p_bench_full.json 2> gpurun_out/r2p_bench_full.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_p_bench_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'])
for w in d.get('other_workloads', []): print(w['workload'], w['ms_per_step'], w.get('stages_ms_per_step'), w.get('e2e_ms_per_step'), [ (k.get('frac'), k.get('ms')) for k in w.get('kernel_rooflines', [])])
PY
