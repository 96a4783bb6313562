#!/usr/bin/env python
"""bench.py — render-quanta/sec of the OfflineAudioContext hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1], "C2"): 1000 independent OfflineAudioContexts per GPU, each
AudioBufferSource(stereo, seeded uniform noise) -> BiquadFilter(lowpass, seeded f0/Q) -> Gain -> destination,
48 kHz stereo, 10 s (3750 render quanta of 128 frames).  A "step" = one render of the whole batch.
  value : graph-quanta/s, kernel-only (source PCM resident in HBM), CUDA events on the engine's stream
  e2e   : the same through the host API with HOST buffers: H2D of the source PCM (pinned) + render + D2H of the
          rendered PCM, every step
  roofline / cpu_baseline : see DESIGN.md "Measurement"
--impl reference times the reference's CPU algorithm (the oracle port, all host threads) on the same config.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR = 48000.0
METRIC = "offline render-quanta/sec (48kHz stereo, 128-frame)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm_sorted = sorted(sm)
        # median of the samples under load (upper half: idle samples before/after the region pull it down)
        med = sm_sorted[len(sm_sorted) * 3 // 4] if sm_sorted else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_c2_batch(pkg, backend, n_graphs, length, seed_base=0):
    import graphs as G
    ctxs = []
    for g in range(n_graphs):
        ctxs.append(G.c2_buffer_biquad_gain(pkg, backend, seed_base + g, length))
    return ctxs


def measure_workload(pkg, eng, oracle, name, build, n_gpu, n_cpu, length, steps, cores, note):
    """One additional workload of BASELINE.json: kernel-only, e2e (pipelined host->host) and the CPU port, same graphs."""
    import torch
    ctxs = [build(eng.backend, g) for g in range(n_gpu)]
    eng.set_option(pkg.OPT_PIPELINE_GROUPS, 1)
    batch = pkg.Batch(ctxs)
    st = batch.stats()
    batch.set_timing(True)
    for _ in range(2):
        batch.run()
    batch.sync()
    ms = []
    for _ in range(steps):
        batch.run()
        batch.sync()
        ms.append(batch.stats().last_run_ms)
    stages = {}
    for n, t, _k in batch.stage_times():
        stages[n] = stages.get(n, 0.0) + t
    quanta = n_gpu * ((length + 127) // 128)
    host = torch.empty(n_gpu * 2 * length, dtype=torch.float32, pin_memory=True)
    hp = ctypes.c_void_p(host.data_ptr())
    eng.set_option(pkg.OPT_PIPELINE_GROUPS, 0)
    be2e = pkg.Batch(ctxs)
    be2e.run_pipelined(hp)
    t0 = time.perf_counter()
    for _ in range(steps):
        be2e.run_pipelined(hp)
    e2e_s = (time.perf_counter() - t0) / steps
    out = {"workload": name, "note": note, "graphs": n_gpu, "frames_per_graph": length, "ms_per_step": float(np.median(ms)),
           "value": quanta / (float(np.median(ms)) * 1e-3), "e2e_value": quanta / e2e_s, "e2e_ms_per_step": e2e_s * 1e3,
           "unit": "graph-quanta/s", "kernel_launches_per_step": int(st.kernel_launches_per_run), "chunks": int(st.chunks),
           "algorithmic_bytes_per_step": int(st.algorithmic_bytes), "stages_ms_per_step": {k: round(v, 4) for k, v in stages.items()}}
    if oracle is not None and n_cpu > 0:
        octx = [build(oracle, g) for g in range(n_cpu)]
        arr = (ctypes.c_void_p * n_cpu)(*[c._g for c in octx])
        ref = np.empty((n_cpu, 2, length), np.float32)
        secs = ctypes.c_double()
        oracle.api.check(oracle.api.render_many(arr, n_cpu, ref.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), min(cores, n_cpu), ctypes.byref(secs)))
        got = host.numpy().reshape(n_gpu, 2, length)[:n_cpu]
        out["cpu_port"] = {"value": n_cpu * ((length + 127) // 128) / secs.value, "cores_used": min(cores, n_cpu),
                           "sample": f"{n_cpu} graphs, {secs.value:.2f} s wall", "max_abs_diff_vs_gpu": float(np.abs(got - ref).max()),
                           "ref_abs_max": float(np.abs(ref).max())}
    batch.destroy()
    be2e.destroy()
    return out


def run_extra_workloads(pkg, eng, oracle, cores, steps=2):
    import graphs as G
    res = []
    ir3 = G.synthetic_ir(144000, 2, decay=0.6)  # 3 s stereo IR at 48 kHz: 141 partitions of 1024 (C4's parking-garage IR is 175)
    res.append(measure_workload(pkg, eng, oracle, "C3", lambda be, g: G.c3_many_voices(pkg, be, 4096, 48000), 1, 1, 48000, steps, cores,
                                "configs[2]: ONE graph, 4096 x (Oscillator -> Biquad) summed in reference order at the destination, 1 s"))
    res.append(measure_workload(pkg, eng, oracle, "C4", lambda be, g: G.c4_convolver(pkg, be, g, 480000, ir3), 128, min(cores, 128), 480000,
                                steps, cores, "configs[3] scaled to 128 graphs/GPU: stereo source -> Convolver(3 s stereo IR, normalize) -> destination, 10 s"))
    res.append(measure_workload(pkg, eng, oracle, "north_star", lambda be, g: G.north_star_voices_convolver(pkg, be, 1000, 480000, ir3, seed=g),
                                8, 8, 480000, steps, cores,
                                "north_star: 8 graphs/GPU, each 1000 voices (Oscillator -> Biquad -> Gain) summed into one Convolver(3 s IR) -> destination, 10 s"))
    # the reference's IRC_1003_C sphere (44.1 kHz, 512 taps, 187 vertices) cannot travel: synthetic data of the same rate and size,
    # resampled to the 48 kHz context rate by the library exactly as the embedded one would be (~417 taps)
    sphere = G.synthetic_hrir_sphere(44100, 512)
    eng.backend.set_hrir_sphere(sphere)
    if oracle is not None:
        oracle.set_hrir_sphere(sphere)
    res.append(measure_workload(pkg, eng, oracle, "C5", lambda be, g: G.c5_full_chain(pkg, be, g, 192000, ir3), 256, min(cores, 64), 192000,
                                steps, cores,
                                "configs[4] per-GPU share (2048 graphs / 8 GPUs): Oscillator -> WaveShaper -> Biquad -> Convolver(3 s IR) -> "
                                "Panner(HRTF, 44.1 kHz / 512-tap sphere resampled to 48 kHz) -> Analyser -> destination, 4 s"))
    return res


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port: the Rust crate cannot be built here, no cargo)."""
    if rank != 0:
        return
    import __graft_entry__ as ge
    pkg = ge.build()
    oracle = pkg.context.Backend(pkg.Api(ctypes.CDLL(ge.ORACLE_SO), "wao_"))
    cores = os.cpu_count() or 1
    length = int(args.seconds * SR)
    n_sample = args.ref_graphs
    quanta_per_graph = (length + 127) // 128
    times = []
    for step in range(args.warmup + args.steps):
        ctxs = build_c2_batch(pkg, oracle, n_sample, length)
        arr = (ctypes.c_void_p * n_sample)(*[c._g for c in ctxs])
        out = np.empty((n_sample, 2, length), np.float32)
        secs = ctypes.c_double()
        oracle.api.check(oracle.api.render_many(arr, n_sample, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), cores,
                                                ctypes.byref(secs)))
        if step >= args.warmup:
            times.append(secs.value)
        del ctxs, out
    t = float(np.mean(times))
    value = n_sample * quanta_per_graph / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "graph-quanta/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 filter state / f32 PCM", "data": "synthetic",
        "config": {"workload": "C2: AudioBufferSource->Biquad->Gain->destination, 48 kHz stereo, %.0f s per graph" % args.seconds,
                   "graphs_per_step": n_sample, "frames_per_graph": length},
        "cpu_baseline": {"value": value, "unit": "graph-quanta/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} graphs x {args.seconds:.0f} s per step, one context per worker thread"},
        "e2e": {"value": value, "unit": "graph-quanta/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--graphs", type=int, default=1000, help="graphs per GPU (C2: 1000)")
    ap.add_argument("--seconds", type=float, default=10.0, help="rendered seconds per graph (C2: 10)")
    ap.add_argument("--ref-graphs", type=int, default=512, help="graphs per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--cpu-sample-graphs", type=int, default=512)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--groups", type=int, default=32, help="graph groups of the e2e pipeline (H2D | render | D2H overlap)")
    ap.add_argument("--serial-filters", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", type=int, default=1, help="also measure C3 / C4 / north_star (rank 0, N=1 only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    pkg = ge.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = pkg.Engine(local_rank)
    eng.set_option(pkg.OPT_CHUNK_FRAMES, args.chunk)
    eng.set_option(pkg.OPT_SERIAL_FILTERS, args.serial_filters)
    length = int(args.seconds * SR)
    quanta_per_graph = (length + 127) // 128
    n_graphs = args.graphs

    # ---- build the batch: graphs with different seeds per rank (independent shards, weak scaling)
    # Two compiled batches of the same graphs: one un-grouped (every stage is ONE launch over all graphs: the clean
    # kernel-only / roofline measurement) and one cut into graph groups for the H2D/render/D2H pipeline (e2e).
    eng.set_option(pkg.OPT_PIPELINE_GROUPS, 1)
    ctxs = build_c2_batch(pkg, eng.backend, n_graphs, length, seed_base=rank * n_graphs)
    batch = pkg.Batch(ctxs)
    eng.set_option(pkg.OPT_PIPELINE_GROUPS, args.groups)
    t_prep = time.perf_counter()
    batch_e2e = pkg.Batch(ctxs)  # wae_batch_prepare: planning + allocation + the first upload of the source PCM
    prepare_ms = (time.perf_counter() - t_prep) * 1e3
    stats0 = batch.stats()
    out_floats = n_graphs * 2 * length
    host_out = torch.empty(out_floats, dtype=torch.float32, pin_memory=True)
    host_out_ptr = ctypes.c_void_p(host_out.data_ptr())
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        batch.sync()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    # ---- kernel-only: inputs resident in HBM (3.84 GB of source PCM per 1000 graphs >> 126 MB L2: every step
    # streams its inputs from HBM again, no explicit L2 flush needed)
    batch.set_timing(True)
    for _ in range(args.warmup):
        batch.run()
    batch.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total, _ = timed(batch.run, args.steps)
    clocks = sampler.stop()
    stage_times = batch.stage_times()  # last run of the timed region
    stats = batch.stats()
    ms_per_step = ms_total / args.steps
    total_quanta = n_graphs * quanta_per_graph * world
    value = total_quanta / (ms_per_step * 1e-3)

    # ---- end to end through the host API: H2D (pinned source PCM) + render + D2H (pinned), every step
    def e2e_step():
        # per graph group: H2D (pinned source PCM) -> render -> D2H (pinned output), overlapped on three streams
        batch_e2e.run_pipelined(host_out_ptr)

    batch.set_timing(False)
    batch.sync()
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    e2e_steps = max(1, min(args.steps, 3))
    _, e2e_wall = timed(e2e_step, e2e_steps)
    e2e_s = e2e_wall / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_quanta / e2e_s
    h2d = stats.asset_bytes * world  # whole job, like `value`: every rank copies its own shard over its own PCIe link
    d2h = out_floats * 4 * world

    # ---- roofline of the dominant kernel (CUDA events around every stage launch, on the launching stream)
    peak, peak_src = load_peaks()
    agg, launches_of = {}, {}
    for name, ms, _n in stage_times:  # one entry per (graph group, stage): aggregate by kernel
        agg[name] = agg.get(name, 0.0) + ms
        launches_of[name] = launches_of.get(name, 0) + int(stats.chunks)
    dom_name = max(agg, key=agg.get) if agg else ""
    dom = (dom_name, agg.get(dom_name, 0.0), 0)
    n_chunks = launches_of.get(dom_name, 1)
    # C2 = 2048 B per graph-quantum (1024 B source read + 1024 B destination write); one launch of the
    # dominant kernel covers all graphs of the batch for one chunk
    alg_bytes_step = 2048 * n_graphs * quanta_per_graph
    alg_bytes_launch = alg_bytes_step / n_chunks
    dom_ms_launch = dom[1] / n_chunks if n_chunks else 0.0
    achieved = alg_bytes_launch / (dom_ms_launch * 1e-3) / 1e9 if dom_ms_launch > 0 else 0.0
    # DRAM traffic of the same launch from an `ncu` capture of this very command (measured/dram_traffic.json, written by
    # tools/record_traffic.py on the GPU box); null when no capture matches the workload size
    traffic = None
    try:
        rec = json.load(open(os.path.join(ROOT, "web-audio-api-rs_b200", "measured", "dram_traffic.json")))
        if rec.get("kernel") == dom[0] and rec.get("graphs") == n_graphs and rec.get("frames_per_graph") == length:
            traffic = rec["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_launch,
                "kernel_ms_per_launch": dom_ms_launch, "kernel_share_of_step": dom[1] / ms_per_step if ms_per_step else None,
                "step_achieved_gbs": alg_bytes_step / (ms_per_step * 1e-3) / 1e9,
                "launches_per_step": n_chunks, "stages_ms_per_step": {n: round(ms, 4) for n, ms in agg.items()}}

    # ---- the final gather of rendered PCM over NCCL (north_star), timed once, outside the steps: in deployment
    # every rank returns its own shard over its own PCIe link, so the gather is reported, not part of `value`
    gather_ms = None
    if world > 1:
        p, nfl = batch.device_ptr()

        class _W:  # the engine's output buffer as a torch tensor (CUDA array interface, no copy)
            __cuda_array_interface__ = {"shape": (n_graphs, 2, length), "typestr": "<f4", "data": (p, False), "version": 2}
        shard = torch.as_tensor(_W(), device="cuda")
        pkg.parallel.gather_pcm(shard[:1], world, dst=0)  # communicator set-up outside the timed gather
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = pkg.parallel.gather_pcm(shard, n_graphs * world, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = pkg.parallel.max_over_ranks(g0.elapsed_time(g1), device="cuda")
        del full

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample of the same workload (rank 0 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = pkg.context.Backend(pkg.Api(ctypes.CDLL(ge.ORACLE_SO), "wao_"))
        cores = os.cpu_count() or 1
        ns = min(args.cpu_sample_graphs, n_graphs)  # bounded sample of the same workload
        octx = build_c2_batch(pkg, oracle, ns, length)
        arr = (ctypes.c_void_p * ns)(*[c._g for c in octx])
        out = np.empty((ns, 2, length), np.float32)
        secs = ctypes.c_double()
        oracle.api.check(oracle.api.render_many(arr, ns, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), cores, ctypes.byref(secs)))
        walls = [secs.value]
        for _ in range(2):  # the sample takes well under a second on a many-core host: median of three fresh renders
            octx = build_c2_batch(pkg, oracle, ns, length)
            arr = (ctypes.c_void_p * ns)(*[c._g for c in octx])
            oracle.api.check(oracle.api.render_many(arr, ns, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), cores, ctypes.byref(secs)))
            walls.append(secs.value)
        wall = float(np.median(walls))
        cpu_baseline = {"value": ns * quanta_per_graph / wall, "unit": "graph-quanta/s", "cores": cores, "kind": "port",
                        "sample": f"{ns} graphs x {args.seconds:.0f} s of the same workload, one context per worker thread, "
                                  f"median of 3 renders: {wall:.3f} s wall ({ns * wall:.1f} core-seconds upper bound)"}
        # parity spot check of the bench output itself against the oracle (first graphs)
        got = host_out.numpy().reshape(n_graphs, 2, length)[:ns]
        cpu_baseline["max_abs_diff_vs_gpu"] = float(np.abs(got - out).max())

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "graph-quanta/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 filter state / f32 PCM", "data": "synthetic",
            "config": {"workload": "C2 (BASELINE configs[1]): %d OfflineAudioContexts/GPU, AudioBufferSource->Biquad->Gain->"
                                   "destination, 48 kHz stereo, %.0f s each" % (n_graphs, args.seconds),
                       "graphs_per_gpu": n_graphs, "frames_per_graph": length, "chunk_frames": int(stats0.chunks and (length + 127) // 128 * 128 // stats0.chunks),
                       "l2": "inputs (%.2f GB/GPU) larger than L2, no flush" % (stats.asset_bytes / 1e9),
                       "sharding": "independent graphs per rank, no data-path collective", "e2e_pipeline_groups": args.groups},
            "samples_per_sec": value * 128, "prepare_ms_once": prepare_ms, "gpu_launches": int(stats.kernel_launches_per_run) * args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "graph-quanta/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if gather_ms is not None:
            line["nccl_gather_pcm_ms"] = gather_ms
        if world == 1 and args.extra:
            batch.destroy()
            batch_e2e.destroy()
            del host_out
            oracle2 = None if args.no_cpu_baseline else pkg.context.Backend(pkg.Api(ctypes.CDLL(ge.ORACLE_SO), "wao_"))
            line["other_workloads"] = run_extra_workloads(pkg, eng, oracle2, os.cpu_count() or 1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
