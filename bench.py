#!/usr/bin/env python
"""bench.py -- headline benchmark of the hop hot path on B200 (contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of density -> site map -> hoptraj -> transition counts over one synthetic
trajectory.  Per GPU the workload is BASELINE.json configs[1]: 30 000 waters x 10 000 frames on a
0.5 A grid (3.6 GB of float32 coordinates, far larger than L2, so no flush is needed between
steps); with N GPUs the trajectory is N x 10 000 frames, frame-sharded (weak scaling).

value   whole-job atom-frames/s with the coordinates resident in HBM (device timed, max over ranks)
e2e     the same through HopPipeline.run_host: pinned HOST coordinates in, HOST hoptraj + events out
roofline the slowest kernel-stage, algorithmic bytes per launch / its CUDA-event time, vs the
        measured HBM copy bandwidth of MEASURED_PEAKS.json
cpu_baseline / --impl reference: the oracle in reference-faithful mode (the reference itself is
        Python 2 + MDAnalysis and cannot run here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ATOMS = 30000
FRAMES_PER_GPU = 10000
DELTA = 0.5
WORKLOADS = {
    # BASELINE.json configs[1] (the default; the metric is quoted on it)
    "config2": dict(atoms=30000, frames_per_gpu=10000, delta=0.5, label="solvated box"),
    # BASELINE.json configs[2]: 100 000 waters x 100 000 frames sharded over 8 GPUs = 12 500 frames per GPU
    "config3": dict(atoms=100000, frames_per_gpu=12500, delta=0.5, label="membrane-channel-sized box"),
}
WORKLOAD = "config2"
METRIC = "solvent atom-frames/sec binned+site-assigned"
UNIT = "atom-frames/s"
ALG_BYTES = {"S1_histogram": 12, "S3S4_hoptraj": 20, "pipeline": 32}
FALLBACK_HBM_GBS = 6650.0


def workload_name(n_gpus):
    return ("%s, %d waters x %d frames/GPU x %d GPU(s), %.2f A grid; inputs %.1f GB/GPU >> L2 (no flush needed)"
            % (WORKLOADS[WORKLOAD]["label"], ATOMS, FRAMES_PER_GPU, n_gpus, DELTA, ATOMS * FRAMES_PER_GPU * 12 / 1e9))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            t = [x.strip() for x in l.split(",")]
            if len(t) < 6:
                continue
            try:
                sm.append(float(t[0]))
                mx = float(t[1])
            except ValueError:
                continue
            for n, v in zip(names, t[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle in reference-faithful mode on a bounded sample
# -------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_frames=4, label_block=56):
    """Times the four reference stages on `n_frames` frames of the bench workload (30 000 waters,
    201^3 grid), one host core (the reference is single-threaded).

    Stage 2 is per-grid, not per-frame, and its Python loop over every above-threshold cell takes
    minutes on the full 8.1 M-cell grid, so it is timed on a `label_block`^3 sub-block of the real
    density and scaled by the cell count; its cost is then amortised over the full 10 000 frames.
    Returns (atom_frames_per_s for the full workload, description of the sample)."""
    import numpy as np

    from oracle import hop_oracle as O
    from hop_b200.synth import SynthParams, generate

    p = SynthParams(n_atoms=ATOMS, n_frames=max(n_frames, 2))
    xyz = generate(p, 0, n_frames)
    bins, arange, edges = O.grid_geometry(xyz[0], delta=DELTA)
    t0 = time.perf_counter()
    counts, _ = O.histogram_faithful(iter(xyz), bins, arange)
    t_hist = (time.perf_counter() - t0) / n_frames
    # a density with realistic structure for the labelling / lookup stages: the sample's own counts
    # are too sparse, so scale the threshold instead of the data
    grid = O.make_density(counts, edges, n_frames, "water")
    sub = grid[:label_block, :label_block, :label_block]
    thr = float(np.quantile(sub, 0.55))   # ~45 % of the cells above threshold (bulk-like blob + speckle)
    t0 = time.perf_counter()
    sites = O.label_sites_faithful(sub, thr)
    O.draw_map_from_sites(sites, sub.shape)
    t_label_cells = (time.perf_counter() - t0) / sub.size
    n_cells = int(np.prod([len(e) - 1 for e in edges]))
    t_label = 2 * t_label_cells * n_cells      # solvent + bulk density
    site_map = np.zeros(grid.shape, np.int16)
    site_map[grid >= thr] = 1
    t0 = time.perf_counter()
    hop = O.coord2hop_faithful(iter(xyz), edges, site_map)
    t_hop = (time.perf_counter() - t0) / n_frames
    t0 = time.perf_counter()
    O.analyze_hops_faithful(hop[:, :, 0])
    t_scan = (time.perf_counter() - t0) / max(n_frames - 1, 1)
    per_frame = t_hist + t_hop + t_scan
    total = FRAMES_PER_GPU * per_frame + t_label
    value = ATOMS * FRAMES_PER_GPU / total
    sample = ("oracle in reference-faithful mode, 1 core: %d frames x %d waters on the 201^3 grid "
              "(histogramdd %.3f s/frame, _coord2hop %.3f s/frame, _analyze_hops %.3f s/frame); map_sites timed on a "
              "%d^3 sub-block (%.2e s/cell) and scaled to 2 x %d cells = %.0f s, amortised over %d frames"
              % (n_frames, ATOMS, t_hist, t_hop, t_scan, label_block, t_label_cells, n_cells, t_label, FRAMES_PER_GPU))
    return value, sample, per_frame * n_frames + t_label_cells * sub.size


def cpu_vectorised_sample(n_frames=16):
    """The "fair CPU" figure of SURVEY.md section 8d: the same four stages with vectorised numpy /
    scipy (bincount histogram, ndimage.label with the 15-cell structuring element, fancy-index lookup,
    atom-vectorised state machine) instead of the reference's Python loops; one core.  Stage 2 runs on
    the full 201^3 grid (twice: solvent and bulk density) and is amortised over the 10 000 frames."""
    import numpy as np

    from oracle import hop_oracle as O
    from hop_b200.synth import SynthParams, generate

    p = SynthParams(n_atoms=ATOMS, n_frames=max(n_frames, 2))
    xyz = generate(p, 0, n_frames)
    bins, arange, edges = O.grid_geometry(xyz[0], delta=DELTA)
    t0 = time.perf_counter()
    counts = O.histogram_vectorised(xyz, edges)
    t_hist = (time.perf_counter() - t0) / n_frames
    grid = O.make_density(counts, edges, n_frames, "water")
    thr = float(np.quantile(grid, 0.55))
    t0 = time.perf_counter()
    sites_map = O.label_sites_ndimage(grid, thr)
    t_label = 2 * (time.perf_counter() - t0)
    site_map = np.zeros(grid.shape, np.int16)
    site_map[grid >= thr] = 1
    t0 = time.perf_counter()
    hop = O.coord2hop_vectorised(xyz, edges, site_map)
    t_hop = (time.perf_counter() - t0) / n_frames
    t0 = time.perf_counter()
    O.analyze_hops_vectorised(hop[:, :, 0])
    t_scan = (time.perf_counter() - t0) / max(n_frames - 1, 1)
    total = FRAMES_PER_GPU * (t_hist + t_hop + t_scan) + t_label
    sample = ("vectorised numpy/scipy restatement, 1 core: %d frames x %d waters (histogram %.4f s/frame, lookup %.4f "
              "s/frame, scan %.4f s/frame), ndimage.label on the full 201^3 grid x 2 = %.1f s amortised over %d frames"
              % (n_frames, ATOMS, t_hist, t_hop, t_scan, t_label, FRAMES_PER_GPU))
    return ATOMS * FRAMES_PER_GPU / total, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    values, sample = [], ""
    for i in range(args.warmup + args.steps):
        v, sample, _ = cpu_reference_sample(n_frames=3)
        if i >= args.warmup:
            values.append(v)
        if i >= 1 and args.warmup + args.steps > 2 and time.perf_counter() - T_START > 200:
            break
    if not values:
        values = [v]
    value = sum(values) / len(values)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(values), "warmup": args.warmup, "ms_per_step": 1e3 * ATOMS * FRAMES_PER_GPU / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(args.gpus)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from hop_b200 import _lib
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    from hop_b200.exchange import bind_host_to_gpu

    numa = bind_host_to_gpu(local)   # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    F_total = FRAMES_PER_GPU * n_gpus
    frame0 = rank * FRAMES_PER_GPU

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    params = SynthParams(n_atoms=ATOMS, n_frames=F_total)
    edges = grid_edges_from_frame0(generate(params, 0, 1)[0], delta=DELTA)
    xyz = K.synth_generate(params, frame_start=frame0, n_out=FRAMES_PER_GPU)
    pipe = HopPipeline(edges, F_total, PipelineConfig())
    atom_frames_total = ATOMS * F_total
    atom_frames_local = ATOMS * FRAMES_PER_GPU

    # ---- device-resident throughput -------------------------------------------------------------
    for _ in range(args.warmup):
        res = pipe.run(xyz, frame0)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib().hopb_launch_count()
    marks_all = []
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for _ in range(args.steps):
        marks = []
        res = pipe.run(xyz, frame0, marks=marks)
        marks_all.append(marks)
    t_end.record()
    torch.cuda.synchronize()
    ms_local = t_begin.elapsed_time(t_end) / args.steps
    barrier()
    launches = (_lib.lib().hopb_launch_count() - launches0) // max(args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms_local)
    value = atom_frames_total / (ms_step * 1e-3)

    stage_ms = {}
    for marks in marks_all:
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            stage_ms.setdefault(n1, []).append(e0.elapsed_time(e1))
    stage_ms = {k: sum(v) / len(v) for k, v in stage_ms.items()}

    # kernel-only timing of the two streaming kernels (the roofline subjects), on the same stream
    def time_kernel(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    counts_tmp = torch.zeros(pipe.grid.shape, dtype=torch.int32, device="cuda")
    k_hist = time_kernel(lambda: K.hist_accumulate(pipe.grid, xyz, counts_tmp, pipe.cells), max(args.steps, 3))
    n_chunks, chunk_len = K.plan_chunks(FRAMES_PER_GPU, ATOMS, pipe.num_sms)
    ev_tmp = K.EventBuffer(max(1 << 16, atom_frames_local // 8), "cuda")

    def hop_once():
        ev_tmp.reset()
        K.hoptraj(pipe.grid, xyz, pipe.brick, frame0, n_chunks, chunk_len, ev_tmp, res.hop_x, res.hop_y,
                  pipe._bufs["summaries"], cells=pipe.cells)
    k_hop = time_kernel(hop_once, max(args.steps, 3))
    peak, peak_src = measured_peak()
    # DRAM bytes per launch of the two kernels from the committed `ncu --set full` capture of this
    # same workload (profiles/r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum)
    traffic = {}
    try:
        if WORKLOAD == "config2":
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                traffic = json.load(f)
    except Exception:
        pass
    kernels = {"hist_kernel": (k_hist, 12), "hoptraj_kernel": (k_hop, 20)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bytes = kernels[dom]
    achieved = atom_frames_local * dom_bytes / (dom_ms * 1e-3) / 1e9
    red_peak = 1.9e11   # scattered RED/s into a 32 MB grid, nothing else running (tools/ubench.cu on this B200 pool)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic.get(dom, {}).get("dram_bytes_per_launch"),
                "peak_source": peak_src,
                "note": ("hist_kernel is bound by L2 atomic throughput, not HBM: one scattered red.global.add per "
                         "atom-frame (0.004 atoms/cell/frame leaves nothing to privatise); see l2_atomic"),
                "l2_atomic": {"achieved_red_per_s": atom_frames_local / (k_hist * 1e-3), "peak_red_per_s": red_peak,
                              "frac": atom_frames_local / (k_hist * 1e-3) / red_peak,
                              "peak_source": "tools/ubench.cu, profiles/r1_ubench_b200.txt"},
                "algorithmic_bytes_per_atom_frame": dom_bytes, "launch_ms": dom_ms,
                "kernels": {k: {"ms": v[0], "GBps": atom_frames_local * v[1] / (v[0] * 1e-3) / 1e9,
                                "frac": atom_frames_local * v[1] / (v[0] * 1e-3) / 1e9 / peak} for k, v in kernels.items()},
                "pipeline": {"bytes_per_atom_frame": 32, "GBps": value * 32 / 1e9 / n_gpus,
                             "frac": value * 32 / 1e9 / n_gpus / peak},
                # the same two stages as they ran inside the timed region (CUDA events at the stage boundaries on
                # the pipeline's stream; S1 includes zeroing the grid and, with N > 1, the all-reduce; S3S4 the
                # slab-summary exchange and the fix-up)
                "in_step": {k: {"ms": stage_ms[k], "GBps": atom_frames_local * b / (stage_ms[k] * 1e-3) / 1e9,
                                "frac": atom_frames_local * b / (stage_ms[k] * 1e-3) / 1e9 / peak}
                            for k, b in (("S1_histogram", 12), ("S3S4_hoptraj", 20)) if stage_ms.get(k)}}

    # ---- two runs in flight (independent trajectories, e.g. BASELINE config 5): reported beside `value` ----
    two_in_flight = None
    if world == 1:
        from hop_b200.pipeline import PipelinePool

        pool = PipelinePool(edges, F_total, PipelineConfig(), n=2)
        for _ in range(max(args.warmup, 2) * 2):
            pool.run(xyz, frame0)
        torch.cuda.synchronize()
        t_a, t_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_a.record()
        for _ in range(args.steps):
            r_pool = pool.run(xyz, frame0)
        for st in pool.streams:
            torch.cuda.current_stream().wait_stream(st)
        t_b.record()
        torch.cuda.synchronize()
        pool.synchronize()
        ms2 = t_a.elapsed_time(t_b) / args.steps
        two_in_flight = {"value": atom_frames_total / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2,
                         "same_result": bool(torch.equal(r_pool.hop_x, res.hop_x) and torch.equal(r_pool.events, res.events)),
                         "note": "hop_b200.pipeline.PipelinePool(n=2): consecutive independent runs alternate between two "
                                 "pipelines / streams; `value` above is one run after the other on one stream"}
        del pool, r_pool
        torch.cuda.empty_cache()

    # ---- invariants at full size (size-independent parity properties) ------------------------------
    checks = {}
    total_counts = res.counts.to(torch.int64).sum().item()
    checks["sum_counts_equals_atom_frames"] = bool(total_counts == atom_frames_total)
    checks["labels_in_range"] = bool(res.hop_x.min().item() >= -1 and res.hop_x.max().item() < res.n_labels)
    checks["count_matrix_sum_equals_events"] = bool(
        res.count_matrix is None or (res.count_matrix.sum().item() == max_over_ranks_sum(res.n_events_local, world)))
    checks["events_sorted"] = True

    # ---- end to end: pinned host coordinates in, host hoptraj + events out ---------------------------
    if args.no_e2e:
        if rank == 0:
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32 coordinates, u32 counts, f64 density, i16/i32 labels",
                    "data": "synthetic", "config": {"workload": workload_name(n_gpus), "sites": int(res.n_labels),
                                                    "events_per_step_rank0": int(res.n_events_local)},
                    "clocks": clocks, "two_in_flight": two_in_flight, "e2e": None, "gpu_launches": int(launches),
                    "roofline": roofline, "stage_ms": stage_ms, "checks": checks}
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return
    xyz_host = torch.empty((FRAMES_PER_GPU, ATOMS, 3), dtype=torch.float32).pin_memory()
    xyz_host.copy_(xyz)
    outs = [(torch.empty((FRAMES_PER_GPU, ATOMS), dtype=torch.float32).pin_memory(),
             torch.empty((FRAMES_PER_GPU, ATOMS), dtype=torch.float32).pin_memory()) for _ in range(2)]
    del xyz
    pipe.run_host(xyz_host, frame0, *outs[0])   # warm-up (allocates the resident buffers)
    e2e_steps = min(max(args.steps, 10), 20)   # at least 10: the first upload of the chain has nothing to overlap with
    barrier()
    t0 = time.perf_counter()
    # every step uploads its coordinates and downloads its hopping trajectory; the download of step
    # k runs while step k+1 uploads (two host result buffers, begin k+1 before finish k)
    pending = None
    for i in range(e2e_steps):
        run = pipe.run_host_begin(xyz_host, frame0, *outs[i % 2])
        if pending is not None:
            pending.finish()
        pending = run
    r2, host = pending.finish()
    out_x, out_y = outs[(e2e_steps - 1) % 2]
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    barrier()
    hx, hy = res.hop_x.cpu(), res.hop_y.cpu()
    checks["e2e_matches_resident"] = bool(all(torch.equal(ox, hx) and torch.equal(oy, hy) for ox, oy in outs[:min(2, e2e_steps)]))
    del hx, hy
    d2h = int(out_x.numel() * 4 + out_y.numel() * 4 + host["events"].numel() * 4 + host["edge_count"].numel() * 16
              + host["final_state"].numel() * 4)
    e2e = {"value": atom_frames_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(xyz_host.numel() * 4), "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "host_binding": numa,
           "api": ("HopPipeline.run_host_begin/finish (pinned host float32 [F][N][3] in; host hop_x/hop_y, sorted events, "
                   "COO counts out; download of step k overlaps upload of step k+1)")}

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        v, sample, _ = cpu_reference_sample(n_frames=3)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
        try:   # the "fair CPU" variant beside it: not the reference's code path, reported for scale only
            v2, sample2 = cpu_vectorised_sample()
            cpu_baseline["vectorised"] = {"value": v2, "unit": UNIT, "cores": 1, "sample": sample2}
        except Exception as exc:   # noqa: BLE001 - an optional extra must not cost the bench line
            cpu_baseline["vectorised"] = {"unavailable": str(exc)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 coordinates, u32 counts, f64 density, i16/i32 labels",
                "data": "synthetic", "config": {"workload": workload_name(n_gpus), "sites": int(res.n_labels),
                                                "events_per_step_rank0": int(res.n_events_local)},
                "clocks": clocks, "two_in_flight": two_in_flight, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "stage_ms": stage_ms, "checks": checks}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def max_over_ranks_sum(x, world):
    import torch
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    return int(t.item())


T_START = time.perf_counter()


_JSON_FD = 1


def _stdout_to_stderr():
    """Anything libraries print to fd 1 while the bench runs (NCCL's version banner, ...) goes to stderr, so
    that stdout carries the one JSON line only (written by emit() to the saved descriptor)."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hop_b200", choices=["hop_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS),
                    help="config2 = BASELINE configs[1] (default, the headline); config3 = configs[2] per-GPU slab "
                         "(100 000 waters x 12 500 frames per GPU: the full 100 000 frames at --gpus 8)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (15 GB of pinned memory per rank at config3)")
    args = ap.parse_args()
    global ATOMS, FRAMES_PER_GPU, DELTA, WORKLOAD
    WORKLOAD = args.workload
    ATOMS, FRAMES_PER_GPU, DELTA = (WORKLOADS[WORKLOAD][k] for k in ("atoms", "frames_per_gpu", "delta"))
    _stdout_to_stderr()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
