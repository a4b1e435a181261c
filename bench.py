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
    # BASELINE.json configs[2] in full on ONE GPU (north_star's target sentence): 100 000 waters x 100 000 frames = 1e10
    # atom-frames, 120 GB of coordinates, streamed through HBM in frame chunks (HopPipeline.run_streamed)
    # BASELINE.json configs[3]: three ion species in a 200k-atom box (126 A), 10^6 frames, 1.0 A grids, one coordinate stream
    "config4": dict(atoms=1200, frames_per_gpu=1000000, delta=1.0, label="ion hopping, 3 species x 400 ions in a 126 A box"),
    # BASELINE.json configs[4]: 64 trajectories x 20 000 frames at 0.25 A (402^3 grid), 8 trajectories per GPU, no collective
    "config5": dict(atoms=30000, frames_per_gpu=20000, delta=0.25, label="many-densities batch, 8 trajectories per GPU"),
    "config3full": dict(atoms=100000, frames_per_gpu=100000, delta=0.5, label="membrane-channel-sized box, chunk-streamed"),
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
CPU_FRAMES_PER_STEP = 50      # at least 50 frames per per-frame stage
CPU_LABEL_BLOCK = 64          # stage 2 is timed on a block of this edge (cells), see cpu_stage2_block


def pin_one_core():
    """The reference has no threading or multiprocessing: the CPU arm runs on ONE host core.
    Returns (previous affinity, description)."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, {allowed[-1]})
        try:
            from threadpoolctl import threadpool_limits

            threadpool_limits(1)
        except Exception:
            pass
        return set(allowed), "pinned to cpu %d of %d allowed (sched_setaffinity), BLAS/OpenMP pools limited to 1 thread" % (
            allowed[-1], len(allowed))
    except Exception as exc:  # noqa: BLE001
        return None, "not pinned (%s)" % (str(exc)[:80],)


def unpin(previous):
    if previous:
        try:
            os.sched_setaffinity(0, previous)
        except Exception:
            pass


def cpu_stage2_block(density, thresholds, block=CPU_LABEL_BLOCK):
    """Density.map_sites as the reference runs it (Python loop over the above-threshold cells x 7 offsets
    into a networkx graph, connected components, size sort, per-cell repaint: hop/sitemap.py:1496-1587) on
    a `block`^3 block of the full-grid density at the real thresholds.  The block starts 8 cells below the
    box centre, so it holds part of the hydration-site region and bulk water in about the proportion of the
    whole box.  Returns (seconds for all thresholds, cells of the block, above-threshold cells per threshold)."""
    import numpy as np

    from oracle import hop_oracle as O

    o = [max(0, min(n // 2 - 8, n - block)) for n in density.shape]
    sub = np.ascontiguousarray(density[o[0]:o[0] + block, o[1]:o[1] + block, o[2]:o[2] + block])
    t0 = time.perf_counter()
    filled = []
    for thr in thresholds:
        sites = O.label_sites_faithful(sub, thr)
        O.draw_map_from_sites(sites, sub.shape)
        filled.append(int((sub >= thr).sum()))
    return time.perf_counter() - t0, int(sub.size), filled


def cpu_per_frame_stages(xyz, bins, arange, edges, site_map):
    """The three per-frame stages of the reference on the frames `xyz` [S][N][3]: numpy.histogramdd per frame
    (hop/density.py:125-131), _coord2hop per frame (hop/trajectory.py:403-468: full map copy + per-atom
    lookup) and the per-atom dict state machine of _analyze_hops (hop/graph.py:212-274).  Returns seconds per
    stage."""
    from oracle import hop_oracle as O

    t0 = time.perf_counter()
    O.histogram_faithful(iter(xyz), bins, arange)
    t1 = time.perf_counter()
    hop = O.coord2hop_faithful(iter(xyz), edges, site_map)
    t2 = time.perf_counter()
    O.analyze_hops_faithful(hop[:, :, 0])
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


def rough_site_map(density, solvent_threshold, bulk_threshold):
    """Input data for the CPU timing of stage 3 when no GPU result is at hand: bulk mask = 1, hydration
    sites numbered 2.. by scipy (not the canonical numbering -- the lookup cost does not depend on it)."""
    import numpy as np
    from scipy import ndimage

    from oracle import hop_oracle as O

    st = np.zeros((3, 3, 3), bool)
    st[1, 1, 1] = True
    for off in O.DELTA_FIRST_OCTANT:
        st[tuple(1 + off)] = st[tuple(1 - off)] = True
    lab, _ = ndimage.label(density >= solvent_threshold, structure=st)
    m = np.where(density >= bulk_threshold, 1, 0).astype(np.int16)
    m[lab > 0] = (lab[lab > 0] + 1).astype(np.int16)
    return m


def expected_density(params, edges, unit="water"):
    """Long-run (infinite-frame) density of the synthetic model of hop_b200/synth.py on the grid `edges`, in
    closed form: bulk atoms uniform in the box outside the central cube, bound atoms a Gaussian of width
    sigma_site on their site centre, occupied p_rebind / (p_rebind + p_hop / 2) of the time.  Input data for
    the CPU timing of stages 2-4 when the full-trajectory density is out of the CPU's reach (the reference arm
    has no GPU result to borrow one from): it has the structure of the 10 000-frame density -- solid bulk,
    compact hydration sites -- which a density from the few hundred sample frames has not (Poisson noise
    riddles its bulk site with holes, and atoms flickering between bulk and interstitial make the
    reference's state machine several times slower than it is on real input)."""
    import numpy as np

    from oracle import hop_oracle as O

    p = params
    centres, half = p.lattice()
    L = float(p.box)
    n_bound = p.n_bound
    rho_bulk = (p.n_atoms - n_bound) / (L ** 3 - (2.0 * half) ** 3)
    mids = [0.5 * (e[:-1] + e[1:]) for e in edges]
    c = 0.5 * L
    inside = [(m >= 0) & (m < L) for m in mids]
    in_cube = [np.abs(m - c) < half for m in mids]
    box = inside[0][:, None, None] & inside[1][None, :, None] & inside[2][None, None, :]
    cube = in_cube[0][:, None, None] & in_cube[1][None, :, None] & in_cube[2][None, None, :]
    rho = np.where(box & ~cube, rho_bulk, 0.0)
    occ = n_bound * (p.p_rebind / (p.p_rebind + 0.5 * p.p_hop)) / len(centres)
    s = float(p.sigma_site)
    norm = occ / ((2.0 * np.pi) ** 1.5 * s ** 3)
    r = int(np.ceil(4.0 * s / float(edges[0][1] - edges[0][0])))
    for ctr in np.asarray(centres, dtype=np.float64):
        idx = [int(np.searchsorted(e, x, side="right")) - 1 for e, x in zip(edges, ctr)]
        sl = [slice(max(0, i - r), min(len(m), i + r + 1)) for i, m in zip(idx, mids)]
        gx, gy, gz = (np.exp(-0.5 * ((m[q] - x) / s) ** 2) for m, q, x in zip(mids, sl, ctr))
        rho[sl[0], sl[1], sl[2]] += norm * gx[:, None, None] * gy[None, :, None] * gz[None, None, :]
    return rho * O.density_conversion_factor("Angstrom^{-3}", unit)


def describe_cpu_sample(n_frames, t_hist, t_hop, t_scan, t_block, block_cells, filled, shape, pin, density_source):
    n_cells = 1
    for n in shape:
        n_cells *= int(n)
    t_label_full = t_block * n_cells / block_cells
    return ("oracle in reference-faithful mode (kind=port: the reference is Python 2 + MDAnalysis and cannot run on this "
            "box), 1 core, %s; per-frame stages timed on %d frames x %d waters on the full %s grid: histogramdd %.4f "
            "s/frame, _coord2hop %.4f s/frame, _analyze_hops %.4f s/frame; map_sites (solvent + bulk threshold) timed on "
            "a %d^3 block of %s at the real thresholds (%s of %d cells above threshold) = %.1f s, scaled by cells to the "
            "full grid = %.0f s and amortised over the %d frames of the workload"
            % (pin, n_frames, ATOMS, "x".join(str(int(n)) for n in shape), t_hist / n_frames,
               t_hop / n_frames, t_scan / n_frames, CPU_LABEL_BLOCK, density_source, "+".join(str(f) for f in filled),
               block_cells, t_block, t_label_full, FRAMES_PER_GPU))


def cpu_vectorised_sample(n_frames=16):
    """The "fair CPU" figure of SURVEY.md section 8d: the same four stages with vectorised numpy /
    scipy (bincount histogram, ndimage.label with the 15-cell structuring element, fancy-index lookup,
    atom-vectorised state machine) instead of the reference's Python loops; one core.  Stage 2 runs on
    the full grid (twice: solvent and bulk density) and is amortised over the frames of the workload."""
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
    O.label_sites_ndimage(grid, thr)
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
              "s/frame, scan %.4f s/frame), ndimage.label on the full grid x 2 = %.1f s amortised over %d frames"
              % (n_frames, ATOMS, t_hist, t_hop, t_scan, t_label, FRAMES_PER_GPU))
    return ATOMS * FRAMES_PER_GPU / total, sample


def cpu_baseline_leg(density, site_map, edges, solvent_threshold, bulk_threshold):
    """cpu_baseline of the GPU arm (rank 0, N = 1): one CPU step on CPU_FRAMES_PER_STEP frames of the bench
    workload with the REAL full-trajectory density and site map (downloaded from the GPU run: input data of
    the CPU timing, the GPU is not on the timed path)."""
    import numpy as np

    from oracle import hop_oracle as O
    from hop_b200.synth import SynthParams, generate

    prev, pin = pin_one_core()
    try:
        S = CPU_FRAMES_PER_STEP
        xyz = generate(SynthParams(n_atoms=ATOMS, n_frames=FRAMES_PER_GPU), 0, S)
        bins, arange, e2 = O.grid_geometry(xyz[0], delta=DELTA)
        assert all(np.array_equal(a, b) for a, b in zip(edges, e2))
        t_hist, t_hop, t_scan = cpu_per_frame_stages(xyz, bins, arange, edges, site_map)
        t_block, block_cells, filled = cpu_stage2_block(density, (solvent_threshold, bulk_threshold))
        n_cells = int(density.size)
        t_label_full = t_block * n_cells / block_cells
        t_step = t_hist + t_hop + t_scan
        value = ATOMS * S / (t_step + t_label_full * S / FRAMES_PER_GPU)
        out = {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": describe_cpu_sample(S, t_hist, t_hop, t_scan, t_block, block_cells, filled, density.shape, pin,
                                             "the full-trajectory density of this run"),
               "timed_s": t_step + t_block,
               "extrapolated": {"full_workload_s": FRAMES_PER_GPU * t_step / S + t_label_full,
                                "stage2_full_grid_s": t_label_full,
                                "note": "per-frame stages scale linearly in frames (measured on %d); stage 2 scaled by cells" % S}}
        try:   # the "fair CPU" variant beside it: not the reference's code path, reported for scale only
            v2, sample2 = cpu_vectorised_sample()
            out["vectorised"] = {"value": v2, "unit": UNIT, "cores": 1, "sample": sample2}
        except Exception as exc:   # noqa: BLE001 - an optional extra must not cost the bench line
            out["vectorised"] = {"unavailable": str(exc)[:200]}
        return out
    finally:
        unpin(prev)


def run_reference(args):
    """--impl reference: W warm-up + exactly K timed steps; a step = the three per-frame stages of the
    reference on CPU_FRAMES_PER_STEP consecutive frames of the bench workload (different frames every step).
    Stage 2 (per grid, not per frame) is timed once on a block at the real thresholds and its full-grid
    cost enters `value` pro rata; `ms_per_step` is what a step actually took."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from oracle import hop_oracle as O
    from hop_b200.synth import SynthParams, generate

    prev, pin = pin_one_core()
    S = CPU_FRAMES_PER_STEP
    n_steps = args.warmup + args.steps
    xyz = generate(SynthParams(n_atoms=ATOMS, n_frames=FRAMES_PER_GPU), 0, S * n_steps)
    bins, arange, edges = O.grid_geometry(xyz[0], delta=DELTA)
    # the site map the lookups run against and the density stage 2 is timed on: the closed-form long-run
    # density of the synthetic model (expected_density) at the real thresholds
    import math

    thr_s, thr_b = math.e, math.exp(-0.5)
    density = expected_density(SynthParams(n_atoms=ATOMS, n_frames=FRAMES_PER_GPU), edges)
    site_map = rough_site_map(density, thr_s, thr_b)
    t_block, block_cells, filled = cpu_stage2_block(density, (thr_s, thr_b))
    n_cells = int(density.size)
    t_label_full = t_block * n_cells / block_cells
    times = []
    for i in range(n_steps):
        t = cpu_per_frame_stages(xyz[i * S:(i + 1) * S], bins, arange, edges, site_map)
        if i >= args.warmup:
            times.append(t)
    t_hist, t_hop, t_scan = (sum(t[k] for t in times) for k in range(3))
    n_frames = S * len(times)
    t_step = (t_hist + t_hop + t_scan) / len(times)
    value = ATOMS * S / (t_step + t_label_full * S / FRAMES_PER_GPU)
    sample = describe_cpu_sample(n_frames, t_hist, t_hop, t_scan, t_block, block_cells, filled, density.shape, pin,
                                 "the closed-form long-run density of the synthetic model (bench.py expected_density)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(args.gpus)},
            "step": "%d frames x %d waters through histogramdd, _coord2hop and _analyze_hops" % (S, ATOMS),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "extrapolated": {"full_workload_s": FRAMES_PER_GPU * t_step / S + t_label_full,
                             "stage2_full_grid_s": t_label_full, "stage2_block_s": t_block,
                             "note": ("value = atom-frames of a step / (its measured time + the pro-rata share of stage 2); "
                                      "the reference itself (Python 2.7, MDAnalysis) cannot run on this box, the port "
                                      "repeats its work pattern (oracle/hop_oracle.py *_faithful)")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    unpin(prev)
    emit(line)


# -------------------------------------------------------------------------------------------------
# config 3 in full on one GPU: the trajectory does not fit into HBM and is streamed in frame chunks
# -------------------------------------------------------------------------------------------------
def run_streamed_workload(args, chunk_frames=2500):
    import numpy as np
    import torch

    from hop_b200 import _lib
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate
    from oracle import hop_oracle as O

    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--workload config3full is the single-GPU run of configs[2]; use --workload config3 --gpus 8 for the sharded one")
    torch.cuda.set_device(0)
    F, N = FRAMES_PER_GPU, ATOMS
    params = SynthParams(n_atoms=N, n_frames=F)
    prefix = generate(params, 0, 64)
    edges = grid_edges_from_frame0(prefix[0], delta=DELTA)
    stream = K.SynthStream(params)
    gen_events = []

    def source(f0, f1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        x = stream(f0, f1)
        b.record()
        gen_events.append((a, b))
        return x

    keep = {}
    sums = []

    def sink(f0, hx, hy):
        if f0 == 0 and "x" not in keep:
            keep["x"], keep["y"] = hx[:64].clone(), hy[:64].clone()

    pipe = HopPipeline(edges, F, PipelineConfig(async_results=False, event_capacity=64 << 20))
    times, stage_ms, gen_ms_all = [], {}, []
    sampler = ClockSampler(0)
    launches = 0
    for it in range(args.warmup + args.steps):
        del gen_events[:]
        if it == args.warmup:
            sampler.start()
            launches0 = _lib.lib().hopb_launch_count()
        marks = []
        torch.cuda.synchronize()
        res = pipe.run_streamed(source, F, chunk_frames, sink=sink, marks=marks)
        torch.cuda.synchronize()
        gen_ms = sum(a.elapsed_time(b) for a, b in gen_events)
        total_ms = marks[0][1].elapsed_time(marks[-1][1])
        if it >= args.warmup:
            times.append(total_ms - gen_ms)
            gen_ms_all.append(gen_ms)
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                stage_ms.setdefault(n1, []).append(e0.elapsed_time(e1))
    launches = (_lib.lib().hopb_launch_count() - launches0) // max(args.steps, 1)
    clocks = sampler.stop()
    ms_step = sum(times) / len(times)
    gen_step = sum(gen_ms_all) / len(gen_ms_all)
    stage_ms = {k: sum(v) / len(v) for k, v in stage_ms.items()}
    stage_ms["S1_histogram"] -= gen_step        # every chunk is generated inside pass 1 (pass 2 reads the cached brick indices)
    value = N * F / (ms_step * 1e-3)
    peak, peak_src = measured_peak()
    checks = {"sum_counts_equals_atom_frames": bool(int(res.counts.to(torch.int64).sum().item()) == N * F)}
    site_map = res.map16.cpu().numpy()
    hop = O.coord2hop_vectorised(prefix, edges, site_map)
    checks["prefix64_hop_x_equals_oracle"] = bool(np.array_equal(keep["x"].cpu().numpy(), hop[:, :, 0]))
    checks["prefix64_hop_y_equals_oracle"] = bool(np.array_equal(keep["y"].cpu().numpy(), hop[:, :, 1]))
    ev_ref, _ = O.analyze_hops_vectorised(hop[:, :, 0])
    ev = K.events_to_numpy(res.events)
    sel = ev[ev["frame"] < 64]
    order = np.lexsort((sel["iatom"], sel["frame"]))
    got = np.stack([sel[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1).astype(np.int64)
    checks["prefix64_events_equal_oracle"] = bool(np.array_equal(got, ev_ref))
    checks["count_matrix_sum_equals_events"] = bool(res.count_matrix is None or int(res.count_matrix.sum().item()) == len(ev))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 coordinates, u32 counts, f64 density, i16/i32 labels, i16 hoptraj planes", "data": "synthetic",
            "config": {"workload": "%s, %d waters x %d frames on ONE GPU, %.2f A grid; 120 GB of coordinates streamed through HBM in "
                                   "%d-frame chunks (3 GB each >> L2), generated in HBM chunk by chunk" % (
                                       WORKLOADS[WORKLOAD]["label"], N, F, DELTA, chunk_frames)},
            "timing": ("CUDA events around HopPipeline.run_streamed minus the CUDA-event time of the chunk generator calls "
                       "(the generator stands for the data source); generator %.1f ms per step on top" % gen_step),
            "result": {"sites": int(res.n_labels), "events": int(len(ev))},
            "clocks": clocks, "gpu_launches": int(launches), "stage_ms": stage_ms, "checks": checks,
            "roofline": {"bound": "hbm", "kernel": "pipeline", "achieved": value * 32 / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": value * 32 / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                         "algorithmic_bytes_per_atom_frame": 32,
                         "stages": {k: {"ms": stage_ms[k], "GBps": N * F * b / (stage_ms[k] * 1e-3) / 1e9,
                                        "frac": N * F * b / (stage_ms[k] * 1e-3) / 1e9 / peak}
                                    for k, b in (("S1_histogram", 12), ("S3S4_hoptraj", 20))}},
            "e2e": None}
    emit(line)


# -------------------------------------------------------------------------------------------------
# config 4: several species from one coordinate stream (hop_b200.multispecies)
# -------------------------------------------------------------------------------------------------
def run_multispecies_workload(args, n_species=3):
    import numpy as np
    import torch

    from hop_b200 import _lib
    from hop_b200 import kernels as K
    from hop_b200.multispecies import MultiSpeciesPipeline
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--workload config4 is a single-GPU workload (species are replicas: one box per GPU)")
    torch.cuda.set_device(0)
    F, N = FRAMES_PER_GPU, ATOMS
    params = SynthParams(n_atoms=N, n_frames=F, rho=N / 126.0 ** 3)
    frame0 = generate(params, 0, 1)[0]
    per = N // n_species
    columns = [(k * per, (k + 1) * per if k < n_species - 1 else N) for k in range(n_species)]
    xyz = K.synth_generate(params)
    cfg = PipelineConfig(event_capacity=max(1 << 22, F * per // 16))
    multi = MultiSpeciesPipeline.from_frame0(frame0, columns, F, delta=DELTA, config=cfg)
    for _ in range(args.warmup):
        res = multi.run(xyz)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = _lib.lib().hopb_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        res = multi.run(xyz)
    b.record()
    torch.cuda.synchronize()
    ms_step = a.elapsed_time(b) / args.steps
    launches = (_lib.lib().hopb_launch_count() - launches0) // max(args.steps, 1)
    clocks = sampler.stop()
    value = N * F / (ms_step * 1e-3)
    # the way hop does it: one run per species, each reading its own coordinates (parity + the cost of K reads)
    checks, t_single = {}, 0.0
    for k, (c0, c1) in enumerate(columns):
        sub = xyz[:, c0:c1].contiguous()
        pipe = HopPipeline(grid_edges_from_frame0(frame0[c0:c1], delta=DELTA), F, PipelineConfig(async_results=False, event_capacity=cfg.event_capacity))
        pipe.run(sub)
        torch.cuda.synchronize()
        a.record()
        r = pipe.run(sub)
        b.record()
        torch.cuda.synchronize()
        t_single += a.elapsed_time(b)
        checks["species%d_equals_single_species_run" % k] = bool(
            torch.equal(r.counts, res[k].counts) and torch.equal(r.map16, res[k].map16) and torch.equal(r.hop_x, res[k].hop_x)
            and torch.equal(r.hop_y, res[k].hop_y) and torch.equal(r.events, res[k].events) and r.n_labels == res[k].n_labels)
        checks["species%d_counts_conserved" % k] = bool(int(res[k].counts.to(torch.int64).sum().item()) == (c1 - c0) * F)
        del pipe, r, sub
    peak, peak_src = measured_peak()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 coordinates, u32 counts, f64 density, i16/i32 labels, i16 hoptraj planes", "data": "synthetic",
            "config": {"workload": "%s x %d frames, %.1f A grids (one per species, %s); one coordinate stream of %.1f GB >> L2" % (
                WORKLOADS[WORKLOAD]["label"], F, DELTA, " / ".join("x".join(str(n) for n in p.grid.shape) for p in multi.pipes),
                N * F * 12 / 1e9)},
            "step": "hop_b200.multispecies.MultiSpeciesPipeline.run: one histogram pass for all species, one site map and one "
                    "hoptraj launch per species on the shared brick indices",
            "result": {"sites": [int(r.n_labels) for r in res], "events": [int(r.n_events_local) for r in res]},
            "one_run_per_species_ms": t_single, "clocks": clocks, "gpu_launches": int(launches), "checks": checks, "e2e": None,
            "roofline": {"bound": "hbm", "kernel": "pipeline", "achieved": value * 32 / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": value * 32 / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                         "algorithmic_bytes_per_atom_frame": 32}}
    emit(line)


# -------------------------------------------------------------------------------------------------
# config 5: a batch of independent trajectories, 8 per GPU (hop-generate-many-densities style); replicas only
# -------------------------------------------------------------------------------------------------
def run_batch_workload(args, per_gpu=8):
    import numpy as np
    import torch
    import torch.distributed as dist

    from hop_b200 import _lib
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, PipelinePool, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))   # barrier and max-over-ranks only
    F, N = FRAMES_PER_GPU, ATOMS
    seeds = [20261019 + 5 + 100 * (rank * per_gpu + k) for k in range(per_gpu)]
    trajs = [K.synth_generate(SynthParams(n_atoms=N, n_frames=F, seed=sd)) for sd in seeds]
    # every trajectory spans its own grid in frame 0 (hop/density.py:271-278): same shape here, edges a few ulps apart
    all_edges = [grid_edges_from_frame0(generate(SynthParams(n_atoms=N, n_frames=2, seed=sd), 0, 1)[0], delta=DELTA) for sd in seeds]
    edges = all_edges[0]
    cfg = PipelineConfig(event_capacity=max(1 << 22, F * N // 32))
    pool = PipelinePool(edges, F, cfg, n=2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_batch(collect=None):
        pending = []
        for k, xyz in enumerate(trajs):
            r = pool.run(xyz, edges=all_edges[k])
            total = None
            if collect is not None:
                # the count grid is reused by the next run of the same pipeline: reduce it on that pipeline's stream now
                with torch.cuda.stream(pool.streams[k % len(pool.streams)]):
                    total = r.counts.sum(dtype=torch.int64)
            pending.append((k, r, total))
            if len(pending) > 2:                      # two runs in flight; look at a result before its buffers are reused
                j, rj, tj = pending.pop(0)
                if collect is not None:
                    collect[j] = (int(rj.n_labels), int(rj.n_events_local), tj)
                else:
                    rj.wait()
        for j, rj, tj in pending:
            if collect is not None:
                collect[j] = (int(rj.n_labels), int(rj.n_events_local), tj)
            else:
                rj.wait()
        if collect is not None:
            pool.synchronize()
            for j in collect:
                collect[j] = collect[j][:2] + (int(collect[j][2].item()),)

    for _ in range(args.warmup):
        one_batch()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib().hopb_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        one_batch()
    for st in pool.streams:
        torch.cuda.current_stream().wait_stream(st)
    b.record()
    torch.cuda.synchronize()
    pool.synchronize()
    ms_local = a.elapsed_time(b) / args.steps
    barrier()
    launches = (_lib.lib().hopb_launch_count() - launches0) // max(args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    value = world * per_gpu * N * F / (ms_step * 1e-3)
    # checks (not timed): conservation for every trajectory; trajectory 0 equals a stand-alone pipeline run
    got = {}
    one_batch(collect=got)
    checks = {"sum_counts_equals_atom_frames": all(v[2] == N * F for v in got.values())}
    single = HopPipeline(edges, F, PipelineConfig(async_results=False, event_capacity=cfg.event_capacity))
    r0 = single.run(trajs[0])
    checks["trajectory0_equals_single_pipeline"] = bool(got[0][0] == r0.n_labels and got[0][1] == r0.n_events_local)
    rp = pool.run(trajs[0], edges=all_edges[0])
    checks["trajectory0_hoptraj_equal"] = bool(torch.equal(rp.hop_x, r0.hop_x) and torch.equal(rp.hop_y, r0.hop_y)
                                               and torch.equal(rp.events, r0.events))
    ok = torch.tensor([0 if all(checks.values()) else 1], device="cuda")
    if world > 1:
        dist.all_reduce(ok)
    checks["all_ranks"] = bool(int(ok.item()) == 0)
    if rank == 0:
        peak, peak_src = measured_peak()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 coordinates, u32 counts, f64 density, i16/i32 labels, i16 hoptraj planes", "data": "synthetic",
                "config": {"workload": "%s: %d trajectories x %d waters x %d frames per GPU x %d GPU(s) = %d trajectories, %.2f A grid "
                                       "(402^3, 260 MB of counts per trajectory > L2: slabbed histogram); inputs 7.2 GB per "
                                       "trajectory >> L2; replicas only, no collective" % (
                                           WORKLOADS[WORKLOAD]["label"], per_gpu, N, F, world, world * per_gpu, DELTA)},
                "step": "one batch = %d independent trajectories per GPU through PipelinePool (two in flight)" % per_gpu,
                "result": {"sites_by_trajectory": [v[0] for _, v in sorted(got.items())],
                           "events_by_trajectory": [v[1] for _, v in sorted(got.items())]},
                "clocks": clocks, "gpu_launches": int(launches), "checks": checks, "e2e": None,
                "ms_per_trajectory": ms_step / per_gpu,
                "roofline": {"bound": "hbm", "kernel": "pipeline", "achieved": value * 32 / 1e9 / world, "peak": peak, "unit": "GB/s",
                             "frac": value * 32 / 1e9 / world / peak, "peak_source": peak_src, "traffic": None,
                             "algorithmic_bytes_per_atom_frame": 32}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    # same workload (profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum)
    traffic = {}
    try:
        if WORKLOAD == "config2":
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
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
    ev_t = res.events.to(torch.int64)      # [n][6] = frame, iatom, s0, s, tau, tbarrier; reference order = (edge, frame, atom)
    if ev_t.shape[0] > 1:
        ekey = (ev_t[:, 2] + 1) * (res.n_labels + 1) + ev_t[:, 3] + 1
        tkey = ev_t[:, 0] * ATOMS + ev_t[:, 1]
        ok_sorted = bool(((ekey[1:] > ekey[:-1]) | ((ekey[1:] == ekey[:-1]) & (tkey[1:] > tkey[:-1]))).all().item())
    else:
        ok_sorted = True
    checks["events_sorted"] = bool(max_over_ranks_sum(0 if ok_sorted else 1, world) == 0)
    del ev_t
    # BASELINE config 1 as hop ITSELF computed it, frame-sharded over the ranks of this run (not timed)
    checks["reference_config1_sharded"] = golden_parity(world, rank)

    # ---- end to end: pinned host coordinates in, host hoptraj + events out ---------------------------
    if args.no_e2e:
        if rank == 0:
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32 coordinates, u32 counts, f64 density, i16/i32 labels, i16 hoptraj planes",
                    "data": "synthetic", "config": {"workload": workload_name(n_gpus)},
                    "result": {"sites": int(res.n_labels), "events_per_step_rank0": int(res.n_events_local)},
                    "clocks": clocks, "two_in_flight": two_in_flight, "e2e": None, "gpu_launches": int(launches),
                    "roofline": roofline, "stage_ms": stage_ms, "checks": checks}
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return
    xyz_host = torch.empty((FRAMES_PER_GPU, ATOMS, 3), dtype=torch.float32).pin_memory()
    xyz_host.copy_(xyz)
    # the hoptraj planes come back as int16 (lossless; HostRun.planes_float32 restores the DCD's float32 on the host)
    outs = [(torch.empty((FRAMES_PER_GPU, ATOMS), dtype=torch.int16).pin_memory(),
             torch.empty((FRAMES_PER_GPU, ATOMS), dtype=torch.int16).pin_memory()) for _ in range(2)]
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
    hx, hy = res.hop_x.cpu().to(torch.int16), res.hop_y.cpu().to(torch.int16)   # no-ops with the default int16 planes
    checks["e2e_matches_resident"] = bool(all(torch.equal(ox, hx) and torch.equal(oy, hy) for ox, oy in outs[:min(2, e2e_steps)]))
    del hx, hy
    d2h = int(out_x.numel() * out_x.element_size() + out_y.numel() * out_y.element_size() + host["events"].numel() * 4 + host["edge_count"].numel() * 16
              + host["final_state"].numel() * 4)
    e2e = {"value": atom_frames_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(xyz_host.numel() * 4), "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "host_binding": numa,
           "api": ("HopPipeline.run_host_begin/finish (pinned host float32 [F][N][3] in; host hop_x/hop_y as int16 label planes, "
                   "sorted events, COO counts out; download of step k overlaps upload of step k+1)")}

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu_baseline = cpu_baseline_leg(res.density.cpu().numpy(), res.map16.cpu().numpy(), edges,
                                        pipe.cfg.solvent_threshold, pipe.cfg.bulk_threshold)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 coordinates, u32 counts, f64 density, i16/i32 labels, i16 hoptraj planes",
                "data": "synthetic", "config": {"workload": workload_name(n_gpus)},
                "result": {"sites": int(res.n_labels), "events_per_step_rank0": int(res.n_events_local)},
                "clocks": clocks, "two_in_flight": two_in_flight, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "stage_ms": stage_ms, "checks": checks}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def golden_parity(world, rank):
    """BASELINE.json configs[0] (1 000 waters x 200 frames, 1.0 A grid) against what hop ITSELF produced for
    it (tests/golden/reference/ref_config1.npz, frozen from the reference's own classes by
    tests/golden/make_reference_golden.py), with the 200 frames sharded over the ranks of this run exactly
    like the timed workload: counts all-reduce, slab-summary all-gather, count-matrix all-reduce.  Run
    twice (the second run takes the asynchronous finalisation).  Returns {check: bool}, AND-ed over ranks."""
    import numpy as np
    import torch

    from hop_b200 import exchange
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from tests.refgolden import load, reference_products

    g = load("ref_config1")
    ref = reference_products(g)
    xyz, dt = g["xyz"], float(g["dt"])
    F, N = xyz.shape[:2]
    edges = grid_edges_from_frame0(xyz[0], delta=float(g["delta"]))
    f0, f1 = exchange.slab_bounds(F, world, rank)
    pipe = HopPipeline(edges, F, PipelineConfig(n_chunks=3))
    slab = torch.from_numpy(np.ascontiguousarray(xyz[f0:f1])).cuda()
    pipe.run(slab, f0)
    res = pipe.run(slab, f0)
    out = {"edges": all(np.array_equal(a, b) for a, b in zip(edges, (g["edges_x"], g["edges_y"], g["edges_z"])))}
    out["density"] = np.array_equal(res.density.cpu().numpy().reshape(g["grid"].shape), g["grid"])
    out["map"] = np.array_equal(res.map16.cpu().numpy(), ref["map"])
    out["hop_x"] = np.array_equal(res.hop_x.cpu().numpy(), ref["hop_x"][f0:f1].astype(np.float32))
    out["hop_y"] = np.array_equal(res.hop_y.cpu().numpy(), ref["hop_y"][f0:f1].astype(np.float32))
    ev = K.events_to_numpy(res.events)
    got = [(int(a), int(b), int(fr), int(ia), float(t) * dt, float(tb) * dt)
           for a, b, fr, ia, t, tb in zip(ev["s0"], ev["s"], ev["frame"], ev["iatom"], ev["tau"], ev["tbarrier"])]
    want = sorted((k[0], k[1], int(fr), int(ia), float(t), float(tb)) for k, d in ref["props"].items()
                  for t, tb, fr, ia in zip(d["tau"], d["tbarrier"], d["frame"], d["iatom"]) if f0 <= fr < f1)
    out["events"] = got == want
    M = res.count_matrix.cpu().numpy()
    out["count_matrix"] = bool(M.sum() == sum(len(d["tau"]) for d in ref["props"].values())
                               and all(M[a + 1, b + 1] == len(d["tau"]) for (a, b), d in ref["props"].items()))
    st = res.final_state.cpu().numpy()
    nodes = sorted((int(st[0][i]), int(i), float((F - 1) - st[1][i]) * dt) for i in np.nonzero(st[0] != 0)[0])
    want_nodes = sorted((int(s0), int(ia), float(t)) for s0, d in ref["nprops"].items() for t, ia in zip(d["tau"], d["iatom"]))
    out["nodes"] = nodes == want_nodes
    return {k: bool(max_over_ranks_sum(0 if v else 1, world) == 0) for k, v in out.items()}


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
                         "(100 000 waters x 12 500 frames per GPU: the full 100 000 frames at --gpus 8); config3full = "
                         "configs[2] in full on one GPU, chunk-streamed")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (15 GB of pinned memory per rank at config3)")
    args = ap.parse_args()
    global ATOMS, FRAMES_PER_GPU, DELTA, WORKLOAD
    WORKLOAD = args.workload
    ATOMS, FRAMES_PER_GPU, DELTA = (WORKLOADS[WORKLOAD][k] for k in ("atoms", "frames_per_gpu", "delta"))
    _stdout_to_stderr()
    if args.impl == "reference":
        run_reference(args)
    elif WORKLOAD == "config3full":
        run_streamed_workload(args)
    elif WORKLOAD == "config5":
        run_batch_workload(args)
    elif WORKLOAD == "config4":
        run_multispecies_workload(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
