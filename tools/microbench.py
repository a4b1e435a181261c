"""Per-stage device timings of the hot path on a synthetic config (CUDA events, current stream).

usage: python tools/microbench.py [--atoms N] [--frames F] [--delta D] [--reps R] [--out file]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from hop_b200 import kernels as K
from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
from hop_b200.synth import SynthParams


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=30000)
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--delta", type=float, default=0.5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0, help="CTAs per SM assumed by the chunk planner")
    ap.add_argument("--out", default="gpurun_out/micro.json")
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    p = SynthParams(n_atoms=args.atoms, n_frames=args.frames)
    t0 = time.time()
    xyz = K.synth_generate(p)
    torch.cuda.synchronize()
    print("generated", tuple(xyz.shape), "in %.2fs" % (time.time() - t0), flush=True)
    edges = grid_edges_from_frame0(xyz[0].cpu().numpy(), delta=args.delta)
    cfg = PipelineConfig(n_chunks=args.chunks or None)
    if args.ctas:
        import functools
        K.plan_chunks = functools.partial(K.plan_chunks, warps_per_sm=8 * args.ctas)
    pipe = HopPipeline(edges, args.frames, cfg)
    AF = args.atoms * args.frames
    rows = {}

    def rec(name, ms, bytes_per_af=None, extra=None):
        r = {"ms": ms}
        if bytes_per_af:
            r["atom_frames_per_s"] = AF / (ms * 1e-3)
            r["GBps"] = AF * bytes_per_af / (ms * 1e-3) / 1e9
            r["frac_hbm"] = r["GBps"] / hbm
        if extra:
            r.update(extra)
        rows[name] = r
        print(name, json.dumps(r), flush=True)

    ms, mn, counts = timed(lambda: pipe.histogram(xyz), args.reps)
    rec("S1_hist", ms, 12, {"min_ms": mn, "grid": list(pipe.grid.shape)})
    ms, mn, sm = timed(lambda: pipe.site_map(counts), args.reps)
    density, own, member, map16, props = sm
    rec("S2_sitemap_total", ms, None, {"min_ms": mn, "cells": pipe.grid.n_cells})
    ms, mn, _ = timed(lambda: K.label_sites(density, cfg.solvent_threshold), args.reps)
    rec("S2_label_solvent", ms, None, {"min_ms": mn})
    ms, mn, _ = timed(lambda: K.label_sites(density, cfg.bulk_threshold), args.reps)
    rec("S2_label_bulk", ms, None, {"min_ms": mn})
    ms, mn, _ = timed(lambda: K.site_properties(pipe.grid, density, own, member, pipe.MAX_LABELS), args.reps)
    rec("S2_properties", ms, None, {"min_ms": mn})
    ms, mn, _ = timed(lambda: K.density_from_counts(pipe.grid, counts, args.frames, 30.0, out=density), args.reps)
    rec("S2_density", ms, None, {"min_ms": mn})
    ms, mn, _ = timed(lambda: K.brick_map(map16), args.reps)
    rec("S2_brick_map", ms, None, {"min_ms": mn})
    pipe.site_map(counts)
    ms, mn, ht = timed(lambda: pipe.hoptraj(xyz, map16, 0), args.reps)
    hop_x, hop_y, events, state, final = ht
    n_labels, n_ev = pipe.sync_scalars(events)
    rec("S3S4_hoptraj", ms, 20, {"min_ms": mn, "n_events": n_ev, "n_labels": n_labels,
                                 "chunks": list(K.plan_chunks(args.frames, args.atoms, pipe.num_sms))})
    ms, mn, _ = timed(lambda: pipe.transitions(events, n_ev, args.atoms, n_labels), args.reps)
    rec("S4_finalize", ms, None, {"min_ms": mn})
    ms, mn, _ = timed(lambda: pipe.run(xyz), args.reps)
    rec("pipeline", ms, 32, {"min_ms": mn})
    hx = hop_x.cpu().numpy() if AF <= 5e8 else None
    if hx is not None:
        rows["label_fractions"] = {"outlier": float((hx == -1).mean()), "interstitial": float((hx == 0).mean()),
                                   "bulk": float((hx == 1).mean()), "sites": float((hx > 1).mean())}
        print(rows["label_fractions"])
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"atoms": args.atoms, "frames": args.frames, "delta": args.delta, "hbm_peak_gbs": hbm, "rows": rows},
              open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
