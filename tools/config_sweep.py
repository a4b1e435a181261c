"""BASELINE.json's five configurations through HopPipeline on one GPU: invariants, an oracle
comparison on a frame prefix, and the time of one pass (inputs resident in HBM).

    python tools/config_sweep.py [--configs 1,2,3,4,5] > gpurun_out/config_sweep.jsonl

Configs that do not fit one GPU are run as the slab one of the eight GPUs would own (config 3:
100 000 frames / 8) or as one of the independent units (config 4: one ion species, config 5: one of
the 64 trajectories).  Checks per config:
  * every atom-frame is counted exactly once (the synthetic atoms stay inside the padded grid),
  * site labels in range, hydration sites == thresholded density, count matrix sums to the events,
  * the first PREFIX frames agree with the oracle bit for bit: histogram, hoptraj x / y (with the
    full-size site map) and the events of those frames,
  * the hopping trajectory and the events do not depend on the chunking or on the lookup mode.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

PREFIX = 16

CONFIGS = {
    1: dict(name="config 1: 1 000 waters x 200 frames, 1.0 A", atoms=1000, frames=200, delta=1.0, rho=None),
    2: dict(name="config 2: 30 000 waters x 10 000 frames, 0.5 A", atoms=30000, frames=10000, delta=0.5, rho=None),
    3: dict(name="config 3: 100 000 waters x 12 500 frames (1/8 slab of 100 000), 0.5 A", atoms=100000, frames=12500,
            delta=0.5, rho=None),
    4: dict(name="config 4: 1 000 ions of one species in a 126 A box x 1 000 000 frames, 1.0 A", atoms=1000,
            frames=1000000, delta=1.0, rho=1000 / 126.0 ** 3),
    5: dict(name="config 5: one of 64 trajectories, 30 000 waters x 20 000 frames, 0.25 A", atoms=30000, frames=20000,
            delta=0.25, rho=None),
}


def run_config(cid):
    from oracle import hop_oracle as O
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    c = CONFIGS[cid]
    N, F = c["atoms"], c["frames"]
    kw = {} if c["rho"] is None else {"rho": c["rho"]}
    p = SynthParams(n_atoms=N, n_frames=F, seed=20261019 + cid, **kw)
    xyz = K.synth_generate(p)
    prefix = generate(p, 0, PREFIX)
    checks = {"device_generator_equals_host": bool(np.array_equal(xyz[:PREFIX].cpu().numpy(), prefix))}
    edges = grid_edges_from_frame0(prefix[0], delta=c["delta"])
    pipe = HopPipeline(edges, F, PipelineConfig())
    res = pipe.run(xyz)
    torch.cuda.synchronize()
    t, stages = [], {}
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        a.record()
        res = pipe.run(xyz, marks=marks)
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            stages.setdefault(n1, []).append(e0.elapsed_time(e1))
    ms = float(np.median(t))
    stages = {k: float(np.median(v)) for k, v in stages.items()}

    checks["counts_conserved"] = bool(int(res.counts.to(torch.int64).sum().item()) == N * F)
    m = res.map16
    checks["labels_in_range"] = bool(int(m.min().item()) == 0 and int(m.max().item()) == res.n_labels - 1
                                     and int(res.hop_x.max().item()) < res.n_labels and int(res.hop_x.min().item()) >= -1)
    checks["sites_are_thresholded_density"] = bool(torch.equal(res.own > 0, res.density >= pipe.cfg.solvent_threshold)
                                                   and torch.equal(m >= 2, res.own > 0))
    checks["count_matrix_sums_to_events"] = bool(res.count_matrix is None or
                                                 int(res.count_matrix.sum().item()) == res.n_events_local)
    checks["edge_counts_sum_to_events"] = bool(int(res.edge_count.sum().item()) == res.n_events_local)

    # prefix against the oracle
    site_map = m.cpu().numpy()
    hop = O.coord2hop_vectorised(prefix, edges, site_map)
    checks["prefix_hoptraj_x"] = bool(np.array_equal(res.hop_x[:PREFIX].cpu().numpy(), hop[:, :, 0]))
    checks["prefix_hoptraj_y"] = bool(np.array_equal(res.hop_y[:PREFIX].cpu().numpy(), hop[:, :, 1]))
    ev_ref, _ = O.analyze_hops_vectorised(hop[:, :, 0])
    ev = K.events_to_numpy(res.events)
    sel = ev[ev["frame"] < PREFIX]
    order = np.lexsort((sel["iatom"], sel["frame"]))
    got = np.stack([sel[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1).astype(np.int64)
    checks["prefix_events"] = bool(np.array_equal(got, ev_ref))
    g = pipe.grid
    cpre = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, xyz[:PREFIX], cpre)
    checks["prefix_histogram"] = bool(np.array_equal(cpre.cpu().numpy().astype(np.int64), O.histogram_vectorised(prefix, edges)))
    del cpre

    # independence of chunking / lookup mode (coordinates instead of cached indices, no block map)
    ref_ev = ev
    nc0, cl0 = K.plan_chunks(F, N, pipe.num_sms)
    n_chunks = max(1, min(7, F // 32)) if nc0 != 7 else 5
    chunk_len = -(-F // n_chunks)
    n_chunks = -(-F // chunk_len)
    evb = K.EventBuffer(max(len(ref_ev) + 16, 1 << 16), "cuda")
    X, Y = torch.empty_like(res.hop_x), torch.empty_like(res.hop_y)
    s = K.hoptraj(g, xyz, pipe.brick, 0, n_chunks, chunk_len, evb, X, Y, block_map="none")
    st = K.new_state(N, "cuda")
    K.hop_advance(None, st, init=True)
    K.hop_fixup(s, chunk_len, F, st, Y, evb)
    same = bool(torch.equal(X, res.hop_x) and torch.equal(Y, res.hop_y) and evb.n() == len(ref_ev))
    if same and len(ref_ev):
        out = K.events_finalize(evb, evb.n(), N, F, res.n_labels)
        same = bool(np.array_equal(K.events_to_numpy(out[0]), ref_ev))
    checks["independent_of_chunking_and_lookup"] = same
    del X, Y, evb, s

    line = {"config": c["name"], "atoms": N, "frames": F, "grid": list(g.shape), "sites": int(res.n_labels),
            "events": int(res.n_events_local), "chunks": [nc0, cl0], "ms_per_pass": ms, "stage_ms": stages,
            "atom_frames_per_s": N * F / (ms * 1e-3), "checks": checks, "ok": bool(all(checks.values()))}
    del res, pipe, xyz
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    args = ap.parse_args()
    ok = True
    for cid in [int(x) for x in args.configs.split(",")]:
        t0 = time.time()
        line = run_config(cid)
        line["wall_s"] = time.time() - t0
        ok &= line["ok"]
        print(json.dumps(line), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
