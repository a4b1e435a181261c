"""Experiment: decompose the hoptraj kernel time (gathers / event slow path / block map kind)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from hop_b200 import kernels as K
from hop_b200.synth import SynthParams, generate
from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

ATOMS, FRAMES = 30000, 10000
params = SynthParams(n_atoms=ATOMS, n_frames=FRAMES)
edges = grid_edges_from_frame0(generate(params, 0, 1)[0], delta=0.5)
xyz = K.synth_generate(params, frame_start=0, n_out=FRAMES)
pipe = HopPipeline(edges, FRAMES, PipelineConfig())
res = pipe.run(xyz, 0)
n_chunks, chunk_len = K.plan_chunks(FRAMES, ATOMS, pipe.num_sms)
ev = K.EventBuffer(1 << 25, "cuda")


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(bm, block_map="auto", planes=True):
    def f():
        ev.reset()
        K.hoptraj(pipe.grid, xyz, bm, 0, n_chunks, chunk_len, ev, res.hop_x if planes else None, res.hop_y if planes else None,
                  pipe._bufs["summaries"], cells=pipe.cells, block_map=block_map)
    t = timeit(f)
    return "%.3f ms  (%d events)" % (t, ev.n())


m = res.map16
for bmk in ("fine", "coarse", "none"):
    print("real map, %-6s     " % bmk, run(pipe.brick, bmk))
print("real map, no planes  ", run(pipe.brick, "fine", planes=False))
ones = torch.ones_like(m)
print("all-bulk map         ", run(K.brick_map(ones), "fine"))
m01 = torch.where(m > 1, torch.ones_like(m), m)      # hydration sites melted into the bulk: gathers stay, events go
print("0/1 map (no sites)   ", run(K.brick_map(m01), "fine"), "fine")
print("0/1 map (no sites)   ", run(K.brick_map(m01), "coarse"), "coarse")
m1s = torch.where(m == 0, torch.ones_like(m), m)     # interstitial melted into the bulk: events stay
print("bulk + sites map     ", run(K.brick_map(m1s), "fine"))
