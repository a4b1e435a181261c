"""Three whole-trajectory runs of HopPipeline on the synthetic box (atoms, frames, grid spacing from the command line): the
program the ncu captures under profiles/ and the HOPB_FUSED_TRACE phase tables were taken from.

    python tools/step_loop.py 30000 10000 0.5        # the bench workload (BASELINE config 2)
    HOPB_FUSED_TRACE=1 python tools/step_loop.py 30000 20000 0.25   # 402^3 grid (config 5), stage-2 phase times"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from hop_b200 import kernels as K
from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
from hop_b200.synth import SynthParams, generate
ATOMS, FR, DELTA = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
p = SynthParams(n_atoms=ATOMS, n_frames=FR)
edges = grid_edges_from_frame0(generate(p, 0, 1)[0], delta=DELTA)
xyz = K.synth_generate(p)
pipe = HopPipeline(edges, FR, PipelineConfig())
for i in range(3):
    res = pipe.run(xyz)
    torch.cuda.synchronize()
print("sites", res.n_labels)
