"""The drop-in classes end to end, the way hop's workflow calls them (hop/interactive.py:235-307,
417-464): DensityCreator -> Density.map_sites / site_insert_bulk -> HoppingTrajectory.write ->
TransportNetwork.HoppingGraph, on an in-memory universe (host coordinates) and with the hopping
trajectory written to a DCD file.  Wall-clock per stage, atom-frames/s for the whole workflow.

    python tools/class_api_bench.py [--atoms 30000] [--frames 2000] [--delta 0.5] > gpurun_out/class_api.json
"""
import argparse
import json
import math
import os
import sys
import tempfile
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=30000)
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--delta", type=float, default=0.5)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    from hop_b200 import kernels as K
    from hop_b200.density import DensityCreator
    from hop_b200.graph import TransportNetwork
    from hop_b200.synth import SynthParams
    from hop_b200.trajectory import HoppingTrajectory
    from hop_b200.universe import Universe

    p = SynthParams(n_atoms=args.atoms, n_frames=args.frames)
    xyz = K.synth_generate(p).cpu().numpy()                  # host coordinates, like a trajectory reader yields
    u = Universe(coordinates=xyz, names=["OH2"] * args.atoms, dt=1.0, filename="synthetic.psf")
    out = None
    for rep in range(args.repeat):                            # the first pass pays allocations and warm-up
        t = {}
        with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            DC = DensityCreator(u, mode="solvent", delta=args.delta, atomselection="name OH2")
            dens = DC.create()["solvent"]
            dens.convert_density("water")
            torch.cuda.synchronize()
            t["density"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            dens.map_sites(math.e)
            bulk = DensityCreator(u, mode="solvent", delta=args.delta, atomselection="name OH2").create()["solvent"]
            bulk.convert_density("water")
            bulk.map_sites(math.exp(-0.5))
            dens.site_insert_bulk(bulk)
            n_sites = len(dens.site_labels("sites"))
            torch.cuda.synchronize()
            t["site_map(+bulk density)"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            h = HoppingTrajectory(u.trajectory, u.select_atoms("name OH2"), dens)
            h.write(os.path.join(tmp, "hoptraj"))
            torch.cuda.synchronize()
            t["hoptraj.write (dcd+psf)"] = time.perf_counter() - t0
            size = os.path.getsize(os.path.join(tmp, "hoptraj.dcd"))
            t0 = time.perf_counter()
            tn = TransportNetwork(h, dens)
            tn.compute_residency_times()
            torch.cuda.synchronize()
            t["transport network + rates"] = time.perf_counter() - t0
            n_edges = tn.graph.number_of_edges()
        total = sum(t.values())
        out = {"atoms": args.atoms, "frames": args.frames, "delta": args.delta, "pass": rep, "seconds": t, "total_s": total,
               "atom_frames_per_s": args.atoms * args.frames / total, "sites": int(n_sites), "edges": int(n_edges),
               "hoptraj_dcd_bytes": int(size)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
