"""Regenerates tests/golden/*.npz from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors (hop/tests/README:1-4),
so these pin the oracle itself: any later change to oracle/ or to the synthetic generator that
alters a result shows up as a diff against the committed files."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import hop_oracle as O  # noqa: E402
from hop_b200.synth import SynthParams, generate  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, n_atoms, n_frames, delta, perturb=False):
    p = SynthParams(n_atoms=n_atoms, n_frames=n_frames)
    xyz = generate(p)
    if perturb:
        rng = np.random.default_rng(7)
        m = rng.random(xyz.shape[:2]) < 0.03
        m[0, :] = False
        xyz[m] += np.float32(50.0)          # outliers
        xyz[0, 5:9] -= np.float32(60.0)     # outliers in the first frame: the (-1, N) quirk
    res = O.pipeline(xyz, delta=delta, faithful=True)
    res2 = O.pipeline(xyz, delta=delta, faithful=False)
    assert np.array_equal(res["map"], res2["map"]) and np.array_equal(res["hop"], res2["hop"])
    ev, nodes = res2["events"], res2["nodes"]
    nz = np.nonzero(res["counts"].ravel())[0]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        n_atoms=n_atoms, n_frames=n_frames, delta=delta, seed=p.seed, perturb=perturb,
        xyz=xyz, edges_x=res["edges"][0], edges_y=res["edges"][1], edges_z=res["edges"][2],
        counts_idx=nz.astype(np.int32), counts_val=res["counts"].ravel()[nz].astype(np.int32),
        grid_sha256=hashlib.sha256(np.ascontiguousarray(res["grid"]).tobytes()).hexdigest(),
        site_map=res["map"], volumes=np.array([len(s) for s in res["sites"]], dtype=np.int32),
        hop_x=res["hop"][:, :, 0].astype(np.int16), hop_y=res["hop"][:, :, 1].astype(np.int16),
        events=ev.astype(np.int32), nodes=nodes.astype(np.int32))
    print(name, "sites", len(res["sites"]) - 1, "events", len(ev), "outlier frames", int((res["hop"][:, :, 0] == -1).sum()))


if __name__ == "__main__":
    case("small_clean", 300, 60, 1.0)
    case("small_outliers", 300, 60, 1.0, perturb=True)
    case("half_angstrom", 400, 40, 0.5)
