"""The oracle against the committed golden vectors, and its two restatements against each other."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import hop_oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    xyz = g["xyz"]
    from hop_b200.synth import SynthParams, generate

    if not bool(g["perturb"]):
        assert np.array_equal(generate(SynthParams(n_atoms=int(g["n_atoms"]), n_frames=int(g["n_frames"]), seed=int(g["seed"]))), xyz)
    res = O.pipeline(xyz, delta=float(g["delta"]), faithful=False)
    for e, k in zip(res["edges"], ("edges_x", "edges_y", "edges_z")):
        assert np.array_equal(e, g[k])
    counts = np.zeros(res["counts"].size, np.int32)
    counts[g["counts_idx"]] = g["counts_val"]
    assert np.array_equal(res["counts"].ravel(), counts)
    assert hashlib.sha256(np.ascontiguousarray(res["grid"]).tobytes()).hexdigest() == str(g["grid_sha256"])
    assert np.array_equal(res["map"], g["site_map"])
    assert np.array_equal(res["hop"][:, :, 0], g["hop_x"].astype(np.float32))
    assert np.array_equal(res["hop"][:, :, 1], g["hop_y"].astype(np.float32))
    assert np.array_equal(res["events"], g["events"])
    assert np.array_equal(res["nodes"], g["nodes"])


def test_faithful_and_vectorised_agree_on_config1():
    from hop_b200.synth import SynthParams, generate

    xyz = generate(SynthParams(n_atoms=1000, n_frames=200))[:80]
    a = O.pipeline(xyz, delta=1.0, faithful=True)
    b = O.pipeline(xyz, delta=1.0, faithful=False)
    assert np.array_equal(a["counts"], b["counts"]) and np.array_equal(a["grid"], b["grid"])
    assert a["sites"] == b["sites"] and np.array_equal(a["map"], b["map"])
    assert np.array_equal(a["hop"], b["hop"])
    ev = b["events"]
    n = 0
    for key, d in a["props"].items():
        if not isinstance(key, tuple):
            continue
        sel = ev[(ev[:, 2] == key[0]) & (ev[:, 3] == key[1])]
        assert np.array_equal(sel[:, 0], d["frame"]) and np.array_equal(sel[:, 1], d["iatom"])
        assert np.array_equal(sel[:, 4], d["tau"]) and np.array_equal(sel[:, 5], d["tbarrier"])
        n += len(sel)
    assert n == len(ev)
    nodes = {k: d for k, d in a["props"].items() if not isinstance(k, tuple)}
    assert sum(len(d["tau"]) for d in nodes.values()) == len(b["nodes"])


def test_connectivity_is_14_not_26():
    g = np.zeros((3, 3, 3))
    g[0, 1, 0] = g[1, 0, 0] = 1.0          # mixed-sign diagonal: NOT connected in the reference
    assert len(O.label_sites_faithful(g, 0.5)) - 1 == 2
    assert len(O.label_sites_ndimage(g, 0.5)) - 1 == 2
    g2 = np.zeros((3, 3, 3))
    g2[0, 0, 0] = g2[1, 1, 1] = 1.0        # (1,1,1) offset: connected
    assert len(O.label_sites_faithful(g2, 0.5)) - 1 == 1


def test_ndimage_and_networkx_partitions_agree():
    rng = np.random.default_rng(0)
    for shape, p in [((24, 24, 24), 0.2), ((9, 31, 17), 0.4), ((30, 5, 5), 0.1)]:
        g = rng.random(shape)
        for minsite in (1, 2):
            assert O.label_sites_faithful(g, 1 - p, minsite) == O.label_sites_ndimage(g, 1 - p, minsite)


def test_state_machine_documented_quirks():
    # returning to the same site through the interstitial is not a transition (hysteresis)
    x = np.array([[2], [0], [0], [2], [0], [3]], np.float32)
    ev, nodes = O.analyze_hops_vectorised(x)
    assert ev.tolist() == [[5, 0, 2, 3, 4, 1]]           # tau = t1 - t0 = 4 - 0, tbarrier = 5 - 4
    assert nodes.tolist() == [[0, 3, 0]]
    # atom starting as outlier produces a (-1, N) edge (hop/graph.py:217-221)
    x = np.array([[-1], [-1], [0], [4]], np.float32)
    ev, nodes = O.analyze_hops_vectorised(x)
    assert ev.tolist() == [[3, 0, -1, 4, 2, 1]]
    # atom starting in the interstitial: first entry is not a transition
    x = np.array([[0], [5], [6]], np.float32)
    ev, _ = O.analyze_hops_vectorised(x)
    assert ev.tolist() == [[2, 0, 5, 6, 1, 0]]
