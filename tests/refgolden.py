"""Helpers for the golden vectors produced by the reference itself (tests/golden/reference/*.npz,
see tests/golden/make_reference_golden.py): loading, and the canonical numbering of equal-volume
sites.  Used by tests/test_reference_golden.py and __graft_entry__.smoke()."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "reference", "*.npz")))
NAMES = [os.path.splitext(os.path.basename(f))[0] for f in FILES]


def load(name):
    g = dict(np.load(os.path.join(HERE, "golden", "reference", name + ".npz"), allow_pickle=False))
    if g["xyz"].shape[0] == 0:          # coordinates of the bench generator: stored as seed + hash
        import hashlib

        from hop_b200.synth import SynthParams, generate

        g["xyz"] = generate(SynthParams(n_atoms=int(g["n_solvent"]), n_frames=g["hop_x"].shape[0],
                                        seed=int(g["synth_seed"])))
        assert hashlib.sha256(np.ascontiguousarray(g["xyz"]).tobytes()).hexdigest() == str(g["xyz_sha256"])
    if "rates" not in g:
        g["rates"] = np.zeros((0, 4))
    return g


def canonical_perm(site_map):
    """perm[label] -> canonical label.  Labels <= 1 (outlier, interstitial, bulk) are fixed; hydration
    sites are ordered by volume descending, then larger minimum C-order cell first."""
    m = np.asarray(site_map).ravel().astype(np.int64)
    S = int(m.max())
    vol = np.bincount(m[m > 0], minlength=S + 1)
    first = np.full(S + 1, -1, np.int64)
    idx = np.nonzero(m > 0)[0]
    order = np.argsort(m[idx], kind="stable")
    ls, cs = m[idx][order], idx[order]
    starts = np.r_[0, np.nonzero(np.diff(ls))[0] + 1]
    first[ls[starts]] = cs[starts]
    perm = np.arange(S + 1)
    for new, old in enumerate(sorted(range(2, S + 1), key=lambda l: (-vol[l], -first[l])), start=2):
        perm[old] = new
    return perm


def relabel(a, perm):
    a = np.asarray(a).astype(np.int64)
    out = a.copy()
    pos = a > 0
    out[pos] = perm[a[pos]]
    return out


def reference_products(g):
    """The reference's outputs in canonical site numbering."""
    perm = canonical_perm(g["map"])
    inv = np.argsort(perm)                         # canonical label -> reference label
    ev = g["events"].copy()
    ev[:, 0], ev[:, 1] = relabel(ev[:, 0], perm), relabel(ev[:, 1], perm)
    nodes = g["nodes"].copy()
    nodes[:, 0] = relabel(nodes[:, 0], perm)
    props = {}
    for s0, s, tau, tb, frame, ia in ev:           # emission order inside an edge is the file order
        d = props.setdefault((int(s0), int(s)), dict(tau=[], tbarrier=[], frame=[], iatom=[]))
        d["tau"].append(tau), d["tbarrier"].append(tb), d["frame"].append(int(frame)), d["iatom"].append(int(ia))
    nprops = {}
    for s0, tau, frame, ia in nodes:
        d = nprops.setdefault(int(s0), dict(tau=[], frame=[], iatom=[]))
        d["tau"].append(tau), d["frame"].append(int(frame)), d["iatom"].append(int(ia))
    for d in nprops.values():                      # node records are emitted in atom order
        o = np.argsort(d["iatom"], kind="stable")
        for k in d:
            d[k] = list(np.asarray(d[k])[o])
    allt = g["alltransitions"].copy()
    allt = np.stack([relabel(allt[:, 0], perm), relabel(allt[:, 1], perm)], axis=1)
    rates = g["rates"].copy()
    rates[:, 0], rates[:, 1] = relabel(rates[:, 0], perm), relabel(rates[:, 1], perm)
    return dict(perm=perm, map=relabel(g["map"], perm), hop_x=relabel(g["hop_x"], perm),
                hop_y=relabel(g["hop_y"], perm), props=props, nprops=nprops,
                alltransitions=set(map(tuple, allt.tolist())),
                rates={(int(a), int(b)): (k, int(n)) for a, b, k, n in rates},
                volume=g["site_volume"][inv], occ_avg=g["site_occupancy_avg"][inv],
                occ_std=g["site_occupancy_std"][inv], center=g["site_center"][inv],
                ncells=g["site_ncells"][inv])
