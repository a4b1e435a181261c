"""The chunk/slab stitching algebra (csrc/hopb_machine.h) and the float32 binning
(csrc/hopb_binning.h), replayed on the CPU and compared with the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import hop_oracle as O

EV = np.dtype([("frame", "i4"), ("iatom", "i4"), ("s0", "i4"), ("s", "i4"), ("tau", "i4"), ("tbar", "i4")])


def random_labels(rng, F, N, nsites, p_stay, p_out, p_zero):
    lab = np.zeros((F, N), np.int32)
    cur = rng.integers(-1, nsites + 1, size=N)
    for f in range(F):
        u = rng.random(N)
        new = rng.integers(1, nsites + 1, size=N)
        cur = np.where(u < p_stay, cur, new)
        v = rng.random(N)
        lab[f] = np.where(v < p_out, -1, np.where(v < p_out + p_zero, 0, cur))
    return lab


def run_emul(emul, lab, slab_len, chunk_len):
    F, N = lab.shape
    cap = F * N + 16
    ev = np.zeros(cap, EV)
    state = np.zeros((N, 4), np.int32)
    orbit = np.zeros((F, N), np.int32)
    n = emul.emul_scan(lab.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(F), ctypes.c_int64(N),
                       ctypes.c_int64(slab_len), ctypes.c_int64(chunk_len),
                       ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(cap),
                       state.ctypes.data_as(ctypes.c_void_p), orbit.ctypes.data_as(ctypes.c_void_p))
    assert n >= 0, "emul_scan failed (%d)" % n
    ev = ev[:n]
    order = np.lexsort((ev["iatom"], ev["frame"]))
    return ev[order], state, orbit


def oracle_orbit(lab):
    y = np.zeros_like(lab)
    last = np.zeros(lab.shape[1], lab.dtype)
    for f in range(lab.shape[0]):
        off = (lab[f] == 0) | (lab[f] == -1)
        last = np.where(off, last, lab[f])
        y[f] = last
    return y


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("cut", [(10 ** 9, 10 ** 9), (10 ** 9, 7), (50, 50), (37, 5), (16, 1), (3, 2)])
def test_stitching_matches_oracle(emul, seed, cut):
    rng = np.random.default_rng(seed)
    F, N = 160, 48
    params = [(0.9, 0.05, 0.3), (0.5, 0.2, 0.2), (0.98, 0.0, 0.6), (0.7, 0.5, 0.1), (0.2, 0.01, 0.01), (0.95, 0.1, 0.8)][seed]
    lab = random_labels(rng, F, N, nsites=3 + seed % 3, p_stay=params[0], p_out=params[1], p_zero=params[2])
    slab_len, chunk_len = cut
    ev, state, orbit = run_emul(emul, lab, min(slab_len, F), min(chunk_len, F))
    ref_ev, ref_nodes = O.analyze_hops_vectorised(lab)
    got = np.stack([ev[k] for k in ("frame", "iatom", "s0", "s", "tau", "tbar")], axis=1).astype(np.int64)
    assert got.shape == ref_ev.shape
    assert np.array_equal(got, ref_ev)
    # node records from the final state
    ia = np.nonzero(state[:, 0] != 0)[0]
    nodes = np.stack([ia, state[ia, 0], (F - 1) - state[ia, 1]], axis=1)
    assert np.array_equal(nodes, ref_nodes)
    assert np.array_equal(orbit, oracle_orbit(lab))


def test_stitching_vs_faithful_transcription(emul):
    rng = np.random.default_rng(11)
    lab = random_labels(rng, 60, 20, nsites=4, p_stay=0.8, p_out=0.1, p_zero=0.3)
    ev, state, _ = run_emul(emul, lab, 17, 4)
    props, edges, nodes = O.analyze_hops_faithful(lab.astype(np.float32), dt=1.0)
    for (a, b) in edges:
        sel = ev[(ev["s0"] == a) & (ev["s"] == b)]
        d = props[(a, b)]
        assert np.array_equal(sel["frame"], d["frame"]) and np.array_equal(sel["iatom"], d["iatom"])
        assert np.array_equal(sel["tau"], d["tau"]) and np.array_equal(sel["tbar"], d["tbarrier"])
    assert len(ev) == sum(len(props[e]["tau"]) for e in edges)


def _bins(emul, edges, xyz, mode):
    out = np.zeros((len(xyz), 3), np.int32)
    top = np.zeros(3, np.float32)
    decs = np.array(O.top_edge_decimals(edges), np.int32)
    e = [np.ascontiguousarray(x, np.float64) for x in edges]
    rc = emul.emul_bin(e[0].ctypes.data_as(ctypes.c_void_p), len(e[0]) - 1, e[1].ctypes.data_as(ctypes.c_void_p),
                       len(e[1]) - 1, e[2].ctypes.data_as(ctypes.c_void_p), len(e[2]) - 1,
                       decs.ctypes.data_as(ctypes.c_void_p), xyz.ctypes.data_as(ctypes.c_void_p),
                       ctypes.c_int64(len(xyz)), mode, out.ctypes.data_as(ctypes.c_void_p),
                       top.ctypes.data_as(ctypes.c_void_p))
    assert rc >= 0
    if mode >= 2:
        assert rc == 1, "fast path not proven exact for these edges"
    return out, top


def adversarial_points(rng, edges, n):
    """random points plus float32 neighbours of every edge, values outside, NaN/inf."""
    lo = np.array([e[0] for e in edges])
    hi = np.array([e[-1] for e in edges])
    pts = (lo - 1.0 + rng.random((n, 3)) * (hi - lo + 2.0)).astype(np.float32)
    near = []
    for d in range(3):
        e32 = edges[d].astype(np.float32)
        for k in range(-3, 4):
            v = e32.copy()
            for _ in range(abs(k)):
                v = np.nextafter(v, np.float32(np.inf if k > 0 else -np.inf))
            p = (lo + rng.random((len(v), 3)) * (hi - lo)).astype(np.float32)
            p[:, d] = v
            near.append(p)
    special = np.array([[np.nan, 1, 1], [np.inf, 1, 1], [-np.inf, 1, 1], [1e30, 1, 1], [-1e30, 1, 1]], np.float32)
    return np.ascontiguousarray(np.concatenate([pts] + near + [special]))


@pytest.mark.parametrize("delta,origin", [(1.0, 0.0), (0.5, -1.9999), (0.25, 3.3), (0.1, -7.7), (1.3, 100.0), (25.0, -300.0)])
def test_binning_matches_numpy(emul, delta, origin):
    rng = np.random.default_rng(5)
    coords0 = (origin + rng.random((200, 3)) * np.array([17.3, 9.1, 23.7]) * max(delta, 1.0)).astype(np.float32)
    bins, arange, edges = O.grid_geometry(coords0, delta=delta, padding=2.0)
    xyz = adversarial_points(rng, edges, 20000)
    # the branch-free pair-table search must agree with the loop search everywhere (incl. NaN/inf)
    for slow, fast in ((0, 2), (1, 3)):
        a, _ = _bins(emul, edges, xyz, slow)
        b, _ = _bins(emul, edges, xyz, fast)
        n = np.array([len(e) - 1 for e in edges])
        a_out = (a < 1) | (a > n)
        b_out = (b < 1) | (b > n)
        assert np.array_equal(a_out, b_out)
        assert np.array_equal(a[~a_out], b[~b_out])
    # the table-free "sure" bin (with the table as fallback) is the histogram bin everywhere, it is
    # never taken for a hoptraj bin that differs from the histogram bin, and it covers nearly all points
    a, _ = _bins(emul, edges, xyz, 0)
    b, _ = _bins(emul, edges, xyz, 4)
    n = np.array([len(e) - 1 for e in edges])
    a_out, b_out = (a < 1) | (a > n), (b < 1) | (b > n)
    assert np.array_equal(a_out, b_out) and np.array_equal(a[~a_out], b[~b_out])
    sure, _ = _bins(emul, edges, xyz, 5)
    hopb, _ = _bins(emul, edges, xyz, 1)
    assert np.array_equal(hopb[sure == 1], a[sure == 1])
    inside = (xyz[:20000] > np.array([e[0] for e in edges])).all(axis=1) & (xyz[:20000] < np.array([e[-1] for e in edges])).all(axis=1)
    assert sure[:20000][inside].mean() > 0.95
    # histogram semantics
    got, _ = _bins(emul, edges, xyz, 0)
    for d in range(3):
        x = xyz[:, d].astype(np.float64)
        ref = np.searchsorted(edges[d], x, side="right")
        ref[x == edges[d][-1]] -= 1
        n = len(edges[d]) - 1
        inside = (ref >= 1) & (ref <= n)
        got_inside = (got[:, d] >= 1) & (got[:, d] <= n)
        assert np.array_equal(inside, got_inside)
        assert np.array_equal(ref[inside], got[inside, d])
    # hoptraj semantics (digitize + float32 top-edge rule)
    got, top = _bins(emul, edges, xyz, 1)
    finite = ~np.isnan(xyz).any(axis=1)
    ref = O.digitize_hop(xyz[finite], edges)
    decs = O.top_edge_decimals(edges)
    for d in range(3):
        assert top[d] == np.float32(np.around(edges[d][-1], decs[d]))
        n = len(edges[d]) - 1
        r = ref[d]
        g = got[finite, d]
        r_out = (r < 1) | (r > n)
        g_out = (g < 1) | (g > n)
        assert np.array_equal(r_out, g_out)
        assert np.array_equal(r[~r_out], g[~g_out])


def test_histogram_counts_via_emul_bins(emul):
    from hop_b200.synth import SynthParams, generate

    xyz = generate(SynthParams(n_atoms=500, n_frames=20))
    bins, arange, edges = O.grid_geometry(xyz[0], delta=1.0)
    ref, _ = O.histogram_faithful(iter(xyz), bins, arange)
    b, _ = _bins(emul, edges, np.ascontiguousarray(xyz.reshape(-1, 3)), 0)
    shape = tuple(int(x) for x in bins)
    ok = np.all((b >= 1) & (b <= np.array(shape)), axis=1)
    lin = np.ravel_multi_index((b[ok] - 1).T, shape)
    counts = np.bincount(lin, minlength=int(np.prod(shape))).reshape(shape)
    assert np.array_equal(counts, ref.astype(np.int64))
