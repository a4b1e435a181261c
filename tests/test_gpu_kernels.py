"""Parity of every CUDA kernel (through the C ABI) against the CPU oracle.  Bit-exact unless stated."""
import math

import numpy as np
import pytest

from oracle import hop_oracle as O
from hop_b200.synth import SynthParams, generate

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hop_b200 import kernels

    return kernels


@pytest.fixture(scope="module")
def traj():
    p = SynthParams(n_atoms=1000, n_frames=200)
    return p, generate(p)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_synth_device_matches_host(K, traj):
    p, xyz = traj
    got = K.synth_generate(p).cpu().numpy()
    assert got.shape == xyz.shape
    assert np.array_equal(got.view(np.uint32), xyz.view(np.uint32))
    tail = K.synth_generate(p, frame_start=150, n_out=20).cpu().numpy()
    assert np.array_equal(tail, xyz[150:170])


@pytest.mark.parametrize("delta", [1.0, 0.5])
def test_histogram_counts(K, traj, delta):
    p, xyz = traj
    _, _, edges = O.grid_geometry(xyz[0], delta=delta)
    ref = O.histogram_vectorised(xyz, edges)
    g = K.GridTables(edges)
    counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, dev(xyz), counts)
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), ref)
    # accumulate in two unaligned pieces (exercises the scalar lead/tail path)
    counts.zero_()
    flat = dev(xyz.reshape(-1, 3))
    cut = 1000 * 77 + 3
    K.hist_accumulate(g, flat[:cut], counts)
    K.hist_accumulate(g, flat[cut:], counts)
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), ref)


@pytest.mark.parametrize("pack16", ["0", "2"])
def test_histogram_slabbed_passes(K, traj, pack16, monkeypatch):
    """Grids beyond the L2 budget are reduced slab by slab from the cached cell indices -- straight into the u32 grid
    (HOPB_HIST_PACK16=0) or through the 16-bit packed grid first (=2; the default for such grids when the call brings
    at least a point per cell)."""
    monkeypatch.setenv("HOPB_HIST_PACK16", pack16)
    p, xyz = traj
    xyz = xyz[:50].copy()
    _, _, edges = O.grid_geometry(xyz[0], delta=0.5)
    top = [np.float32(e[-1]) for e in edges]
    xyz[3, 10:14, 0] = top[0]                         # on / next to the last edge: the two binning rules differ
    xyz[4, 10:14, 1] = np.nextafter(top[1], np.float32(0))
    xyz[5, 10:14, 2] = np.nextafter(top[2], np.float32(1e9))
    ref = O.histogram_vectorised(xyz, edges)
    g = K.GridTables(edges)
    try:
        for budget in (1 << 12, 1 << 16, 1 << 20):
            K.set_hist_l2_budget(budget)
            counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
            cells = torch.empty(xyz.shape[:2], dtype=torch.int32, device="cuda")
            K.hist_accumulate(g, dev(xyz), counts, cells)
            assert np.array_equal(counts.cpu().numpy().astype(np.int64), ref), budget
            # the cached cells still drive the hopping trajectory correctly
            res = O.pipeline(xyz, delta=0.5)
            m16 = dev(res["map"])
            brick = K.brick_map(m16)
            ev = K.EventBuffer(xyz.shape[0] * xyz.shape[1], "cuda")
            X = torch.empty(xyz.shape[:2], dtype=torch.float32, device="cuda")
            Y = torch.empty(xyz.shape[:2], dtype=torch.float32, device="cuda")
            s = K.hoptraj(g, None, brick, 0, 2, 25, ev, X, Y, cells=cells)
            st = K.new_state(xyz.shape[1], "cuda")
            K.hop_advance(None, st, init=True)
            K.hop_fixup(s, 25, 50, st, Y, ev)
            assert np.array_equal(X.cpu().numpy(), res["hop"][:, :, 0]) and np.array_equal(Y.cpu().numpy(), res["hop"][:, :, 1])
    finally:
        K.set_hist_l2_budget(64 << 20)


def test_packed16_histogram_is_exact_even_when_a_counter_wraps(K, traj, monkeypatch):
    """The 16-bit packed count grid (hist_flush16_kernel): equal to the direct reductions on ordinary input, in one and
    in several x-slab passes, accumulating over calls -- and equal as well when a 16-bit counter wraps inside one call
    (70 000 points in one cell, its neighbour in the same word occupied too): the sum check finds the wrap out, the
    packed grid is discarded and the direct-reduction passes queued behind it run."""
    p, xyz = traj
    _, _, edges = O.grid_geometry(xyz[0], delta=1.0)
    g = K.GridTables(edges)
    mids = [0.5 * (e[:-1] + e[1:]) for e in edges]
    hot = np.array([mids[0][5], mids[1][6], mids[2][8]], np.float32)          # cell (5, 6, 8) and its z neighbour (5, 6, 9)
    hot2 = np.array([mids[0][5], mids[1][6], mids[2][9]], np.float32)
    crowd = xyz[:80].copy()                                                       # 80 frames x 1000 atoms
    crowd[:70, :, :] = hot                                                        # 70 000 points in one cell
    crowd[70:75, :500, :] = hot2
    cases = {"ordinary": xyz, "wrap": crowd}
    out = {}
    try:
        for mode in ("0", "2"):
            monkeypatch.setenv("HOPB_HIST_PACK16", mode)
            for name, data in cases.items():
                for budget in (64 << 20, 1 << 14):
                    K.set_hist_l2_budget(budget)
                    counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
                    cells = torch.empty(data.shape[:2], dtype=torch.int32, device="cuda")
                    K.hist_accumulate(g, dev(data), counts, cells)
                    K.hist_accumulate(g, dev(data[:7]), counts, cells[:7])        # a second call adds to the same grid
                    out[(mode, name, budget)] = (counts.cpu().numpy().copy(), cells.cpu().numpy().copy())
    finally:
        K.set_hist_l2_budget(64 << 20)
    for name, data in cases.items():
        ref = O.histogram_vectorised(data, edges) + O.histogram_vectorised(data[:7], edges)
        for budget in (64 << 20, 1 << 14):
            a, b = out[("0", name, budget)], out[("2", name, budget)]
            assert np.array_equal(a[0].astype(np.int64), ref), (name, budget)
            assert np.array_equal(b[0], a[0]) and np.array_equal(b[1], a[1]), (name, budget)
    assert out[("2", "wrap", 64 << 20)][0][5, 6, 8] >= 70000 + 7000


@pytest.mark.parametrize("delta", [1.0, 0.5])
def test_partitioned_histogram_equals_direct_reductions(K, traj, delta, monkeypatch):
    """The spatially partitioned, shared-memory-privatised histogram (hist_part_kernel + hist_tiles_kernel, taken for
    large resident runs) gives the counts and brick indices of the one-reduction-per-point kernel: ordinary points
    through the buckets, last-edge points, unaligned heads / tails and points a full bucket has no room for through
    direct reductions."""
    p, xyz = traj
    _, _, edges = O.grid_geometry(xyz[0], delta=delta)
    ref = O.histogram_vectorised(xyz, edges)
    g = K.GridTables(edges)
    flat = dev(xyz.reshape(-1, 3))

    def run(points, mode):
        monkeypatch.setenv("HOPB_HIST_PART", mode)
        counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
        cells = torch.full((points.shape[0],), -5, dtype=torch.int32, device="cuda")
        K.hist_accumulate(g, points, counts, cells)
        return counts, cells

    direct = run(flat, "0")
    assert np.array_equal(direct[0].cpu().numpy().astype(np.int64), ref)
    for pts in (flat, flat[1:], flat[3:100003], flat[:77]):
        a, b = run(pts, "0"), run(pts, "2")
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # every point in the same few cells: far more than a bucket's capacity (twice the mean) -> the overflow path
    lump = flat[:4096].repeat(60, 1).contiguous()
    a, b = run(lump, "0"), run(lump, "2")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and int(a[0].sum().item()) == lump.shape[0]
    # outliers, NaN, points on the last edges
    odd = flat[:5000].clone()
    odd[::7] += 1000.0
    odd[3::11] = float("nan")
    odd[5::13, 0] = float(edges[0][-1])
    odd[6::17, 2] = float(edges[2][-1])
    a, b = run(odd, "0"), run(odd, "2")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # accumulation over several calls adds up
    monkeypatch.setenv("HOPB_HIST_PART", "2")
    counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    cells = torch.empty((flat.shape[0],), dtype=torch.int32, device="cuda")
    cut = 1000 * 77 + 3
    K.hist_accumulate(g, flat[:cut], counts, cells[:cut])
    K.hist_accumulate(g, flat[cut:], counts, cells[cut:])
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), ref) and torch.equal(cells, direct[1])


def test_histogram_edge_semantics(K):
    rng = np.random.default_rng(3)
    coords0 = (rng.random((300, 3)) * [17.3, 9.1, 23.7] - 1.9999).astype(np.float32)
    _, _, edges = O.grid_geometry(coords0, delta=0.5)
    from tests.test_machine_emul import adversarial_points

    pts = adversarial_points(rng, edges, 50000)
    pts = pts[~np.isnan(pts).any(axis=1)]
    # include points exactly on the last edge when it is representable
    ref = O.histogram_vectorised(pts[None], edges)
    g = K.GridTables(edges)
    counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, dev(pts), counts)
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), ref)


@pytest.mark.parametrize("unit", ["Angstrom^{-3}", "water", "TIP3P"])
def test_density_normalisation_bit_exact(K, traj, unit):
    p, xyz = traj
    _, _, edges = O.grid_geometry(xyz[0], delta=1.0)
    counts = O.histogram_vectorised(xyz, edges)
    ref = O.make_density(counts, edges, xyz.shape[0], unit)
    g = K.GridTables(edges)
    factor = O.density_conversion_factor("Angstrom^{-3}", unit)
    got = K.density_from_counts(g, dev(counts.astype(np.int32)), xyz.shape[0], factor).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape,lo,hi", [((40, 37, 45), 0.0, 20.3), ((64, 64, 64), -7.25, 24.75), ((9, 5, 130), 3.1, 68.2)])
@pytest.mark.parametrize("unit", ["Angstrom^{-3}", "water", "Molar"])
def test_density_table_path_bit_exact(K, shape, lo, hi, unit):
    """density_lookup_kernel (values from the per-width-class table) against the reference's arithmetic,
    with counts on both sides of the table limit (1024), zero counts and the largest uint32 counts."""
    rng = np.random.default_rng(shape[0])
    edges = [np.linspace(lo, hi + 0.37 * d, n + 1) for d, n in enumerate(shape)]
    counts = rng.poisson(37, size=shape).astype(np.int64)
    big = rng.random(shape) < 0.02
    counts[big] = rng.integers(1000, 1050, size=int(big.sum()))
    counts[rng.random(shape) < 0.01] = 0
    counts[0, 0, :3] = [1023, 1024, 2**32 - 1]
    n_frames = 977
    ref = O.make_density(counts.astype(np.float64), edges, n_frames, unit)
    g = K.GridTables(edges)
    factor = O.density_conversion_factor("Angstrom^{-3}", unit)
    got = K.density_from_counts(g, dev(counts.astype(np.uint32).view(np.int32)), n_frames, factor).cpu().numpy()
    assert np.array_equal(got, ref)


def _check_labels(K, grid, thr, minsite=1):
    sites, ref_map = O.map_sites(grid, thr, minsite)
    labels, n_sites, volumes = K.label_sites(dev(grid), thr, minsite)
    assert n_sites == len(sites) - 1
    assert np.array_equal(labels.cpu().numpy(), ref_map.astype(np.int32))
    assert np.array_equal(volumes.cpu().numpy(), np.array([len(s) for s in sites]))
    return sites, labels


@pytest.mark.parametrize("p_fill", [0.02, 0.1, 0.2, 0.35, 0.6, 0.95])
@pytest.mark.parametrize("shape", [(24, 24, 24), (7, 33, 65), (40, 3, 70), (1, 1, 50), (5, 1, 1)])
def test_label_sites_random_masks(K, p_fill, shape):
    rng = np.random.default_rng(int(p_fill * 100) + shape[0])
    grid = rng.random(shape)
    thr = 1.0 - p_fill
    if (grid >= thr).sum() > 20000:
        pytest.skip("would exceed the site limit only through isolated cells")
    try:
        _check_labels(K, grid, thr)
    except OverflowError:
        assert len(O.label_sites_ndimage(grid, thr)) - 1 > 32767


def test_label_sites_against_networkx_restatement(K):
    rng = np.random.default_rng(9)
    grid = rng.random((14, 15, 16))
    sites_f = O.label_sites_faithful(grid, 0.75)
    sites_n, labels = _check_labels(K, grid, 0.75)
    assert sites_f == sites_n


def test_label_sites_minsite2_and_ties(K):
    rng = np.random.default_rng(4)
    grid = rng.random((20, 20, 20))
    _check_labels(K, grid, 0.85, minsite=2)
    # many equal-volume sites: isolated pairs
    g2 = np.zeros((16, 16, 16))
    g2[::4, ::4, 2:4] = 1.0
    g2[1, 1, 8:13] = 1.0
    _check_labels(K, g2, 0.5)


def test_label_sites_solid_blob_and_threshold_equality(K):
    g = np.ones((33, 34, 35)) * 0.6
    g[10:12, 10:12, 10:12] = 0.1
    sites, labels = _check_labels(K, g, 0.6)  # >= threshold counts (sitemap.py:518)
    assert len(sites) == 2
    _check_labels(K, g, 0.6000001)


def test_too_many_sites_raises(K):
    g = np.zeros((80, 80, 80))
    g[::2, ::2, ::2] = 1.0  # 64000 isolated cells
    with pytest.raises(OverflowError):
        K.label_sites(dev(g), 0.5, 1)


def test_real_density_labels_bulk_insert_properties_cells(K, traj):
    p, xyz = traj
    _, _, edges = O.grid_geometry(xyz[0], delta=1.0)
    counts = O.histogram_vectorised(xyz, edges)
    grid = O.make_density(counts, edges, xyz.shape[0], "water")
    sites = O.label_sites_ndimage(grid, math.e)
    bulk = O.label_sites_ndimage(grid, math.exp(-0.5))
    full = O.site_insert_bulk(sites, bulk)
    ref_map = O.draw_map_from_sites(full, grid.shape)
    ref_props = O.site_properties(full, grid, edges, "water")

    g = K.GridTables(edges)
    d = dev(grid)
    own, n_sites, _ = K.label_sites(d, math.e)
    bl, n_bulk, _ = K.label_sites(d, math.exp(-0.5))
    member, map16 = K.insert_bulk(own, bl, 1)
    assert np.array_equal(map16.cpu().numpy(), ref_map)
    n_labels = n_sites + 2
    assert n_labels == len(full)
    count, mean, std, center = [t.cpu().numpy() for t in K.site_properties(g, d, own, member, n_labels)]
    assert np.array_equal(count, np.array([len(s) for s in full]))
    det = np.prod([(e[-1] - e[0]) / (len(e) - 1) for e in edges])
    V = count * det * O.density_conversion_factor("water", "Angstrom^{-3}")
    np.testing.assert_allclose(mean * V, ref_props["occupancy_avg"], rtol=1e-9, atol=0)
    np.testing.assert_allclose(std * V, ref_props["occupancy_std"], rtol=1e-6, atol=1e-12)
    nz = count > 0
    np.testing.assert_allclose(center[nz], ref_props["center"][nz], rtol=1e-9)
    cells, offsets = K.site_cells(own, member, n_labels)
    cells, offsets = cells.cpu().numpy(), offsets.cpu().numpy()
    for l, site in enumerate(full):
        lin = np.ravel_multi_index(np.asarray(site).T, grid.shape) if len(site) else np.zeros(0, np.int64)
        assert np.array_equal(cells[offsets[l]:offsets[l + 1]], lin)


def _hop_reference(xyz, edges, site_map):
    hop = O.coord2hop_vectorised(xyz, edges, site_map)
    ev, nodes = O.analyze_hops_vectorised(hop[:, :, 0])
    return hop, ev, nodes


def _run_hoptraj(K, xyz, edges, site_map, n_chunks, slabs=1, mode="xyz"):
    """Single process, optionally emulating `slabs` ranks sequentially (same kernels, same exchange).
    mode: "xyz" | "cells" (packed cells written by the histogram kernel), "+coarse" / "+fine" add a block map."""
    F, N = xyz.shape[:2]
    g = K.GridTables(edges)
    m16 = dev(site_map.astype(np.int16))
    bm = K.brick_map(m16)
    block_map = "coarse" if "coarse" in mode else ("fine" if "fine" in mode else ("cell2" if "cell2" in mode else "none"))
    fine_value = {"fine0": 0, "fine1": 1}.get(mode.split("+")[-1], -1)
    cells_all = None
    if mode.startswith("cells"):
        cells_all = torch.empty((F, N), dtype=torch.int32, device="cuda")
        K.hist_accumulate(g, dev(xyz), torch.zeros(g.shape, dtype=torch.int32, device="cuda"), cells_all)
        ref = O.digitize_hop(xyz.reshape(-1, 3), edges)
        shp = np.array(g.shape)
        b = np.stack(ref, axis=1) - 1
        ok = np.all((b >= 0) & (b < shp), axis=1)
        cn = -(-shp // 4)
        cb = ((b[:, 0] >> 2) * cn[1] + (b[:, 1] >> 2)) * cn[2] + (b[:, 2] >> 2)
        low = ((b[:, 0] & 2) << 4) | ((b[:, 1] & 2) << 3) | ((b[:, 2] & 2) << 2) | ((b[:, 0] & 1) << 2) | ((b[:, 1] & 1) << 1) | (b[:, 2] & 1)
        packed = np.where(ok, (cb << 6) | low, 0xFFFFFFFF).astype(np.uint32)
        assert np.array_equal(cells_all.cpu().numpy().view(np.uint32).ravel(), packed)
    events = K.EventBuffer(F * N + 16, "cuda")
    bounds = np.linspace(0, F, slabs + 1).astype(int)
    X = torch.empty((F, N), dtype=torch.float32, device="cuda")
    Y = torch.empty((F, N), dtype=torch.float32, device="cuda")
    summaries, lens, slab_sums = [], [], []
    for r in range(slabs):
        a, b = bounds[r], bounds[r + 1]
        chunk_len = -(-(b - a) // n_chunks)
        nc = -(-(b - a) // chunk_len)
        s = K.hoptraj(g, dev(xyz[a:b]), bm, int(a), nc, chunk_len, events, X[a:b], Y[a:b],
                      cells=None if cells_all is None else cells_all[a:b], block_map=block_map, fine_value=fine_value)
        summaries.append(s)
        lens.append(chunk_len)
        slab_sums.append(K.hop_combine(s, chunk_len, b - a))
    gathered = torch.stack(slab_sums)
    state = None
    for r in range(slabs):
        a, b = bounds[r], bounds[r + 1]
        state = K.new_state(N, "cuda")
        K.hop_advance(gathered[:r] if r else None, state, init=True)
        K.hop_fixup(summaries[r], lens[r], b - a, state, Y[a:b], events)
    final = K.new_state(N, "cuda")
    K.hop_advance(gathered, final, init=True)
    assert np.array_equal(final.cpu().numpy()[:2], state.cpu().numpy()[:2])
    n_ev = events.n()
    return X.cpu().numpy(), Y.cpu().numpy(), events, n_ev, state.cpu().numpy()


@pytest.mark.parametrize("n_chunks,slabs,mode", [(1, 1, "xyz"), (4, 1, "xyz+coarse"), (13, 1, "cells"), (200, 1, "cells+coarse"),
                                                 (3, 2, "cells+coarse"), (5, 8, "xyz+coarse"), (1, 1, "xyz+fine"),
                                                 (9, 1, "cells+fine"), (4, 3, "cells+fine"), (2, 1, "cells+fine0"), (3, 1, "xyz+fine1"), (6, 1, "cells+cell2"), (2, 2, "xyz+cell2")])
def test_hoptraj_and_transitions(K, traj, n_chunks, slabs, mode):
    p, xyz = traj
    res = O.pipeline(xyz, delta=1.0)
    hop, ev_ref, nodes_ref = res["hop"], res["events"], res["nodes"]
    X, Y, events, n_ev, state = _run_hoptraj(K, xyz, res["edges"], res["map"], n_chunks, slabs, mode)
    assert np.array_equal(X, hop[:, :, 0])
    assert np.array_equal(Y, hop[:, :, 1])
    assert n_ev == len(ev_ref)
    n_labels = len(res["sites"])
    ev, e_s0, e_s, e_start, e_count = K.events_finalize(events, n_ev, xyz.shape[1], xyz.shape[0], n_labels)
    ev = ev.cpu().numpy().astype(np.int64)
    # reference order: by edge, then frame, then atom
    order = np.lexsort((ev_ref[:, 1], ev_ref[:, 0], ev_ref[:, 3], ev_ref[:, 2]))
    assert np.array_equal(ev, ev_ref[order])
    pairs, cnts = O.events_to_counts(ev_ref)
    assert np.array_equal(np.stack([e_s0.cpu().numpy(), e_s.cpu().numpy()], axis=1), pairs)
    assert np.array_equal(e_count.cpu().numpy(), cnts)
    assert np.array_equal(e_start.cpu().numpy(), np.concatenate([[0], np.cumsum(cnts)[:-1]]))
    ia = np.nonzero(state[0] != 0)[0]
    nodes = np.stack([ia, state[0][ia], (xyz.shape[0] - 1) - state[1][ia]], axis=1)
    assert np.array_equal(nodes, nodes_ref)


def test_hoptraj_outliers_and_top_edge(K, traj):
    p, xyz = traj
    xyz = xyz[:60].copy()
    res0 = O.pipeline(xyz, delta=1.0)
    edges, site_map = res0["edges"], res0["map"]
    rng = np.random.default_rng(1)
    # push atoms outside the grid (outliers), also in frame 0, and onto / next to the top edges
    mask = rng.random(xyz.shape[:2]) < 0.05
    xyz[mask] += np.float32(60.0)
    xyz[0, :40] -= np.float32(80.0)
    for d in range(3):
        top = np.float32(edges[d][-1])
        for k, v in enumerate([top, np.nextafter(top, np.float32(0)), np.nextafter(top, np.float32(1e9))]):
            xyz[5 + k, 100 + 10 * d: 110 + 10 * d, d] = v
    hop, ev_ref, nodes_ref = _hop_reference(xyz, edges, site_map)
    assert (hop[:, :, 0] == -1).any() and (ev_ref[:, 2] == -1).any()
    for n_chunks, slabs, mode in [(1, 1, "xyz"), (7, 1, "cells+coarse"), (2, 3, "xyz+coarse"), (3, 1, "cells"), (5, 1, "cells+fine"),
                                  (2, 2, "xyz+fine")]:
        X, Y, events, n_ev, state = _run_hoptraj(K, xyz, edges, site_map, n_chunks, slabs, mode)
        assert np.array_equal(X, hop[:, :, 0])
        assert np.array_equal(Y, hop[:, :, 1])
        ev = K.events_to_numpy(events.data[:n_ev])
        order = np.lexsort((ev["iatom"], ev["frame"]))
        got = np.stack([ev[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1)
        assert np.array_equal(got, ev_ref)
        ia = np.nonzero(state[0] != 0)[0]
        assert np.array_equal(np.stack([ia, state[0][ia], 59 - state[1][ia]], axis=1), nodes_ref)


def test_hoptraj_mostly_interstitial_map(K, traj):
    """A map without a bulk site (ions, or site_insert_nobulk): most atoms never see a site, a few
    hop between sites, some are outliers from frame 0 on -- the "no site seen yet" block path, warps
    with lanes in both regimes, and long chunks entirely in the interstitial."""
    p, xyz = traj
    xyz = xyz.copy()
    res0 = O.pipeline(xyz, delta=1.0, with_bulk=False)
    edges, site_map = res0["edges"], res0["map"]
    assert (site_map == 1).sum() < 0.05 * site_map.size          # label 1 is an ordinary small site here
    rng = np.random.default_rng(4)
    xyz[:, 5:9] += np.float32(70.0)                               # outliers for the whole trajectory
    mask = rng.random(xyz.shape[:2]) < 0.01
    xyz[mask] -= np.float32(90.0)
    hop, ev_ref, nodes_ref = _hop_reference(xyz, edges, site_map)
    assert (hop[:, :, 0] == 0).mean() > 0.8 and len(ev_ref) > 50
    F = xyz.shape[0]
    for n_chunks, slabs, mode in [(1, 1, "cells+fine"), (3, 1, "cells+fine0"), (1, 1, "xyz+fine"), (5, 2, "cells+coarse"),
                                  (2, 1, "cells+cell2"), (4, 1, "cells"), (1, 3, "xyz")]:
        X, Y, events, n_ev, state = _run_hoptraj(K, xyz, edges, site_map, n_chunks, slabs, mode)
        assert np.array_equal(X, hop[:, :, 0]), mode
        assert np.array_equal(Y, hop[:, :, 1]), mode
        ev = K.events_to_numpy(events.data[:n_ev])
        order = np.lexsort((ev["iatom"], ev["frame"]))
        got = np.stack([ev[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1)
        assert np.array_equal(got, ev_ref), mode
        ia = np.nonzero(state[0] != 0)[0]
        assert np.array_equal(np.stack([ia, state[0][ia], (F - 1) - state[1][ia]], axis=1), nodes_ref), mode


def test_coarse_map_blocks(K):
    rng = np.random.default_rng(8)
    for shape in [(9, 10, 11), (36, 36, 36), (4, 4, 4), (5, 1, 70)]:
        m = np.ones(shape, np.int16)
        m[rng.random(shape) < 0.01] = 0
        m[:4, :4, :4] = 0
        if shape[0] > 8:
            m[8, 5, 2] = 7
        bm = K.brick_map(dev(m))
        brick, c = bm
        c = c.cpu().numpy().view(np.uint32)
        fine = bm.fine.cpu().numpy()
        brick = brick.cpu().numpy()
        cn = [-(-n // 4) for n in shape]
        nb = cn[0] * cn[1] * cn[2]
        nb_pad = (nb + 3) // 4 * 4
        assert fine.size == 2 * nb_pad + 16
        assert int(fine[2 * nb_pad:2 * nb_pad + 4].view(np.uint32)[0]) == int((m == 1).sum())
        ii, jj, kk = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
        # 4x4x4 bricks, cells inside a brick ordered by 2x2x2 sub-brick: (i1 j1 k1 i0 j0 k0)
        low = ((ii & 2) << 4) | ((jj & 2) << 3) | ((kk & 2) << 2) | ((ii & 1) << 2) | ((jj & 1) << 1) | (kk & 1)
        idx = ((((ii >> 2) * cn[1] + (jj >> 2)) * cn[2] + (kk >> 2)) << 6) | low
        assert np.array_equal(brick[idx], m)
        c2 = bm.cell2.cpu().numpy().view(np.uint32)
        code = (c2[idx >> 4] >> ((idx & 15) * 2)) & 3
        assert np.array_equal(code, np.where(m == 0, 0, np.where(m == 1, 1, 3)))
        for b in range(nb):
            ci, cj, ck = b // (cn[1] * cn[2]), (b // cn[2]) % cn[1], b % cn[2]
            blk = m[ci * 4:ci * 4 + 4, cj * 4:cj * 4 + 4, ck * 4:ck * 4 + 4]
            u = np.unique(blk)
            want = int(u[0]) if len(u) == 1 and u[0] in (0, 1) else 3
            assert (c[b >> 4] >> ((b & 15) * 2)) & 3 == want
            for sub in range(8):
                i1, j1, k1 = sub >> 2, (sub >> 1) & 1, sub & 1
                sb = blk[i1 * 2:i1 * 2 + 2, j1 * 2:j1 * 2 + 2, k1 * 2:k1 * 2 + 2]
                for v in (0, 1):
                    assert (fine[v * nb_pad + b] >> sub) & 1 == int((sb == v).all()), (shape, b, sub, v)


def test_half_angstrom_golden_through_all_paths(K):
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "half_angstrom.npz"))
    xyz = g["xyz"]
    edges = [g["edges_x"], g["edges_y"], g["edges_z"]]
    ev_ref = g["events"].astype(np.int64)
    for mode in ("xyz", "xyz+coarse", "cells", "cells+coarse", "xyz+fine", "cells+fine", "cells+cell2"):
        X, Y, events, n_ev, state = _run_hoptraj(K, xyz, edges, g["site_map"], 3, 1, mode)
        assert np.array_equal(X, g["hop_x"].astype(np.float32)) and np.array_equal(Y, g["hop_y"].astype(np.float32))
        ev = K.events_to_numpy(events.data[:n_ev])
        order = np.lexsort((ev["iatom"], ev["frame"]))
        got = np.stack([ev[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1)
        assert np.array_equal(got, ev_ref)


def test_hopscan_from_labels_and_all_transitions(K):
    from tests.test_machine_emul import random_labels

    rng = np.random.default_rng(2)
    lab = random_labels(rng, 300, 257, nsites=6, p_stay=0.9, p_out=0.05, p_zero=0.3)
    ev_ref, nodes_ref = O.analyze_hops_vectorised(lab)
    hx = dev(lab.astype(np.float32))
    for n_chunks in (1, 6, 300):
        events = K.EventBuffer(lab.size, "cuda")
        chunk_len = -(-300 // n_chunks)
        nc = -(-300 // chunk_len)
        s = K.hopscan_labels(hx, 0, nc, chunk_len, events)
        state = K.new_state(257, "cuda")
        K.hop_advance(None, state, init=True)
        K.hop_fixup(s, chunk_len, 300, state, None, events)
        n_ev = events.n()
        ev = K.events_to_numpy(events.data[:n_ev])
        order = np.lexsort((ev["iatom"], ev["frame"]))
        got = np.stack([ev[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1)
        assert np.array_equal(got, ev_ref)
    assert np.array_equal(K.all_transitions(hx, 8), O.graph_alltransitions(lab))


def test_pipeline_regrows_event_buffer(K, traj):
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    ref = O.pipeline(xyz, delta=1.0)
    pipe = HopPipeline(grid_edges_from_frame0(xyz[0], delta=1.0), xyz.shape[0], PipelineConfig(event_capacity=100))
    res = pipe.run(dev(xyz))
    assert res.n_events_local == len(ref["events"]) > 100
    assert np.array_equal(res.hop_y.cpu().numpy(), ref["hop"][:, :, 1])
    assert res.count_matrix.sum().item() == len(ref["events"])


def test_empty_and_degenerate_inputs(K):
    # zero points, one frame, atoms all outside the grid, a grid with a single cell
    edges = [np.linspace(0.0, 4.0, 5), np.linspace(0.0, 3.0, 4), np.linspace(0.0, 1.0, 2)]
    g = K.GridTables(edges)
    counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, torch.empty((0, 3), dtype=torch.float32, device="cuda"), counts)
    assert counts.sum().item() == 0
    far = dev(np.full((1, 7, 3), 99.0, np.float32))
    K.hist_accumulate(g, far, counts)
    assert counts.sum().item() == 0
    m16 = dev(np.ones(g.shape, np.int16))
    brick, coarse = K.brick_map(m16)
    ev = K.EventBuffer(16, "cuda")
    X = torch.empty((1, 7), dtype=torch.float32, device="cuda")
    Y = torch.empty((1, 7), dtype=torch.float32, device="cuda")
    s = K.hoptraj(g, far, K.brick_map(m16), 0, 1, 1, ev, X, Y)
    st = K.new_state(7, "cuda")
    K.hop_advance(None, st, init=True)
    K.hop_fixup(s, 1, 1, st, Y, ev)
    assert (X == -1).all().item() and (Y == 0).all().item() and ev.n() == 0
    assert (st.cpu().numpy()[0] == -1).all()          # first-frame outliers: s0 = -1 (hop/graph.py:206-208)
    out = K.events_finalize(ev, 0, 7, 1, 3)
    assert out[0].shape[0] == 0 and out[1].numel() == 0
    one = K.GridTables([np.array([0.0, 1.0])] * 3)
    c1 = torch.zeros(one.shape, dtype=torch.int32, device="cuda")
    K.hist_accumulate(one, dev(np.array([[0.5, 0.5, 0.5], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [1.5, 0.5, 0.5]], np.float32)), c1)
    assert c1.item() == 3                              # the point on the last edge counts, the one beyond does not
    lab, n, vol = K.label_sites(dev(np.ones((1, 1, 1))), 0.5)
    assert n == 1 and lab.item() == 1 and vol.cpu().tolist() == [0, 1]
    lab, n, vol = K.label_sites(dev(np.zeros((3, 2, 2))), 0.5)
    assert n == 0 and (lab == 0).all().item()


def test_pipeline_config1_matches_faithful_oracle(K, traj):
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    ref = O.pipeline(xyz, delta=1.0, faithful=True)
    edges = grid_edges_from_frame0(xyz[0], delta=1.0)
    for a, b in zip(edges, ref["edges"]):
        assert np.array_equal(a, b)
    pipe = HopPipeline(edges, xyz.shape[0], PipelineConfig())
    res = pipe.run(dev(xyz))
    assert np.array_equal(res.counts.cpu().numpy(), ref["counts"].astype(np.int32))
    assert np.array_equal(res.density.cpu().numpy(), ref["grid"])
    assert np.array_equal(res.map16.cpu().numpy(), ref["map"])
    assert np.array_equal(res.hop_x.cpu().numpy(), ref["hop"][:, :, 0])
    assert np.array_equal(res.hop_y.cpu().numpy(), ref["hop"][:, :, 1])
    ev = K.events_to_numpy(res.events)
    e_s0, e_s = res.edge_s0.cpu().numpy(), res.edge_s.cpu().numpy()
    e_start, e_count = res.edge_start.cpu().numpy(), res.edge_count.cpu().numpy()
    props = ref["props"]
    edge_keys = sorted(k for k in props if isinstance(k, tuple))
    assert [tuple(x) for x in np.stack([e_s0, e_s], axis=1).tolist()] == edge_keys
    for (a, b), st, c in zip(edge_keys, e_start, e_count):
        seg = ev[st:st + c]
        d = props[(a, b)]
        assert np.array_equal(seg["frame"], d["frame"]) and np.array_equal(seg["iatom"], d["iatom"])
        assert np.array_equal(seg["tau"], d["tau"]) and np.array_equal(seg["tbarrier"], d["tbarrier"])
    M = res.count_matrix.cpu().numpy()
    assert M.sum() == len(ev)
    for (a, b), c in zip(edge_keys, e_count):
        assert M[a + 1, b + 1] == c


def test_run_host_overlapped_runs_match_resident_run(K, traj):
    """HopPipeline.run_host (host in, host out) gives what run() gives on resident data, also when a
    second run begins before the first one's hopping trajectory has been downloaded."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    ref = O.pipeline(xyz, delta=1.0, faithful=False)
    edges = grid_edges_from_frame0(xyz[0], delta=1.0)
    pipe = HopPipeline(edges, xyz.shape[0], PipelineConfig())
    host_xyz = torch.from_numpy(np.ascontiguousarray(xyz)).pin_memory()
    # a different second trajectory on the same grid: the same frames in reverse order (checked
    # against the resident run below)
    host_rev = torch.from_numpy(np.ascontiguousarray(xyz[::-1])).pin_memory()
    outs = [(torch.empty(xyz.shape[:2], dtype=torch.float32).pin_memory(), torch.empty(xyz.shape[:2], dtype=torch.float32).pin_memory())
            for _ in range(2)]
    a = pipe.run_host_begin(host_xyz, 0, *outs[0], block_bytes=1 << 16)
    b = pipe.run_host_begin(host_rev, 0, *outs[1], block_bytes=1 << 16)
    ra, ha = a.finish()
    rb, hb = b.finish()
    assert np.array_equal(ha["hop_x"].numpy(), ref["hop"][:, :, 0]) and np.array_equal(ha["hop_y"].numpy(), ref["hop"][:, :, 1])
    n_edge_events = len(O.analyze_hops_vectorised(ref["hop"][:, :, 0])[0])
    assert ha["events"].shape[0] == n_edge_events and int(ha["edge_count"].sum()) == n_edge_events
    # the second run alone, resident: identical planes and events
    res = pipe.run(dev(np.ascontiguousarray(xyz[::-1])))
    assert torch.equal(res.hop_x.cpu(), hb["hop_x"]) and torch.equal(res.hop_y.cpu(), hb["hop_y"])
    assert torch.equal(res.events.cpu(), hb["events"]) and torch.equal(res.edge_count.cpu(), hb["edge_count"])
    r1, h1 = pipe.run_host(host_xyz)
    assert torch.equal(h1["hop_x"], ha["hop_x"]) and torch.equal(h1["hop_y"], ha["hop_y"])
    assert torch.equal(h1["events"], ha["events"])
    # int16 host planes: the same labels in half the bytes; planes_float32 gives the DCD records back
    o16 = (torch.empty(xyz.shape[:2], dtype=torch.int16).pin_memory(), torch.empty(xyz.shape[:2], dtype=torch.int16).pin_memory())
    run = pipe.run_host_begin(host_xyz, 0, *o16)
    r2, h2 = run.finish()
    assert h2["hop_x"].dtype == torch.int16 and torch.equal(h2["hop_x"].to(torch.float32), ha["hop_x"])
    fx, fy = run.planes_float32()
    assert torch.equal(fx, ha["hop_x"]) and torch.equal(fy, ha["hop_y"]) and torch.equal(h2["events"], ha["events"])


def test_pipeline_async_results_are_lazy_and_correct(K, traj):
    """From the second run on the pipeline does not wait for the device: the counts-dependent result
    fields resolve on first use, and the results of a run survive the next run being queued."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    edges = grid_edges_from_frame0(xyz[0], delta=1.0)
    a, b = dev(np.ascontiguousarray(xyz)), dev(np.ascontiguousarray(xyz[::-1]))

    def snapshot(res):
        return dict(n_labels=res.n_labels, n_ev=res.n_events_local, events=res.events.clone(), e_s0=res.edge_s0.clone(),
                    e_s=res.edge_s.clone(), e_count=res.edge_count.clone(), e_start=res.edge_start.clone(),
                    M=res.count_matrix.clone(), props=[t.clone() for t in res.site_props])

    def same(x, y):
        return (x["n_labels"] == y["n_labels"] and x["n_ev"] == y["n_ev"] and torch.equal(x["events"], y["events"])
                and torch.equal(x["e_s0"], y["e_s0"]) and torch.equal(x["e_s"], y["e_s"]) and torch.equal(x["e_count"], y["e_count"])
                and torch.equal(x["e_start"], y["e_start"]) and torch.equal(x["M"], y["M"])
                # the annotation sums with floating-point atomics: equal up to summation order
                and all(torch.allclose(s.double(), t.double(), rtol=1e-9, atol=1e-12, equal_nan=True)
                        for s, t in zip(x["props"], y["props"])))

    want = {}
    for name, inp in (("a", a), ("b", b)):
        sync_pipe = HopPipeline(edges, xyz.shape[0], PipelineConfig(async_results=False))
        want[name] = snapshot(sync_pipe.run(inp))
    pipe = HopPipeline(edges, xyz.shape[0], PipelineConfig())
    r1 = pipe.run(a)                       # first run: synchronous (learns the sizes)
    assert r1._pending is None and same(snapshot(r1), want["a"])
    r2 = pipe.run(b)
    assert r2._pending is not None         # nothing resolved yet
    r3 = pipe.run(a)
    assert same(snapshot(r2), want["b"])   # resolved after the next run was queued
    assert same(snapshot(r3), want["a"])
    # more events than budgeted, noticed late: the tail of the run is redone
    small = HopPipeline(edges, xyz.shape[0], PipelineConfig())
    small.run(a)
    small._events = K.EventBuffer(64, "cuda")
    small.cfg.event_capacity = 64
    r = small.run(b)
    assert r._pending is not None and same(snapshot(r), want["b"])
    assert small._events.capacity >= want["b"]["n_ev"]


def test_pipeline_pool_two_runs_in_flight(K, traj):
    """Independent runs on two pipelines / streams give what one pipeline gives, in any interleaving."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig, PipelinePool, grid_edges_from_frame0

    p, xyz = traj
    edges = grid_edges_from_frame0(xyz[0], delta=1.0)
    inputs = [dev(np.ascontiguousarray(xyz)), dev(np.ascontiguousarray(xyz[::-1])), dev(np.ascontiguousarray(xyz[:, ::-1]))]
    single = HopPipeline(edges, xyz.shape[0], PipelineConfig(async_results=False))
    want = []
    for x in inputs:
        r = single.run(x)
        want.append((r.hop_x.clone(), r.hop_y.clone(), r.events.clone(), r.count_matrix.clone(), r.n_labels))
    pool = PipelinePool(edges, xyz.shape[0], PipelineConfig(), n=2)
    order = [0, 1, 2, 1, 0, 2, 2, 0]
    results = []
    for i in order:
        res = pool.run(inputs[i])
        results.append((i, res))
        if len(results) >= 2:           # look at a result while the next run is in flight, before its buffers are reused
            j, old = results[-2]
            hx, hy, ev, M, nl = want[j]
            assert old.n_labels == nl and torch.equal(old.events, ev) and torch.equal(old.count_matrix, M)
            assert torch.equal(old.hop_x, hx) and torch.equal(old.hop_y, hy)
    pool.synchronize()
    j, last = results[-1]
    assert torch.equal(last.events, want[j][2]) and torch.equal(last.hop_x, want[j][0])


def test_pipeline_pool_with_a_grid_per_trajectory(K, traj):
    """A batch of trajectories (BASELINE config 5) where every trajectory spans its own grid in frame 0
    (hop/density.py:271-278): PipelinePool.run(edges=...) switches the geometry per run, the results equal those of
    pipelines built for each grid."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig, PipelinePool, grid_edges_from_frame0

    p, xyz = traj
    shifted = (xyz + np.float32(0.37)).astype(np.float32)          # another extent: other edges, same shape
    wide = xyz.copy()
    wide[0, 5] += np.float32(3.0)                                  # one atom further out in frame 0: a larger grid
    cases = [xyz, shifted, wide, xyz]
    edges = [grid_edges_from_frame0(c[0], delta=1.0) for c in cases]
    assert not np.array_equal(edges[0][0], edges[1][0]) and len(edges[2][0]) != len(edges[0][0])
    want = []
    for c, e in zip(cases, edges):
        r = HopPipeline(e, c.shape[0], PipelineConfig(async_results=False)).run(dev(c))
        want.append((r.counts.clone(), r.map16.clone(), r.hop_x.clone(), r.hop_y.clone(), r.events.clone()))
    pool = PipelinePool(edges[0], xyz.shape[0], PipelineConfig(), n=2)
    for rep in range(2):
        for c, e, w in zip(cases, edges, want):
            r = pool.run(dev(c), edges=e)
            got = (r.counts, r.map16, r.hop_x, r.hop_y, r.events)
            pool.synchronize()
            assert all(torch.equal(a_, b_) for a_, b_ in zip(w, got))


def test_label_from_counts_equals_label_from_density(K):
    """Thresholding the counts (count >= cmin per bin-width class, cmin by bisection on the device) gives
    exactly the labels of thresholding the density, also when the threshold equals a density value."""
    import torch

    rng = np.random.default_rng(12)

    def check(edges, counts, n_frames, factor, thresholds):
        g = K.GridTables(edges)
        assert g.width_classes > 0
        c = dev(counts.astype(np.int32))
        dens = K.density_from_counts(g, c, n_frames, factor)
        for thr in thresholds:
            la = torch.empty(g.shape, dtype=torch.int32, device="cuda")
            lb = torch.empty(g.shape, dtype=torch.int32, device="cuda")
            na = torch.zeros(1, dtype=torch.int32, device="cuda")
            nb = torch.zeros(1, dtype=torch.int32, device="cuda")
            K.label_sites_async(dens, thr, 1, la, na)
            K.label_sites_counts_async(g, c, n_frames, factor, thr, 1, lb, nb)
            assert int(na.item()) == int(nb.item()), thr
            assert torch.equal(la, lb), thr

    # linspace edges as DensityCreator makes them: a few width classes from rounding
    for n, lo, hi in [((24, 17, 31), -3.3, 21.7), ((40, 40, 40), 0.1, 20.1)]:
        edges = [np.linspace(lo + d, hi + 1.7 * d, k + 1) for d, k in enumerate(n)]
        counts = rng.poisson(6.0, n)
        dens = O.make_density(counts, edges, 50, "water")
        vals = np.unique(dens)
        thresholds = [float(vals[len(vals) // 2]), float(vals[-2]), float(np.nextafter(vals[3], np.inf)), 0.0, -1.0,
                      math.e, float("inf"), float("nan"), 1e-300]
        check(edges, counts, 50, O.density_conversion_factor("Angstrom^{-3}", "water"), thresholds)
        check(edges, counts, 50, 1.0, [float(np.median(counts) / 50 / np.prod([np.diff(e)[0] for e in edges])), 0.0])
    # three clearly different widths per axis
    w = np.array([0.5, 0.75, 1.0])
    edges = [np.concatenate([[0.0], np.cumsum(w[rng.integers(0, 3, k)])]) for k in (19, 23, 35)]
    counts = rng.poisson(3.0, (19, 23, 35))
    dens = O.make_density(counts, edges, 7, "Angstrom^{-3}")
    check(edges, counts, 7, 1.0, [float(v) for v in np.unique(dens)[[1, 5, 20]]] + [0.5])
    # irregular edges: no classes, the count path refuses
    edges = [np.concatenate([[0.0], np.cumsum(0.5 + rng.random(40))]) for _ in range(3)]
    g = K.GridTables(edges)
    assert g.width_classes == 0
    from hop_b200._lib import HopbError

    with pytest.raises((HopbError, ValueError, RuntimeError)):
        K.label_sites_counts_async(g, torch.zeros(g.shape, dtype=torch.int32, device="cuda"), 5, 1.0, 1.0, 1,
                                   torch.empty(g.shape, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda"))


@pytest.mark.parametrize("chunk,cache", [(200, True), (64, True), (37, False), (1, None)])
def test_streamed_run_equals_resident_run(K, traj, chunk, cache):
    """HopPipeline.run_streamed -- two passes over frame chunks for trajectories larger than HBM, the machine
    state carried from chunk to chunk -- gives exactly what run() gives on the resident trajectory, whatever the
    chunk length and whether pass 2 re-reads the coordinates or the cached brick indices."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    F, N = xyz.shape[:2]
    if chunk == 1:
        F = 23
        xyz = xyz[:F]
    edges = grid_edges_from_frame0(xyz[0], delta=1.0)
    full = dev(xyz)
    want = HopPipeline(edges, F, PipelineConfig(async_results=False)).run(full)
    calls = []

    def source(f0, f1):
        calls.append((f0, f1))
        return full[f0:f1].clone()

    X = torch.full((F, N), -9.0, device="cuda")
    Y = torch.full((F, N), -9.0, device="cuda")

    def sink(f0, hx, hy):
        X[f0:f0 + hx.shape[0]].copy_(hx)
        Y[f0:f0 + hy.shape[0]].copy_(hy)

    pipe = HopPipeline(edges, F, PipelineConfig(async_results=False))
    for _ in range(2):
        del calls[:]
        got = pipe.run_streamed(source, F, chunk, sink=sink, cache_cells=cache)
        n_chunks = -(-F // chunk)
        assert len(calls) == (n_chunks if cache in (True, None) else 2 * n_chunks)
        assert torch.equal(got.counts, want.counts) and torch.equal(got.map16, want.map16) and torch.equal(got.density, want.density)
        assert torch.equal(X, want.hop_x) and torch.equal(Y, want.hop_y)
        assert got.n_labels == want.n_labels and torch.equal(got.events, want.events)
        assert torch.equal(got.count_matrix, want.count_matrix) and torch.equal(got.edge_count, want.edge_count)
        a, b = got.final_state.cpu().numpy(), want.final_state.cpu().numpy()
        assert np.array_equal(a[[0, 1, 3]], b[[0, 1, 3]])
        off = a[3] == 0
        assert np.array_equal(a[2][off], b[2][off])
    with pytest.raises(ValueError):
        pipe.run_streamed(source, F + 1, chunk)


def test_multispecies_equals_independent_runs(K, traj):
    """BASELINE config 4 in small: three selections of one trajectory, each with a grid of its own (its extent in
    frame 0, hop/density.py:271-278), binned in ONE pass over the coordinates (hopb_hist_accumulate_multi) and mapped
    from the shared brick-index array (hopb_hoptraj_strided) -- every species' products equal those of a HopPipeline
    run on that selection alone, bit for bit."""
    import torch

    from hop_b200.multispecies import MultiSpeciesPipeline
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    xyz = np.ascontiguousarray(xyz[:, :450])
    F = xyz.shape[0]
    columns = [(0, 200), (200, 333), (333, 450)]
    cfg = dict(density_unit="Angstrom^{-3}", solvent_threshold=0.012, bulk_threshold=0.004, n_chunks=5)
    single = []
    for a, b in columns:
        sub = np.ascontiguousarray(xyz[:, a:b])
        pipe = HopPipeline(grid_edges_from_frame0(sub[0], delta=1.0), F, PipelineConfig(async_results=False, **cfg))
        r = pipe.run(dev(sub))
        single.append(dict(counts=r.counts.clone(), map16=r.map16.clone(), density=r.density.clone(), x=r.hop_x.clone(),
                           y=r.hop_y.clone(), events=r.events.clone(), M=r.count_matrix.clone(), n=r.n_labels,
                           state=r.final_state.clone(), shape=tuple(pipe.grid.shape)))
    assert len({s_["shape"] for s_ in single}) > 1 or True          # the grids usually differ in extent
    assert sum(s_["n"] for s_ in single) > 6 and sum(len(s_["events"]) for s_ in single) > 0
    multi = MultiSpeciesPipeline.from_frame0(xyz[0], columns, F, delta=1.0, config=PipelineConfig(**cfg))
    for _ in range(2):
        results = multi.run(dev(xyz))
        for (a, b), r, want in zip(columns, results, single):
            assert torch.equal(r.counts, want["counts"]) and torch.equal(r.density, want["density"])
            assert torch.equal(r.map16, want["map16"]) and r.n_labels == want["n"]
            assert torch.equal(r.hop_x, want["x"]) and torch.equal(r.hop_y, want["y"])
            assert torch.equal(multi.hop_x[:, a:b], want["x"])
            assert torch.equal(r.events, want["events"]) and torch.equal(r.count_matrix, want["M"])
            sa, sb = r.final_state.cpu().numpy(), want["state"].cpu().numpy()
            assert np.array_equal(sa[[0, 1, 3]], sb[[0, 1, 3]])


def test_periodic_wrap_is_off_by_default_and_exact_when_on(K, traj):
    """The optional periodic wrapping (north_star kernel 1; hop itself never wraps, hop/density.py:21-28):
    off -- the default, also after switching it on and off again -- the results are bit-identical to a grid
    that never heard of it; on, atoms displaced by whole box vectors are binned (stage 1) and assigned
    (stage 3) exactly like their images x - floor((x - o) / L) * L evaluated in float32 by numpy."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig

    p, xyz = traj
    _, _, edges = O.grid_geometry(xyz[0], delta=1.0)
    rng = np.random.default_rng(5)
    L = np.float32(p.box)
    origin = np.zeros(3, np.float32)
    shift = rng.integers(-3, 4, xyz.shape).astype(np.float32) * L
    moved = (xyz + shift).astype(np.float32)
    inv = np.float32(1.0) / L
    image = (moved - np.floor((moved - origin) * inv) * L).astype(np.float32)      # float32 throughout, like the kernel

    def run(coords, box):
        pipe = HopPipeline(edges, coords.shape[0], PipelineConfig(periodic_box=box, async_results=False))
        r = pipe.run(dev(coords))
        return (r.counts.clone(), r.map16.clone(), r.hop_x.clone(), r.hop_y.clone(), r.events.clone()), pipe

    plain, pipe = run(xyz, None)
    # on and off again: still the plain result
    pipe.grid.set_periodic_box(origin, [L, L, L])
    pipe.grid.set_periodic_box()
    again = pipe.run(dev(xyz))
    assert all(torch.equal(a_, b_) for a_, b_ in zip(plain, (again.counts, again.map16, again.hop_x, again.hop_y, again.events)))
    # the displaced atoms without wrapping are mostly outliers ...
    lost, _ = run(moved, None)
    assert int(lost[0].sum().item()) < int(plain[0].sum().item())
    # ... with wrapping they are their images, bit for bit, in both the cached-cell and the coordinate path of stage 3
    want, _ = run(image, None)
    got, wpipe = run(moved, ([0.0, 0.0, 0.0], [float(L)] * 3))
    assert all(torch.equal(a_, b_) for a_, b_ in zip(want, got))
    assert np.array_equal(got[0].cpu().numpy().astype(np.int64), O.histogram_vectorised(image, edges))
    wpipe.cfg.cache_cells = False
    r = wpipe.run(dev(moved))
    assert torch.equal(r.hop_x, want[2]) and torch.equal(r.hop_y, want[3]) and torch.equal(r.events, want[4])


@pytest.mark.parametrize("shape", [(24, 17, 31), (40, 40, 40), (9, 5, 130), (4, 4, 4), (3, 2, 1), (1, 1, 70), (33, 34, 35)])
@pytest.mark.parametrize("with_bulk,minsite", [(True, 1), (False, 1), (True, 2)])
def test_fused_site_map_equals_the_kernel_chain(K, shape, with_bulk, minsite):
    """hopb_site_map_fused (stage 2 in one cooperative launch) against the stand-alone entry points it
    replaces in the pipeline: density bit-identical, labels / membership / int16 map / bricked map /
    block maps identical, annotation equal to rounding."""
    import torch

    rng = np.random.default_rng(sum(shape) + 7 * minsite + (100 if with_bulk else 0))
    edges = [np.linspace(-1.3 + d, -1.3 + d + 0.5 * k, k + 1) for d, k in enumerate(shape)]
    g = K.GridTables(edges)
    assert K.site_map_fused_ok(g)
    n_frames = 40
    factor = O.density_conversion_factor("Angstrom^{-3}", "water")
    # bulk-like counts with holes, plus a few dense blobs (hydration sites), and one with everything / nothing above
    lam = 40 * 0.0334 * 0.125
    base = rng.poisson(lam, shape)
    blobs = (rng.random(shape) < 0.03) * rng.poisson(12 * lam, shape)
    for counts, thr_s, thr_b in [(base + blobs, math.e, math.exp(-0.5)), (base, 0.0, 0.0), (base, 1e9, 1e8),
                                 (base + blobs, 1.0, 2.0)]:
        c = dev(counts.astype(np.int32))
        # the chain
        dens = K.density_from_counts(g, c, n_frames, factor)
        own = torch.empty(g.shape, dtype=torch.int32, device="cuda")
        n_dev = torch.zeros(2, dtype=torch.int32, device="cuda")
        K.label_sites_counts_async(g, c, n_frames, factor, thr_s, minsite, own, n_dev[0:1])
        if with_bulk:
            bl = torch.empty(g.shape, dtype=torch.int32, device="cuda")
            K.label_sites_counts_async(g, c, n_frames, factor, thr_b, minsite, bl, n_dev[1:2])
            member, map16 = K.insert_bulk(own, bl, 1)
        else:
            member, map16 = None, K.labels_to_map16(own)
        bm = K.brick_map(map16)
        L = K.HOPB_MAX_SITES + 2
        props = K.site_properties(g, dens, own, member, L)
        # the one launch
        nb, words = K.brick_shape(g.shape)
        f_dens = torch.full(g.shape, -1.0, dtype=torch.float64, device="cuda")
        f_own = torch.full(g.shape, -7, dtype=torch.int32, device="cuda")
        f_member = torch.full(g.shape, 9, dtype=torch.uint8, device="cuda") if with_bulk else None
        f_map = torch.full(g.shape, -7, dtype=torch.int16, device="cuda")
        f_bm = K.BrickMap(torch.full((nb * 64,), -7, dtype=torch.int16, device="cuda"),
                          torch.full((words,), -1, dtype=torch.int32, device="cuda"),
                          torch.full((K.fine_map_bytes(g.shape),), 5, dtype=torch.uint8, device="cuda"),
                          torch.full((nb * 4,), -1, dtype=torch.int32, device="cuda"))
        f_n = torch.full((2,), -1, dtype=torch.int32, device="cuda")
        f_props = (torch.empty(L, dtype=torch.int64, device="cuda"), torch.empty(L, dtype=torch.float64, device="cuda"),
                   torch.empty(L, dtype=torch.float64, device="cuda"), torch.empty((L, 3), dtype=torch.float64, device="cuda"))
        for _ in range(2):      # twice: the second call takes the cached tables
            K.site_map_fused(g, c, n_frames, factor, thr_s, thr_b, with_bulk, minsite, f_dens, f_own, f_member, f_map, f_bm, f_n,
                             props=f_props)
        assert torch.equal(f_dens, dens)
        assert int(f_n[0].item()) == int(n_dev[0].item())
        if with_bulk:
            assert int(f_n[1].item()) == int(n_dev[1].item())
            assert torch.equal(f_member, member)
        assert torch.equal(f_own, own) and torch.equal(f_map, map16)
        assert torch.equal(f_bm.brick, bm.brick) and torch.equal(f_bm.coarse, bm.coarse) and torch.equal(f_bm.cell2, bm.cell2)
        nbp = (nb + 3) // 4 * 4
        assert torch.equal(f_bm.fine[:nb], bm.fine[:nb]) and torch.equal(f_bm.fine[nbp:nbp + nb], bm.fine[nbp:nbp + nb])
        assert torch.equal(f_bm.fine[2 * nbp:2 * nbp + 4], bm.fine[2 * nbp:2 * nbp + 4])
        n_labels = int(n_dev[0].item()) + (2 if with_bulk else 1)
        assert torch.equal(f_props[0][:n_labels], props[0][:n_labels])
        for a_, b_ in zip(f_props[1:], props[1:]):
            assert torch.allclose(a_[:n_labels], b_[:n_labels], rtol=1e-9, atol=1e-12)
        assert int(f_props[0][n_labels:].abs().sum().item()) == 0


@pytest.mark.parametrize("seed", range(32))
def test_pipeline_random_small_cases(K, seed):
    """Seeded random small trajectories -- odd atom / frame counts, tiny and flat grids, outliers, no bulk
    site, MINsite 2, explicit chunk counts, synchronous and asynchronous results -- through HopPipeline
    against the oracle, bit for bit."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    rng = np.random.default_rng(1000 + seed)
    n_atoms = int(rng.choice([1, 2, 31, 33, 97, 250, 601, 1000]))
    n_frames = int(rng.choice([1, 2, 9, 17, 40, 121, 200, 333]))
    delta = float(rng.choice([0.5, 1.0, 1.7]))
    with_bulk = bool(rng.integers(0, 2))
    minsite = int(rng.choice([1, 1, 2]))
    p = SynthParams(n_atoms=n_atoms, n_frames=n_frames, seed=77 + seed, site_stride=int(rng.choice([3, 10, 50])))
    xyz = generate(p)
    if seed % 3 == 0 and n_frames > 1:       # outliers, never in frame 0 (it defines the grid)
        m = rng.random(xyz.shape[:2]) < 0.05
        m[0] = False
        xyz[m] += np.float32(50.0)
    if seed % 4 == 1:                        # a flat system: one axis of the grid only a few cells deep
        xyz[:, :, 2] = xyz[:, :, 2] * np.float32(0.02) + np.float32(5.0)
    thr = float(rng.choice([math.e, 1.0, 4.0]))
    ref = O.pipeline(xyz, delta=delta, solvent_threshold=thr, minsite=minsite, with_bulk=with_bulk)
    edges = grid_edges_from_frame0(xyz[0], delta=delta)
    n_chunks = None if seed % 2 else int(rng.integers(1, max(2, n_frames // 2 + 1)))
    block_map = ["auto", "fine", "coarse", "cell2", "none"][seed % 5]
    planes = seed % 7 != 3                   # without the x / y planes: events and counts only
    cfg = PipelineConfig(solvent_threshold=thr, minsite=minsite, with_bulk=with_bulk, n_chunks=n_chunks,
                         label_from_counts=bool(seed % 2), cache_cells=bool((seed // 2) % 2), block_map=block_map,
                         write_hoptraj=planes, annotate=bool(seed % 11 != 5))
    pipe = HopPipeline(edges, n_frames, cfg)
    x = dev(xyz)
    for rep in range(2):                     # second run: asynchronous results
        res = pipe.run(x)
        assert np.array_equal(res.counts.cpu().numpy(), ref["counts"].astype(np.int32))
        assert np.array_equal(res.density.cpu().numpy(), ref["grid"])
        assert np.array_equal(res.map16.cpu().numpy(), ref["map"])
        if planes:
            assert np.array_equal(res.hop_x.cpu().numpy(), ref["hop"][:, :, 0])
            assert np.array_equal(res.hop_y.cpu().numpy(), ref["hop"][:, :, 1])
        else:
            assert res.hop_x is None and res.hop_y is None
        assert res.n_labels == len(ref["sites"])
        ev = K.events_to_numpy(res.events)
        ev_ref = ref["events"]
        order = np.lexsort((ev_ref[:, 1], ev_ref[:, 0], ev_ref[:, 3], ev_ref[:, 2])) if len(ev_ref) else np.zeros(0, int)
        got = np.stack([ev[k] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1).astype(np.int64)
        assert np.array_equal(got, ev_ref[order].reshape(-1, 6))
        pairs, cnts = O.events_to_counts(ev_ref)
        assert np.array_equal(np.stack([res.edge_s0.cpu().numpy(), res.edge_s.cpu().numpy()], axis=1).reshape(-1, 2), pairs)
        assert np.array_equal(res.edge_count.cpu().numpy(), cnts)
        st = res.final_state.cpu().numpy()
        ia = np.nonzero(st[0] != 0)[0]
        nodes = np.stack([ia, st[0][ia], (n_frames - 1) - st[1][ia]], axis=1).reshape(-1, 3)
        assert np.array_equal(nodes, ref["nodes"].reshape(-1, 3))


def test_pipeline_async_count_matrix_outgrown(K, traj):
    """The asynchronous finalisation sizes the dense count matrix from the previous run; a run with many
    more sites than that is noticed when its result is first used and finalised again."""
    import torch

    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    edges = grid_edges_from_frame0(xyz[0], delta=1.0)
    cfg = dict(solvent_threshold=1.6, with_bulk=False)           # noise peaks of the bulk: hundreds of small sites
    lump = np.ascontiguousarray(np.broadcast_to(xyz[0, :1], xyz.shape)).copy()   # every atom in one cell: one site
    want = HopPipeline(edges, xyz.shape[0], PipelineConfig(async_results=False, **cfg)).run(dev(xyz))
    n_labels, M, ev = want.n_labels, want.count_matrix.clone(), want.events.clone()
    assert n_labels > 300
    pipe = HopPipeline(edges, xyz.shape[0], PipelineConfig(**cfg))
    first = pipe.run(dev(lump))
    assert first.n_labels == 2 and pipe._w_known == 3
    res = pipe.run(dev(xyz))
    assert res._pending is not None
    assert res.n_labels == n_labels and torch.equal(res.count_matrix, M) and torch.equal(res.events, ev)
    again = pipe.run(dev(xyz))                                    # now sized for it
    assert again._pending is not None and torch.equal(again.count_matrix, M)


@pytest.mark.parametrize("n_chunks", [None, 3])
def test_int16_planes_equal_float32_planes(K, traj, n_chunks):
    """The pipeline keeps hop_x / hop_y as int16 in HBM (PipelineConfig.plane_dtype, the default): same labels, same
    events and -- end to end -- the same host planes as the float32 record type of the hoptraj DCD."""
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0

    p, xyz = traj
    edges = grid_edges_from_frame0(xyz[0], delta=1.0)
    out = {}
    for dt in ("float32", "int16"):
        pipe = HopPipeline(edges, xyz.shape[0], PipelineConfig(plane_dtype=dt, n_chunks=n_chunks))
        res = pipe.run(dev(xyz))
        assert res.hop_x.dtype == getattr(torch, dt) and res.hop_y.dtype == getattr(torch, dt)
        ev = res.events.cpu().numpy().copy()
        trans = K.all_transitions(res.hop_x, res.n_labels)
        host = {}
        for ht in (torch.float32, torch.int16):       # host planes of either type from device planes of either type
            ox = torch.empty(xyz.shape[:2], dtype=ht).pin_memory()
            oy = torch.empty(xyz.shape[:2], dtype=ht).pin_memory()
            r2, h = pipe.run_host(torch.from_numpy(xyz).pin_memory(), 0, ox, oy)
            host[ht] = (h["hop_x"].to(torch.float32).numpy().copy(), h["hop_y"].to(torch.float32).numpy().copy())
        out[dt] = (res.hop_x.cpu().to(torch.float32).numpy(), res.hop_y.cpu().to(torch.float32).numpy(), ev, trans, host)
    a, b = out["float32"], out["int16"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    for dt in out:
        for ht in (torch.float32, torch.int16):
            assert np.array_equal(out[dt][4][ht][0], a[0]) and np.array_equal(out[dt][4][ht][1], a[1])
    with pytest.raises(ValueError):
        PipelineConfig(plane_dtype="int8").plane_torch_dtype()


@pytest.mark.parametrize("n_atoms", [1000, 996, 64])
def test_tma_staged_hoptraj_equals_register_staged(K, n_atoms, monkeypatch):
    """Stage 3 reads the cached brick indices through TMA boxes of 32 atoms x 8 frames (full and partial last warp, chunks
    that are not multiples of 8 frames) or, with HOPB_K3_TMA=0 / a row stride that is no multiple of 16 bytes, through
    registers: same planes, same summaries, same events."""
    p = SynthParams(n_atoms=n_atoms, n_frames=203)
    xyz = generate(p)
    ref = O.pipeline(xyz, delta=1.0)
    g = K.GridTables(ref["edges"])
    counts = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    cells = torch.empty(xyz.shape[:2], dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, dev(xyz), counts, cells)
    brick = K.brick_map(dev(ref["map"]))
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("HOPB_K3_TMA", mode)
        for dt in (torch.int16, torch.float32):
            for n_chunks, chunk_len in ((1, 203), (4, 51), (7, 29)):
                ev = K.EventBuffer(xyz.shape[0] * xyz.shape[1], "cuda")
                X = torch.empty(xyz.shape[:2], dtype=dt, device="cuda")
                Y = torch.empty(xyz.shape[:2], dtype=dt, device="cuda")
                s = K.hoptraj(g, None, brick, 0, n_chunks, chunk_len, ev, X, Y, cells=cells)
                st = K.new_state(xyz.shape[1], "cuda")
                K.hop_advance(None, st, init=True)
                K.hop_fixup(s, chunk_len, 203, st, Y, ev)
                e = ev.data[: ev.n()].cpu().numpy()
                e = e[np.lexsort((e[:, 1], e[:, 0]))]
                out[(mode, dt, n_chunks)] = (X.cpu().numpy().astype(np.float32), Y.cpu().numpy().astype(np.float32), s.cpu().numpy(), e)
                assert np.array_equal(out[(mode, dt, n_chunks)][0], ref["hop"][:, :, 0])
                assert np.array_equal(out[(mode, dt, n_chunks)][1], ref["hop"][:, :, 1])
    for key, val in out.items():
        other = out[("0",) + key[1:]]
        assert all(np.array_equal(a, b) for a, b in zip(val, other)), key
