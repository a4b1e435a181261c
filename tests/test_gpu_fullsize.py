"""Size-independent properties at the full sizes of BASELINE.json's configurations plus an oracle
comparison on a frame prefix.  One GPU holds

  config 2   30 000 waters x 10 000 frames, 0.5 A grid (201^3)           -- the bench workload
  config 3   100 000 waters x 12 500 frames (the slab one of 8 GPUs owns of 100 000), 297^3:
             the count grid (105 MB) is histogrammed in two L2-sized slabs (hist_cells_kernel)
  config 4   1 000 ions of one species in a 126 A box x 1 000 000 frames, 1.0 A grid (130^3):
             147 frame chunks, no bulk site
  config 5   one of the 64 trajectories, 30 000 waters x 20 000 frames, 0.25 A grid (402^3):
             7-slab histogram, block maps too large for shared memory (MAP_CELL2 gather path)

The oracle cannot run 1e8..1e9 atom-frames, so parity at these sizes rests on (1) properties that must
hold for any input -- conservation of counts, independence of the result from how the frames are cut
into chunks / slabs, agreement of the two input paths (coordinates vs cached brick indices),
idempotence of the event ordering, a third-opinion labelling (scipy.ndimage with the reference's
15-cell structuring element, hop/sitemap.py:415-422) -- and (2) an exact comparison of the first 64
frames against the oracle, using the full-size site map (events are causal: the events of a prefix do
not depend on later frames)."""
import numpy as np
import pytest

from oracle import hop_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

PREFIX = 64

CONFIGS = {
    2: dict(atoms=30000, frames=10000, delta=0.5, rho=None, mem=40e9),
    3: dict(atoms=100000, frames=12500, delta=0.5, rho=None, mem=100e9),
    4: dict(atoms=1000, frames=1000000, delta=1.0, rho=1000 / 126.0 ** 3, mem=60e9),
    5: dict(atoms=30000, frames=20000, delta=0.25, rho=None, mem=60e9),
}


class Full(object):
    pass


@pytest.fixture(scope="module", params=sorted(CONFIGS), ids=lambda c: "config%d" % c)
def full(request):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cid = request.param
    c = CONFIGS[cid]
    if torch.cuda.get_device_properties(0).total_memory < c["mem"]:
        pytest.skip("needs a large-memory GPU")
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    kw = {} if c["rho"] is None else {"rho": c["rho"]}
    p = SynthParams(n_atoms=c["atoms"], n_frames=c["frames"], seed=20261019 + (cid if cid != 2 else 0), **kw)
    f = Full()
    f.cid, f.K, f.atoms, f.frames, f.delta = cid, K, c["atoms"], c["frames"], c["delta"]
    f.xyz = K.synth_generate(p)
    f.host_prefix = generate(p, 0, PREFIX)
    assert np.array_equal(f.xyz[:PREFIX].cpu().numpy(), f.host_prefix)      # device generator == host generator
    f.edges = grid_edges_from_frame0(f.host_prefix[0], delta=c["delta"])
    f.pipe = HopPipeline(f.edges, c["frames"], PipelineConfig())
    f.res = f.pipe.run(f.xyz)
    yield f
    del f.res, f.pipe, f.xyz
    torch.cuda.empty_cache()


def test_grid_is_the_configured_one(full):
    want = {2: 201, 3: 297, 4: 130, 5: 402}[full.cid]
    assert tuple(full.pipe.grid.shape) == (want, want, want)


def test_counts_are_conserved(full):
    res = full.res
    # every synthetic atom stays inside the padded grid, so every atom-frame is counted exactly once
    assert int(res.counts.to(torch.int64).sum().item()) == full.atoms * full.frames
    assert int(res.counts.min().item()) >= 0


def test_prefix_matches_oracle(full):
    K, res, host_prefix, edges = full.K, full.res, full.host_prefix, full.edges
    site_map = res.map16.cpu().numpy()
    hop = O.coord2hop_vectorised(host_prefix, edges, site_map)
    assert np.array_equal(res.hop_x[:PREFIX].cpu().numpy(), hop[:, :, 0])
    assert np.array_equal(res.hop_y[:PREFIX].cpu().numpy(), hop[:, :, 1])
    ev_ref, _ = O.analyze_hops_vectorised(hop[:, :, 0])
    ev = K.events_to_numpy(res.events)
    sel = ev[ev["frame"] < PREFIX]
    order = np.lexsort((sel["iatom"], sel["frame"]))
    got = np.stack([sel[k][order] for k in ("frame", "iatom", "s0", "s", "tau", "tbarrier")], axis=1).astype(np.int64)
    assert np.array_equal(got, ev_ref)
    # prefix histogram (a grid of its own: the slab budget applies as in the full run)
    counts_ref = O.histogram_vectorised(host_prefix, edges)
    g = K.GridTables(edges)
    c = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, full.xyz[:PREFIX], c)
    assert np.array_equal(c.cpu().numpy().astype(np.int64), counts_ref)
    # and with the cached brick indices, i.e. through the slabbed passes where the grid outgrows L2
    c.zero_()
    cells = torch.empty((PREFIX, full.atoms), dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, full.xyz[:PREFIX], c, cells)
    assert np.array_equal(c.cpu().numpy().astype(np.int64), counts_ref)


def test_density_is_the_reference_arithmetic(full):
    """density == counts / n_frames / dx / dy / dz * factor in the reference's operation order
    (hop/density.py:133-137, hop/sitemap.py:196-216, 236-264), bit for bit, on an 8-plane slice."""
    res = full.res
    sl = slice(full.pipe.grid.shape[0] // 2, full.pipe.grid.shape[0] // 2 + 8)
    counts = res.counts[sl].cpu().numpy().astype(np.int64)
    got = res.density[sl].cpu().numpy()
    grid = counts.astype(np.float64)
    grid /= float(full.frames)
    e = full.edges
    dedges = [np.diff(x) for x in e]
    shape = [1, 1, 1]
    for d in range(3):
        s = list(shape)
        w = dedges[d][sl] if d == 0 else dedges[d]
        s[d] = len(w)
        grid /= w.reshape(s)
    from hop_b200.constants import density_conversion_factor

    grid *= density_conversion_factor("Angstrom^{-3}", full.pipe.cfg.density_unit)
    assert np.array_equal(got, grid)


def test_site_map_is_a_valid_labelling(full):
    pipe, res = full.pipe, full.res
    m = res.map16.cpu().numpy()
    dens = res.density.cpu().numpy()
    own = res.own.cpu().numpy()
    assert m.min() == 0 and m.max() == res.n_labels - 1
    assert np.array_equal(own > 0, dens >= pipe.cfg.solvent_threshold)       # hydration sites = thresholded density
    assert np.array_equal(m >= 2, own > 0)
    assert np.all(dens[m == 1] >= pipe.cfg.bulk_threshold)
    vol = np.bincount(own.ravel())[2:]
    assert np.all(vol[:-1] >= vol[1:])                                        # labels ordered by decreasing volume
    # same partition as scipy on the real full-size density (third opinion, vectorised)
    from scipy import ndimage

    st = np.zeros((3, 3, 3), bool)
    st[1, 1, 1] = True
    for o in O.DELTA_FIRST_OCTANT:
        st[tuple(1 + o)] = st[tuple(1 - o)] = True
    lab, n = ndimage.label(dens >= pipe.cfg.solvent_threshold, structure=st)
    assert n == own.max() - 1
    pairs = np.unique(np.stack([lab[own > 0], own[own > 0]], axis=1), axis=0)
    assert len(pairs) == n                                                    # one-to-one between the two labellings
    # canonical numbering of equal volumes: larger minimum C-order cell first
    idx = np.nonzero(own.ravel() > 0)[0]
    first = np.full(own.max() + 1, -1, np.int64)
    first[own.ravel()[idx][::-1]] = idx[::-1]                                 # reversed: the smallest index is written last
    tie = vol[:-1] == vol[1:]
    assert np.all(first[2:][:-1][tie] > first[2:][1:][tie])
    # the bulk site is the largest component of the bulk-threshold mask
    member = res.member.cpu().numpy().astype(bool)
    labb, nb = ndimage.label(dens >= pipe.cfg.bulk_threshold, structure=st)
    if nb:
        sizes = np.bincount(labb.ravel())[1:]
        assert member.sum() == sizes.max()
        assert len(np.unique(labb[member])) == 1
    else:
        assert not member.any()
    del lab, labb


def test_result_does_not_depend_on_chunking_or_input_path(full):
    K, pipe, xyz, res = full.K, full.pipe, full.xyz, full.res
    ATOMS, FRAMES = full.atoms, full.frames
    ref_x, ref_y = res.hop_x, res.hop_y
    ref_ev = K.events_to_numpy(res.events)
    for n_chunks, use_cells, block_map in [(1, True, "coarse"), (7, False, "none"), (23, True, "auto"), (5, False, "cell2")]:
        ev = K.EventBuffer(max(len(ref_ev) + 16, 1 << 20), "cuda")
        X = torch.empty_like(ref_x)
        Y = torch.empty_like(ref_y)
        chunk_len = -(-FRAMES // n_chunks)
        nc = -(-FRAMES // chunk_len)
        s = K.hoptraj(pipe.grid, xyz, pipe.brick, 0, nc, chunk_len, ev, X, Y,
                      cells=pipe.cells if use_cells else None, block_map=block_map)
        st = K.new_state(ATOMS, "cuda")
        K.hop_advance(None, st, init=True)
        K.hop_fixup(s, chunk_len, FRAMES, st, Y, ev)
        assert torch.equal(X, ref_x) and torch.equal(Y, ref_y)
        n_ev = ev.n()
        assert n_ev == len(ref_ev)
        out = K.events_finalize(ev, n_ev, ATOMS, FRAMES, res.n_labels)
        assert np.array_equal(K.events_to_numpy(out[0]), ref_ev)              # same events in the same order
        a, b = st.cpu().numpy(), res.final_state.cpu().numpy()
        assert np.array_equal(a[[0, 1, 3]], b[[0, 1, 3]])                     # s0, t0, on
        off = a[3] == 0
        assert np.array_equal(a[2][off], b[2][off])                           # t1 only matters while off the site
        del X, Y, ev, s


def test_events_are_consistent_with_the_label_stream(full):
    K, res = full.K, full.res
    ATOMS = full.atoms
    ev = K.events_to_numpy(res.events)
    x = res.hop_x
    # at the event frame the atom is on site s; on the frame before it was not
    idx = torch.from_numpy(ev["frame"].astype(np.int64)).cuda() * ATOMS + torch.from_numpy(ev["iatom"].astype(np.int64)).cuda()
    assert torch.equal(x.view(-1)[idx].to(torch.int32), torch.from_numpy(ev["s"]).cuda())
    assert bool((x.view(-1)[idx - ATOMS].to(torch.int32) != torch.from_numpy(ev["s"]).cuda()).all())
    assert np.all(ev["tau"] >= 0) and np.all(ev["tbarrier"] >= 0) and np.all(ev["s0"] != ev["s"]) and np.all(ev["s"] > 0)
    # sorted by (edge, frame, atom), edges strictly increasing, counts add up
    key = (ev["s0"].astype(np.int64) + 1) * (res.n_labels + 1) + ev["s"] + 1
    t = ev["frame"].astype(np.int64) * ATOMS + ev["iatom"]
    assert np.all((key[1:] > key[:-1]) | ((key[1:] == key[:-1]) & (t[1:] > t[:-1])))
    assert int(res.edge_count.sum().item()) == len(ev) == int(res.count_matrix.sum().item())
    # orbit column: equals x where x is a site, never -1 (frame slices bound the temporaries)
    y = res.hop_y
    step = max(1, (1 << 28) // ATOMS)
    for f0 in range(0, full.frames, step):
        xs, ys = x[f0:f0 + step], y[f0:f0 + step]
        assert bool(((ys == xs) | (xs <= 0)).all()) and bool((ys >= 0).all())
