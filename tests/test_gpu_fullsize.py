"""Size-independent properties at BASELINE.json's full single-GPU size (configs[1]: 30 000 waters x
10 000 frames, 0.5 A grid = 3e8 atom-frames) plus an oracle comparison on a frame prefix.

The oracle cannot run 3e8 atom-frames, so parity at this size rests on (1) properties that must hold
for any input -- conservation of counts, independence of the result from how the frames are cut into
chunks / slabs, agreement of the two input paths (coordinates vs cached brick indices), idempotence
of the event ordering -- and (2) an exact comparison of the first 64 frames against the oracle, using
the full-size site map (events are causal: the events of a prefix do not depend on later frames)."""
import numpy as np
import pytest

from oracle import hop_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ATOMS, FRAMES, DELTA = 30000, 10000, 0.5
PREFIX = 64


@pytest.fixture(scope="module")
def full():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a large-memory GPU")
    from hop_b200 import kernels as K
    from hop_b200.pipeline import HopPipeline, PipelineConfig, grid_edges_from_frame0
    from hop_b200.synth import SynthParams, generate

    p = SynthParams(n_atoms=ATOMS, n_frames=FRAMES)
    xyz = K.synth_generate(p)
    host_prefix = generate(p, 0, PREFIX)
    assert np.array_equal(xyz[:PREFIX].cpu().numpy(), host_prefix)      # device generator == host generator
    edges = grid_edges_from_frame0(host_prefix[0], delta=DELTA)
    pipe = HopPipeline(edges, FRAMES, PipelineConfig())
    res = pipe.run(xyz)
    return K, pipe, xyz, edges, res, host_prefix


def test_counts_are_conserved(full):
    K, pipe, xyz, edges, res, _ = full
    # every synthetic atom stays inside the padded grid, so every atom-frame is counted exactly once
    assert int(res.counts.to(torch.int64).sum().item()) == ATOMS * FRAMES
    assert int(res.counts.min().item()) >= 0


def test_prefix_matches_oracle(full):
    K, pipe, xyz, edges, res, host_prefix = full
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
    # prefix histogram
    counts_ref = O.histogram_vectorised(host_prefix, edges)
    g = K.GridTables(edges)
    c = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    K.hist_accumulate(g, xyz[:PREFIX], c)
    assert np.array_equal(c.cpu().numpy().astype(np.int64), counts_ref)


def test_site_map_is_a_valid_labelling(full):
    K, pipe, xyz, edges, res, _ = full
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


def test_result_does_not_depend_on_chunking_or_input_path(full):
    K, pipe, xyz, edges, res, _ = full
    ref_x, ref_y = res.hop_x, res.hop_y
    ref_ev = K.events_to_numpy(res.events)
    for n_chunks, use_cells, block_map in [(1, True, "coarse"), (7, False, "none"), (23, True, "fine"), (5, False, "fine")]:
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


def test_events_are_consistent_with_the_label_stream(full):
    K, pipe, xyz, edges, res, _ = full
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
    # orbit column: equals x where x is a site, never 0 after the first site visit, never -1
    y = res.hop_y
    on = x > 0
    assert bool((y[on] == x[on]).all()) and bool((y >= 0).all())
