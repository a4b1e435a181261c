"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
the Universe stand-in, selections, DCD/PSF/DX formats, grid geometry, chunk planning."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import hop_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_in_header():
    from hop_b200 import _lib

    header = open(os.path.join(ROOT, "include", "hopb200.h")).read()
    declared = set(re.findall(r"^(?:int|void|uint64_t|const char\*)\s+(hopb_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations found in include/hopb200.h"
    if not os.path.exists(_lib.LIB_PATH):       # fresh checkout: the library is a build product (nvcc needs no GPU)
        from hop_b200.csrc.build import build

        build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libhopb200.so does not export %s" % name
    assert declared == set(_lib.exported_symbols())
    assert _lib.lib().hopb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch

    from hop_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        _lib.handle()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hop_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "tests.emul" not in src, f


def test_fixedwidth_bins_and_edges_match_histogramdd():
    from hop_b200.pipeline import grid_edges_from_frame0
    from hop_b200.utilities import fixedwidth_bins

    rng = np.random.default_rng(0)
    c0 = (rng.random((500, 3)) * [31.3, 17.9, 44.4] - 3.3).astype(np.float32)
    for delta in (1.0, 0.5, 0.25, 0.7):
        bins, arange, edges = O.grid_geometry(c0, delta=delta, padding=2.0)
        got = grid_edges_from_frame0(c0, delta=delta, padding=2.0)
        for a, b in zip(edges, got):
            assert np.array_equal(a, b)
        B = fixedwidth_bins(delta, c0.min(0) - 2.0, c0.max(0) + 2.0)
        assert np.array_equal(B["Nbins"], bins)
    with pytest.raises(ValueError):
        fixedwidth_bins(1.0, np.array([1.0, 0, 0]), np.array([0.0, 1, 1]))


def test_universe_selections_and_iteration():
    from hop_b200.universe import Universe

    xyz = np.arange(5 * 6 * 3, dtype=np.float32).reshape(5, 6, 3)
    names = ["OH2", "H1", "H2", "OH2", "CA", "NA"]
    resn = ["TIP3", "TIP3", "TIP3", "TIP3", "ALA", "ION"]
    u = Universe(coordinates=xyz, names=names, resnames=resn, dt=2.0)
    assert u.trajectory.n_frames == 5 and u.trajectory.totaltime == 8.0
    w = u.select_atoms("name OH2")
    assert list(w.indices) == [0, 3]
    assert list(u.select_atoms("name H*").indices) == [1, 2]
    assert list(u.select_atoms("protein and not name H*").indices) == [4]
    assert list(u.select_atoms("resname TIP3 and not (name OH2 or name H1)").indices) == [2]
    assert list(u.select_atoms("all").indices) == list(range(6))
    assert list(u.select_atoms("index 1:2").indices) == [1, 2]
    frames = [ts.frame for ts in u.trajectory]
    assert frames == [0, 1, 2, 3, 4]
    for ts in u.trajectory[1:4:2]:
        assert np.array_equal(w.positions, xyz[ts.frame][[0, 3]])
    u.trajectory.rewind()
    assert u.trajectory.ts.frame == 0
    assert u.trajectory.next().frame == 1
    from hop_b200.exceptions import SelectionError

    with pytest.raises(SelectionError):
        u.select_atoms("name")


def test_dcd_psf_roundtrip(tmp_path):
    from hop_b200.dcd import DCDFile, DCDWriter, read_psf_atoms, write_dcd
    from hop_b200.universe import Universe

    rng = np.random.default_rng(1)
    x = rng.integers(-1, 40, size=(7, 11)).astype(np.float32)
    y = rng.integers(0, 40, size=(7, 11)).astype(np.float32)
    fn = str(tmp_path / "hop.dcd")
    with DCDWriter(fn, 11, dt=2.5, remarks="Hopping trajectory: x=site y=orbit_site z=0") as w:
        for f in range(7):
            w.write_planes(x[f], y[f], None, [41, 41, 1, 90, 90, 90])
    d = DCDFile(fn)
    assert d.n_frames == 7 and d.n_atoms == 11
    assert abs(d.dt - 2.5) < 1e-5
    assert d.remarks[0].startswith("Hopping trajectory")
    assert np.array_equal(d.plane_array(0), x) and np.array_equal(d.plane_array(1), y)
    assert np.all(d.plane_array(2) == 0)
    assert np.allclose(d.dimensions(3), [41, 41, 1, 90, 90, 90])
    # byte layout: 84-byte header record, CORD magic, per-frame records X,Y,Z of 4N bytes
    raw = open(fn, "rb").read()
    assert raw[:8] == b"T\x00\x00\x00CORD"
    # full coordinates round trip through the Universe reader
    xyz = rng.random((4, 5, 3)).astype(np.float32)
    fn2 = str(tmp_path / "c.dcd")
    write_dcd(fn2, xyz, dt=1.0)
    psf = str(tmp_path / "c.psf")
    with open(psf, "w") as f:
        f.write("PSF EXT\n\n      1 !NTITLE\n* test\n\n         5 !NATOM\n")
        for i in range(5):
            f.write("%10d %8s %-8d %8s %-8s %4s %-14.6f%-14.4f%8d\n" % (i + 1, "W", i + 1, "TIP3", "OH2", "OT", -0.834, 15.9994, 0))
    top = read_psf_atoms(psf)
    assert list(top["name"]) == ["OH2"] * 5
    u = Universe(psf, fn2)
    got = np.stack([np.array(u.atoms.positions) for ts in u.trajectory])
    assert np.array_equal(got, xyz)
    assert np.array_equal(u.trajectory.coordinate_array(1, 3), xyz[1:3])


def test_dx_roundtrip(tmp_path):
    from hop_b200.dx import read_dx, write_dx

    rng = np.random.default_rng(2)
    g = rng.random((4, 5, 7))
    fn = str(tmp_path / "d.dx")
    write_dx(fn, g, [0.5, 1.0, 1.5], [1.0, 0.5, 0.25], ["hello"])
    g2, edges = read_dx(fn)
    np.testing.assert_allclose(g2, g, rtol=1e-8)
    assert len(edges[2]) == 8 and abs(edges[2][1] - edges[2][0] - 0.25) < 1e-12


def test_plan_chunks_tiles_exactly():
    from hop_b200.kernels import plan_chunks

    for F, N in [(200, 1000), (10000, 30000), (100000, 100000), (1, 5), (31, 7), (1000000, 300)]:
        n, l = plan_chunks(F, N)
        assert n >= 1 and l >= 1 and n * l >= F and (n - 1) * l < F and n <= 65535


def test_unit_factors_and_labels():
    from hop_b200 import constants as C

    assert C.SITELABEL == dict(outlier=-1, interstitial=0, bulk=1)
    assert C.get_conversion_factor("density", "Angstrom^{-3}", "water") == O.density_conversion_factor("Angstrom^{-3}", "water")
    assert abs(C.get_conversion_factor("density", "Angstrom^{-3}", "water") * 0.0333679 - 1.0) < 2e-3


def test_synth_generator_properties():
    from hop_b200.synth import SynthParams, generate

    p = SynthParams(n_atoms=600, n_frames=30)
    xyz = generate(p)
    assert xyz.dtype == np.float32 and xyz.shape == (30, 600, 3)
    assert xyz.min() >= 0.0 and xyz.max() < p.box
    assert np.array_equal(generate(p, 10, 20), xyz[10:20])
    assert np.allclose(xyz[0, 1], 0.0) and xyz[0, 2].min() > p.box * 0.999


def test_survivalfunction_and_rates_match_oracle():
    """hop_b200.graph.HoppingGraph rates vs the oracle's restatement of hop/graph.py:1677-1748
    (tolerance 1e-6 relative, north_star); the survival function must be identical."""
    import warnings

    import networkx as NX

    from hop_b200.graph import HoppingGraph, Unitstep, fitExp2, survivalfunction

    rng = np.random.default_rng(7)
    cases = {
        (2, 3): 2.0 * np.round(rng.exponential(30, 400)),                       # single population
        (3, 2): np.concatenate([np.round(rng.exponential(4, 300)), np.round(rng.exponential(90, 300))]) + 1.0,
        (2, 4): np.array([3.0, 1.0, 8.0, 2.0]),                                  # <= 10 events -> single exponential
        (4, 2): np.full(25, 5.0),                                                # all equal: step at t = 5
    }
    for taus in cases.values():
        x = np.arange(0, taus.max() + 2, 1.0)
        S = survivalfunction(taus)
        assert np.array_equal(S(x), O.survival_faithful(taus, x))
        assert np.array_equal(S(x + 0.5), O.survival_faithful(taus, x + 0.5))
    assert Unitstep(1.0, 1.0) == 0.5 and Unitstep(2.0, 1.0) == 1.0 and Unitstep(0.0, 1.0) == 0.0
    g = NX.DiGraph()
    g.add_edges_from(cases.keys())
    props = {e: {"tau": t, "tbarrier": 0 * t, "frame": t, "iatom": t} for e, t in cases.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hg = HoppingGraph(g, props, trjdata={"dt": 1.0}, rate_method="minpack")   # the host path (no GPU in this suite)
        for e, taus in cases.items():
            k, n, p = O.rate_faithful(taus)
            d = hg.graph[e[0]][e[1]]
            assert d["N"] == n == hg.number_of_hops(e)
            assert abs(d["k"] - k) <= 1e-6 * abs(k), (e, d["k"], k)
            assert np.allclose(d["fit"].parameters, p, rtol=1e-6, atol=0)
            f = hg.waitingtime_fit(e)
            kk = f.parameters[0] if len(f.parameters) == 1 else (f.parameters[1] if f.parameters[0] > 0.5 else f.parameters[2])
            assert hg.rate(e) == 1000 * kk
    assert len(hg.waitingtime_fit((2, 4)).parameters) == 1
    assert isinstance(hg.waitingtime_fit((3, 2)), fitExp2)
    assert set(hg.rates(2)) == set(cases)


def test_rate_fits_in_worker_processes_are_identical():
    """The fits of a graph are spread over fresh worker interpreters (hop_b200/_ratefit.py); the results
    must be the very same numbers as in-process fits, in the same order."""
    import warnings

    from hop_b200 import _ratefit

    rng = np.random.default_rng(11)
    taus = [np.round(rng.exponential(15, int(n))) + 1.0 for n in rng.integers(2, 120, 23)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        serial = _ratefit.fit_all(taus, workers=0)
        pooled = _ratefit.fit_all(taus, workers=3)
    assert len(serial) == len(pooled) == len(taus)
    for a, b in zip(serial, pooled):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[3] == b[3] and a[4] == b[4]
    assert _ratefit.default_workers(10) == 0


def test_pipeline_config_plane_dtype_and_exchange_fallbacks():
    """Host-side switches of the later part of round 2: the plane type of the pipeline is int16 by default and checked;
    without a process group nothing asks for symmetric memory and the exchanges are no-ops on one rank."""
    torch = pytest.importorskip("torch")
    from hop_b200 import exchange
    from hop_b200.pipeline import PipelineConfig

    assert PipelineConfig().plane_torch_dtype() == torch.int16
    assert PipelineConfig(plane_dtype="float32").plane_torch_dtype() == torch.float32
    with pytest.raises(ValueError):
        PipelineConfig(plane_dtype="int8").plane_torch_dtype()
    assert exchange.dist() is None and not exchange.nvls_wanted()
    assert exchange.symm_entry(None) is None and exchange.symm_entry(torch.zeros(4)) is None
    t = torch.arange(6, dtype=torch.int32)
    assert exchange.reduce_counts(t) is t and exchange.reduce_count_matrix(t) is t
    assert tuple(exchange.gather_slab_summaries(t.view(2, 3)).shape) == (1, 2, 3)
    assert exchange.slab_bounds(10, 4, 0) == (0, 3) and exchange.slab_bounds(10, 4, 3) == (8, 10)


def test_packed16_sum_check_detects_every_wrap():
    """The invariant behind hist_flush16_kernel (csrc/hist.cu), replayed with numpy's wrapping uint32 arithmetic: adding
    1 << 16 * (cell & 1) into word cell >> 1 per hit, the sum of all decoded 16-bit counters equals the number of hits
    exactly when no counter passed 65535 -- every wrap takes 65535 (low half: the carry lands in the high half) or 65536
    (high half: the carry is lost) out of it."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        n_cells = int(rng.integers(2, 40))
        hits = rng.integers(0, 3000, n_cells).astype(np.int64)
        if trial % 2:                                     # make some counters wrap, in low and in high halves
            for c in rng.choice(n_cells, size=int(rng.integers(1, 4)), replace=False):
                hits[c] = int(rng.integers(65536, 200000))
        words = np.zeros((n_cells + 1) // 2, np.uint32)
        for c, h in enumerate(hits):
            words[c >> 1] += np.uint32((int(h) << (16 * (c & 1))) & 0xFFFFFFFF)      # h reductions of 1 << shift, mod 2^32
        decoded_sum = int((words & 0xFFFF).astype(np.int64).sum() + (words >> 16).astype(np.int64).sum())
        wrapped = bool((hits > 65535).any())
        assert (decoded_sum == int(hits.sum())) == (not wrapped), (trial, hits)
        if not wrapped:
            dec = np.stack([words & 0xFFFF, words >> 16], axis=1).ravel()[:n_cells]
            assert np.array_equal(dec.astype(np.int64), hits)
