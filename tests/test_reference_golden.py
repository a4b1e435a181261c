"""Parity against outputs of the reference ITSELF.

tests/golden/reference/*.npz were produced by hop's own classes (DensityCreator, Density.map_sites,
site_insert_bulk, HoppingTrajectory.write, TransportNetwork.compute_residency_times, HoppingGraph)
executed from /root/reference by tests/golden/reference_py3.py + make_reference_golden.py; see those
files for how Python-2 hop is run under Python 3.  Here

* the oracle is held to them on the CPU (this pins oracle/hop_oracle.py to the reference), and
* the CUDA path is held to them through the drop-in classes (`-m gpu`).

Integer products are compared bit-exactly.  The only normalisation is the numbering of hydration
sites of EQUAL volume, which the reference leaves to set-iteration order (hop/sitemap.py:1560-1568);
both sides are brought to the canonical order (volume descending, larger minimum C-order cell first)
by `canonical_perm`.  Densities are compared bit-exactly, site properties and rates to 1e-6 relative
(BASELINE.json north_star).
"""
import math
import os
import sys

import numpy as np
import pytest

from oracle import hop_oracle as O

from tests.refgolden import FILES, HERE, NAMES, canonical_perm, load, reference_products, relabel  # noqa: F401


def check_transitions(gp, ref, dt):
    """graph_properties (edge tuple or node int -> arrays) against the reference's records."""
    edges = {k: v for k, v in gp.items() if isinstance(k, tuple)}
    nodes = {k: v for k, v in gp.items() if not isinstance(k, tuple)}
    assert set(edges) == set(ref["props"]) and set(nodes) == set(ref["nprops"])
    for k, d in ref["props"].items():
        assert np.array_equal(edges[k]["frame"], d["frame"]) and np.array_equal(edges[k]["iatom"], d["iatom"]), k
        assert np.allclose(edges[k]["tau"], d["tau"], rtol=1e-12, atol=0) and \
            np.allclose(edges[k]["tbarrier"], d["tbarrier"], rtol=1e-12, atol=0), k
    for k, d in ref["nprops"].items():
        assert np.array_equal(nodes[k]["iatom"], d["iatom"]) and np.array_equal(nodes[k]["frame"], d["frame"])
        assert np.allclose(np.asarray(nodes[k]["tau"], dtype=float), d["tau"], rtol=1e-12, atol=0)
        assert all(v is None for v in nodes[k]["tbarrier"])


def check_site_properties(volume, occ_avg, occ_std, center, ref):
    assert np.allclose(volume, ref["volume"], rtol=1e-9, atol=0)
    assert np.allclose(occ_avg, ref["occ_avg"], rtol=1e-6, atol=1e-12)
    assert np.allclose(occ_std, ref["occ_std"], rtol=1e-6, atol=1e-9)
    has = ref["ncells"] > 0
    assert np.allclose(np.asarray(center)[has], ref["center"][has], rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------------------------------------
# CPU: the oracle against the reference
# ------------------------------------------------------------------------------------------------
def test_reference_golden_files_exist():
    assert set(NAMES) >= {"ref_clean", "ref_outliers", "ref_half", "ref_solute", "ref_config1"}


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("faithful", [False, True])
def test_oracle_matches_reference_run(name, faithful):
    g = load(name)
    if faithful and name == "ref_half":
        pytest.skip("the loop restatement is exercised on the three 1.0 A cases")
    n_solv = int(g["n_solvent"])
    xyz, solute = g["xyz"][:, :n_solv], g["xyz"][:, n_solv:]
    unit, dt = str(g["density_unit"]), float(g["dt"])
    ref = reference_products(g)
    F = xyz.shape[0]
    bins, arange, edges = O.grid_geometry(xyz[0], float(g["delta"]), 2.0)
    for d in range(3):
        assert np.array_equal(edges[d], g["edges_" + "xyz"[d]])
    counts = O.histogram_faithful(iter(xyz), bins, arange)[0] if faithful else O.histogram_vectorised(xyz, edges)
    grid = O.make_density(counts, edges, F, unit)
    assert np.array_equal(grid, g["grid"])                      # bit-exact density
    if solute.shape[1]:
        bulk_counts = O.histogram_bulk(xyz, solute, edges, float(g["cutoff"]))
        bulk_grid = O.make_density(bulk_counts, edges, F, unit)
        assert np.array_equal(O.make_density(bulk_counts, edges, F, "Angstrom^{-3}"), g["bulk_grid_hist"])
    else:
        bulk_grid = grid
    label = O.label_sites_faithful if faithful else O.label_sites_ndimage
    sites = label(grid, float(g["solvent_threshold"]), 1)
    bulk_sites = label(bulk_grid, float(g["bulk_threshold"]), 1)
    assert np.array_equal(O.draw_map_from_sites(bulk_sites, grid.shape) == 1, g["bulk_map"] == 1)
    assert len(bulk_sites[1]) == int((g["bulk_map"] == 1).sum())
    sites = O.site_insert_bulk(sites, bulk_sites)
    site_map = O.draw_map_from_sites(sites, grid.shape)
    assert site_map.dtype == np.int16
    assert np.array_equal(site_map, ref["map"])                 # oracle ties are canonical already
    hop = (O.coord2hop_faithful(iter(xyz), edges, site_map) if faithful
           else O.coord2hop_vectorised(xyz, edges, site_map))
    assert hop.dtype == np.float32 and np.all(hop[:, :, 2] == 0)
    assert np.array_equal(hop[:, :, 0], ref["hop_x"]) and np.array_equal(hop[:, :, 1], ref["hop_y"])
    props, e, n = O.analyze_hops_faithful(hop[:, :, 0], dt=dt)
    check_transitions(props, ref, dt)
    ev, _ = O.analyze_hops_vectorised(hop[:, :, 0])
    pairs, cnt = O.events_to_counts(ev)
    assert {tuple(p): c for p, c in zip(pairs.tolist(), cnt.tolist())} == \
        {k: len(d["tau"]) for k, d in ref["props"].items()}
    assert set(map(tuple, O.graph_alltransitions(hop[:, :, 0]).tolist())) == ref["alltransitions"]
    sp = O.site_properties(sites, grid, edges, unit)
    check_site_properties(sp["volume"], sp["occupancy_avg"], sp["occupancy_std"], sp["center"], ref)
    if not faithful:                                            # rate constants of the busiest edges
        busiest = sorted(ref["rates"], key=lambda k: -ref["rates"][k][1])[:12]
        for k in busiest:
            k_ref, n_ref = ref["rates"][k]
            k_or, n_or, _ = O.rate_faithful(props[k]["tau"])
            assert n_or == n_ref and abs(k_or - k_ref) <= 1e-6 * abs(k_ref), (k, k_or, k_ref)


@pytest.mark.skipif(not os.path.isdir("/root/reference/hop"), reason="reference sources not present")
def test_reference_run_is_reproducible():
    """Re-runs hop on the first case and compares with the committed file (only where /root/reference
    exists, i.e. in the build container -- never on the GPU box)."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_reference_golden as M

    g = load("ref_clean")
    try:
        res = _rerun(M, g)
    finally:
        M.R.unload_reference()              # do not leave the reference or the fakes importable
    assert np.array_equal(res["grid"], g["grid"]) and np.array_equal(res["map"], g["map"])
    assert np.array_equal(res["hop"][:, :, 0], g["hop_x"]) and np.array_equal(res["events"], g["events"])


def _rerun(M, g):
    return M.run_reference(g["xyz"], int(g["n_solvent"]), delta=float(g["delta"]),
                          density_unit=str(g["density_unit"]), solvent_threshold=float(g["solvent_threshold"]),
                          bulk_threshold=float(g["bulk_threshold"]), rates=False)


# ------------------------------------------------------------------------------------------------
# GPU: the drop-in classes (CUDA path through the C ABI) against the reference
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_reference_run(name, tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hop_b200.density import DensityCreator
    from hop_b200.graph import TransportNetwork
    from hop_b200.sitemap import Density
    from hop_b200.trajectory import HoppingTrajectory
    from hop_b200.universe import Universe

    g = load(name)
    n_solv = int(g["n_solvent"])
    n_total = g["xyz"].shape[1]
    unit, dt = str(g["density_unit"]), float(g["dt"])
    ref = reference_products(g)
    names = ["OH2"] * n_solv + ["CA"] * (n_total - n_solv)
    resn = ["TIP3"] * n_solv + ["ALA"] * (n_total - n_solv)
    u = Universe(coordinates=g["xyz"], names=names, resnames=resn, dt=dt, filename="mem.psf")
    if n_total > n_solv:
        DC = DensityCreator(u, mode="all", delta=float(g["delta"]), atomselection="name OH2",
                            soluteselection="protein", cutoff=float(g["cutoff"]), padding=2.0)
        DC.create()
        assert np.array_equal(DC.densities["bulk"].grid, g["bulk_grid_hist"])
        dens = DC.DensityWithBulk(density_unit=unit, solvent_threshold=float(g["solvent_threshold"]),
                                  bulk_threshold=float(g["bulk_threshold"]))
        bulk = DC.densities["bulk"]
    else:
        DC = DensityCreator(u, mode="solvent", delta=float(g["delta"]), atomselection="name OH2", padding=2.0)
        dens = DC.create()["solvent"]
        bulk = Density(grid=np.array(dens.grid), edges=[np.array(e) for e in dens.edges],
                       parameters={"isDensity": True}, unit={"length": "Angstrom", "density": "Angstrom^{-3}"})
        dens.convert_density(unit)
        dens.map_sites(float(g["solvent_threshold"]))
        bulk.convert_density(unit)
        bulk.map_sites(float(g["bulk_threshold"]))
        dens.site_insert_bulk(bulk)
    for d in range(3):
        assert np.array_equal(dens.edges[d], g["edges_" + "xyz"[d]])
    assert np.array_equal(dens.grid, g["grid"])                 # bit-exact density
    assert np.array_equal(np.asarray(bulk.map) == 1, g["bulk_map"] == 1)
    assert dens.map.dtype == np.int16 and np.array_equal(dens.map, ref["map"])
    assert dens.P["threshold"] == float(g["P_threshold"]) and dens.P["bulk_threshold"] == float(g["P_bulk_threshold"])
    assert dens.P["bulk_site"] == int(g["P_bulk_site"])
    assert [len(s) for s in dens.sites] == ref["ncells"].tolist()
    sp = dens.site_properties
    assert np.array_equal(sp.label, np.arange(len(ref["ncells"])))
    centers = np.array([np.full(3, np.nan) if c is None else np.asarray(c, dtype=float) for c in sp.center])
    check_site_properties(sp.volume, sp.occupancy_avg, sp.occupancy_std, centers, ref)

    group = u.select_atoms("name OH2")
    h = HoppingTrajectory(u.trajectory, group, dens, verbosity=0)
    fn = str(tmp_path / "hop")
    h.write(fn)
    assert open(fn + ".psf").read() == str(g["psf_text"])        # byte-identical PSF
    frames = np.stack([ts._pos.copy() for ts in h])
    assert frames.dtype == np.float32 and np.all(frames[:, :, 2] == 0)
    assert np.array_equal(frames[:, :, 0], ref["hop_x"]) and np.array_equal(frames[:, :, 1], ref["hop_y"])
    h2 = HoppingTrajectory(hopdcd=fn + ".dcd", hoppsf=fn + ".psf")
    assert np.allclose(h2.ts.dimensions, g["hop_cell"])
    checked_rates = False
    for traj in (h, HoppingTrajectory(u.trajectory, group, dens, verbosity=0)):   # from the DCD, and fused
        tn = TransportNetwork(traj, dens)
        hg = tn.HoppingGraph()
        assert abs(tn.dt - dt) < 1e-6
        gp = {k: dict(v, tau=np.asarray(v["tau"]) * (dt / tn.dt)) for k, v in tn.graph_properties.items()}
        for k, v in gp.items():
            if isinstance(k, tuple):
                v["tbarrier"] = np.asarray(v["tbarrier"]) * (dt / tn.dt)
        check_transitions(gp, ref, dt)
        assert tn.transition_counts() == {k: len(d["tau"]) for k, d in ref["props"].items()}
        busiest = sorted(ref["rates"], key=lambda k: -ref["rates"][k][1])[:12]
        for k in busiest:
            k_ref, n_ref = ref["rates"][k]
            assert hg.graph[k[0]][k[1]]["N"] == n_ref
            # The survival function weighs a waiting time that EQUALS a mesh point by 1/2
            # (hop/graph.py:2501-2518), so the fit depends on whether dt*frames lands exactly on the
            # integer mesh.  The reference run had the exact dt (in-memory reader); through the DCD
            # header dt makes a float32 AKMA round trip (as it does with MDAnalysis) and the fit is
            # only comparable when it comes back exact.
            if tn.dt == dt:
                checked_rates = True
                assert abs(hg.graph[k[0]][k[1]]["k"] - k_ref) <= 1e-6 * abs(k_ref), (k, hg.graph[k[0]][k[1]]["k"], k_ref)
    assert checked_rates or not ref["rates"]
    tn.graph_alltransitions()
    assert set(tn.graph.edges()) == ref["alltransitions"]


# ------------------------------------------------------------------------------------------------
# secondary site-map behaviours (MINsite = 2, re-mapping, site_insert_nobulk, site_labels, stats)
# ------------------------------------------------------------------------------------------------
def load_variants():
    return np.load(os.path.join(HERE, "golden", "reference", "variants", "variants_clean.npz"), allow_pickle=False)


def canonical(site_map, first=1):
    """Map with its sites of equal volume renumbered canonically; `first` = first label that may move."""
    m = np.asarray(site_map).astype(np.int64)
    shifted = np.where(m >= first, m + (2 - first), np.where(m > 0, 1, m))     # reuse canonical_perm (moves >= 2)
    perm = canonical_perm(shifted)
    out = relabel(shifted, perm)
    return np.where(out >= 2, out - (2 - first), m)


def test_oracle_matches_reference_sitemap_variants():
    g = load_variants()
    xyz, unit = g["xyz"], str(g["density_unit"])
    bins, arange, edges = O.grid_geometry(xyz[0], float(g["delta"]), 2.0)
    grid = O.make_density(O.histogram_vectorised(xyz, edges), edges, xyz.shape[0], unit)
    for label in (O.label_sites_faithful, O.label_sites_ndimage):
        s1 = label(grid, float(g["solvent_threshold"]), 1)
        s2 = label(grid, float(g["solvent_threshold"]), 2)
        assert np.array_equal(O.draw_map_from_sites(s1, grid.shape), canonical(g["minsite1_map"]))
        assert np.array_equal(O.draw_map_from_sites(s2, grid.shape), canonical(g["minsite2_map"]))
        assert [len(s) for s in s2] == g["minsite2_ncells"].tolist()
        s3 = label(grid, 1.5 * float(g["solvent_threshold"]), 1)
        assert np.array_equal(O.draw_map_from_sites(s3, grid.shape), canonical(g["remap_map"]))


@pytest.mark.gpu
def test_cuda_path_matches_reference_sitemap_variants():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import warnings

    from hop_b200.density import DensityCreator
    from hop_b200.universe import Universe

    g = load_variants()
    unit = str(g["density_unit"])
    u = Universe(coordinates=g["xyz"], dt=1.0, filename="mem.psf")

    def fresh(**parameters):
        d = DensityCreator(u, mode="solvent", delta=float(g["delta"]), atomselection="name OH2", padding=2.0,
                           parameters=dict(parameters)).create()["solvent"]
        d.convert_density(unit)
        return d

    d2 = fresh(MINsite=2)
    d2.map_sites(float(g["solvent_threshold"]))
    assert np.array_equal(d2.map, canonical(g["minsite2_map"]))
    assert [len(s) for s in d2.sites] == g["minsite2_ncells"].tolist()
    d1 = fresh()
    d1.map_sites(float(g["solvent_threshold"]))
    assert np.array_equal(d1.map, canonical(g["minsite1_map"]))
    bulk = fresh()
    bulk.map_sites(float(g["bulk_threshold"]))
    d1.site_insert_bulk(bulk)
    assert np.array_equal(d1.map, canonical(g["with_bulk_map"], first=2))
    for key, kw in (("default", {}), ("sites", dict(include="sites")), ("all_none", dict(include="all", exclude=None)),
                    ("no_bulk", dict(exclude=["default", "bulk"])), ("explicit", dict(include=[2, 3, 5], exclude=None))):
        assert np.array_equal(np.asarray(d1.site_labels(**kw)), g["labels_" + key]), key
    labels, vol = d1.site_volume(include="sites")
    assert np.array_equal(labels, g["site_volume_labels"]) and np.allclose(vol, g["site_volume"], rtol=1e-9)
    # equal-volume sites may be numbered differently: compare occupancies as multisets per volume
    labels, occ, occ_std = d1.site_occupancy(include="sites")
    for v in np.unique(g["site_volume"]):
        sel = g["site_volume"] == v
        assert np.allclose(np.sort(np.asarray(occ)[sel]), np.sort(g["site_occupancy"][sel]), rtol=1e-6, atol=1e-12)
        assert np.allclose(np.sort(np.asarray(occ_std)[sel]), np.sort(g["site_occupancy_std"][sel]), rtol=1e-6, atol=1e-9)
    st = d1.stats()
    assert sorted(st) == g["stats_keys"].tolist()
    assert np.allclose([float(st[k]) for k in sorted(st)], g["stats_values"], rtol=1e-6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                      # "Removed bulk site in order to map sites ..."
        d1.map_sites(1.5 * float(g["solvent_threshold"]))
    assert np.array_equal(d1.map, canonical(g["remap_map"]))
    assert d1.has_bulk() == bool(g["remap_has_bulk"]) and sorted(d1.P.keys()) == g["remap_P_keys"].tolist()
    d1.site_insert_nobulk()
    assert np.array_equal(d1.map, canonical(g["nobulk_map"], first=2))
    assert [len(s) for s in d1.sites] == g["nobulk_ncells"].tolist()
    assert d1.has_bulk() == bool(g["nobulk_has_bulk"]) and (d1.P["bulk_threshold"] is None) == bool(g["nobulk_bulk_threshold_is_none"])


# ------------------------------------------------------------------------------------------------
# live: the reference against the oracle on random small systems (only where /root/reference exists)
# ------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.isdir("/root/reference/hop"), reason="reference sources not present")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_live_reference_random(seed):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_reference_golden as M
    from hop_b200.synth import SynthParams, generate

    rng = np.random.default_rng(1000 + seed)
    n_atoms, n_frames = int(rng.integers(60, 160)), int(rng.integers(12, 40))
    delta = float(rng.choice([1.0, 0.8, 0.6]))
    xyz = generate(SynthParams(n_atoms=n_atoms, n_frames=n_frames, seed=77 + seed, site_stride=int(rng.integers(3, 9))))
    if seed % 2:                                             # outliers, also in the first frame
        m = rng.random(xyz.shape[:2]) < 0.05
        xyz[m] += np.float32(rng.choice([-40.0, 35.0]))
    unit = "water" if seed % 3 else "Angstrom^{-3}"
    water = 1.0 / (1e-24 * 6.02214179e+23 * 0.997 / 18.016)
    scale = 1.0 if unit == "water" else 1.0 / water
    thr, bthr = float(rng.uniform(1.5, 3.0)) * scale, float(rng.uniform(0.3, 0.8)) * scale
    try:
        res = M.run_reference(xyz, n_atoms, delta=delta, density_unit=unit, solvent_threshold=thr,
                              bulk_threshold=bthr, dt=float(rng.choice([1.0, 0.5])), rates=False)
    except ValueError as exc:                                # hop itself refuses (e.g. fewer than 2 sites)
        pytest.skip("reference raised: %s" % exc)
    finally:
        M.R.unload_reference()
    g = dict(res, n_solvent=n_atoms, hop_x=res["hop"][:, :, 0], hop_y=res["hop"][:, :, 1], rates=np.zeros((0, 4)))
    ref = reference_products(g)
    out = O.pipeline(xyz, delta=delta, solvent_threshold=thr, bulk_threshold=bthr, density_unit=unit, faithful=False)
    for a, b in zip(out["edges"], res["edges"]):
        assert np.array_equal(a, b)
    assert np.array_equal(out["grid"], res["grid"])
    assert np.array_equal(out["map"], ref["map"])
    assert np.array_equal(out["hop"][:, :, 0], ref["hop_x"]) and np.array_equal(out["hop"][:, :, 1], ref["hop_y"])
    props, _, _ = O.analyze_hops_faithful(out["hop"][:, :, 0], dt=1.0)
    ev_ref = {k: (d["frame"], d["iatom"]) for k, d in ref["props"].items()}
    assert {k: (list(v["frame"]), list(v["iatom"])) for k, v in props.items() if isinstance(k, tuple)} == ev_ref


@pytest.mark.skipif(not os.path.isdir("/root/reference/hop"), reason="reference sources not present")
def test_reference_equivalence_sites_cannot_run():
    """Why `find_equivalence_sites_with` (hop/sitemap.py:1082-1279, the last part of SURVEY.md 8f rank 4)
    is not built: with any numpy >= 1.3 (hop pins >= 1.12) `find_common_sites` (:1757-1790) gets a FLAT
    array from `numpy.unique(list of tuples)` instead of the unique pairs the code was written for, and
    the caller fails on it -- there is no reference behaviour to be in parity with."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_reference_golden as M

    try:
        hop = M.R.load_reference()
        xyz = M.synth(300, 60)

        def dens(frames):
            u = M.R.FakeUniverse(xyz[frames], selections={"name OH2": np.arange(300)})
            d = hop.density.DensityCreator(u, mode="solvent", delta=1.0, atomselection="name OH2",
                                           padding=2.0).create()["solvent"]
            d.convert_density("water")
            d.map_sites(math.e)
            return d

        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a, b = dens(np.r_[0:60]), dens(np.r_[0, 30:60])          # same frame 0 -> same grid
            m = hop.sitemap.find_common_sites(a, b)
            assert np.ndim(m) == 1                                   # labels, not (label_a, label_b) pairs
            with pytest.raises(ValueError):
                a.find_equivalence_sites_with(b)
    finally:
        M.R.unload_reference()


# ------------------------------------------------------------------------------------------------
# on-disk products written by hop itself (tests/golden/reference/pickles/*.pickle.gz: Saveable.save of the
# Density and HoppingGraph.save of the ref_clean run) must load here, without hop installed
# ------------------------------------------------------------------------------------------------
def _gunzip_to(tmp_path, name):
    import gzip

    dst = tmp_path / name
    with gzip.open(os.path.join(HERE, "golden", "reference", "pickles", name + ".gz"), "rb") as src:
        dst.write_bytes(src.read())
    return str(dst)


def test_density_pickle_written_by_hop_loads(tmp_path):
    from hop_b200.sitemap import Density

    assert "hop" not in sys.modules or not getattr(sys.modules["hop"], "__reference_py3__", False)
    g = load("ref_clean")
    fn = _gunzip_to(tmp_path, "density.pickle")
    d = Density(filename=fn)
    assert np.array_equal(d.grid, g["grid"]) and np.array_equal(d.map, g["map"]) and d.map.dtype == np.int16
    for a, k in zip(d.edges, ("edges_x", "edges_y", "edges_z")):
        assert np.array_equal(a, g[k])
    assert d.P["threshold"] == float(g["P_threshold"]) and d.P["bulk_threshold"] == float(g["P_bulk_threshold"])
    assert d.P["bulk_site"] == 1 and d.P["isDensity"] and d.has_bulk()
    assert d.unit["length"] == "Angstrom" and d.unit["density"] == "Angstrom^{-3}"
    assert d.metadata["collector"] == "solvent" and d.metadata["n_frames"] == g["hop_x"].shape[0]
    assert [len(s) for s in d.sites] == g["site_ncells"].tolist()
    assert sorted(map(tuple, d.sites[1])) == sorted(map(tuple, np.argwhere(g["bulk_map"] == 1).tolist()))
    sp = d.site_properties
    assert np.allclose(sp.volume, g["site_volume"]) and np.allclose(sp.occupancy_avg, g["site_occupancy_avg"])
    assert np.array_equal(np.asarray(d.site_labels("sites")), np.arange(2, len(g["site_ncells"])))
    # and what is saved again loads again, with the same content
    d.save(str(tmp_path / "again"))
    d2 = Density(filename=str(tmp_path / "again"))
    assert np.array_equal(d2.grid, d.grid) and np.array_equal(d2.map, d.map) and dict(d2.P) == dict(d.P)


def test_hopgraph_pickle_written_by_hop_loads(tmp_path):
    from hop_b200.graph import HoppingGraph

    g = load("ref_clean")
    fn = _gunzip_to(tmp_path, "hopgraph.pickle")
    hg = HoppingGraph(filename=fn)
    rates = {(int(a), int(b)): (k, int(n)) for a, b, k, n in g["rates"]}
    assert set(hg.graph.edges()) == set(rates)
    for (a, b), (k, n) in rates.items():
        assert hg.graph[a][b]["N"] == n and hg.graph[a][b]["k"] == k
        assert hg.number_of_hops(a, b) == n and len(hg.waitingtimes(a, b)) == n
    e = max(rates, key=lambda e: rates[e][1])
    fit = hg.waitingtime_fit(e)
    assert type(fit).__module__ == "hop_b200.graph" and np.all(np.isfinite(fit.fit(np.arange(5.0))))
    assert hg.trjdata["time_unit"] == "ps" and hg.trjdata["dt"] == float(g["dt"])
    assert np.allclose(hg.site_properties.volume, g["site_volume"])
    hg.save(str(tmp_path / "again"))
    hg2 = HoppingGraph(filename=str(tmp_path / "again"))
    assert set(hg2.graph.edges()) == set(hg.graph.edges()) and hg2.trjdata == hg.trjdata
    assert all(np.array_equal(hg2.properties[k]["tau"], hg.properties[k]["tau"]) for k in hg.properties)


def test_site_labels_and_stats_on_hop_density_match_hop(tmp_path):
    """Host-side selections (Density.site_labels, hop/sitemap.py:629-760, and Density.stats, :1301-1339) on
    the density hop itself saved, against what hop returned for the same arguments (CPU only)."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from hop_b200.sitemap import Density

    combos_src = open(os.path.join(HERE, "golden", "make_reference_golden.py")).read()
    ns = {}
    exec(combos_src[combos_src.index("SITE_LABEL_COMBOS = ["):combos_src.index("def synth(")], ns)
    g = np.load(os.path.join(HERE, "golden", "reference", "pickles", "site_labels.npz"), allow_pickle=False)
    d = Density(filename=_gunzip_to(tmp_path, "density.pickle"))
    for i, (inc, exc) in enumerate(ns["SITE_LABEL_COMBOS"]):
        want = g["labels_%02d" % i]
        if "error_%02d" % i in g.files:
            with pytest.raises(Exception):
                d.site_labels(include=inc, exclude=exc)
            continue
        got = np.asarray(d.site_labels(include=inc, exclude=exc), dtype=np.int64)
        assert np.array_equal(got, want), (inc, exc, got[:8], want[:8])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st = d.stats()
    assert sorted(st) == g["stats_keys"].tolist()
    assert np.allclose([float(st[k]) for k in sorted(st)], g["stats_values"], rtol=1e-12)


# ------------------------------------------------------------------------------------------------
# Per-site observables (hop/siteanalysis.py)
# ------------------------------------------------------------------------------------------------
SA_FILE = os.path.join(HERE, "golden", "reference", "siteanalysis", "siteanalysis.npz")
SA_CASES = [(tag, obs, plane) for tag in ("default", "with_bulk") for obs, plane in (("distance", "hop_x"), ("orbit", "hop_y"))]


def _sa_rows(g, tag):
    L = len(g["site_center"])
    sites = g[tag + "_sites"]
    s2i = np.full(L, len(sites), dtype=np.int64)
    s2i[sites] = np.arange(len(sites))
    return s2i


@pytest.mark.parametrize("tag,obs,plane", SA_CASES)
def test_oracle_site_distance_histograms_match_reference(tag, obs, plane):
    """oracle.site_distance_histogram against the histograms hop's own Collection / Distance / Orbit produced
    (hop/siteanalysis.py:722-802, 531-544), on a trajectory with outliers: the fancy-index `+=`, the outlier
    label indexing from the end and the NaN distance of the interstitial are all part of the reference's answer."""
    g = dict(np.load(SA_FILE))
    assert (g["hop_x"] == -1).sum() > 100
    H = O.site_distance_histogram(g["xyz"], g[plane], g["site_center"], _sa_rows(g, tag), g["%s_%s_edges" % (tag, obs)],
                                  0, int(g["stop"]))
    assert np.array_equal(H, g["%s_%s_histogram" % (tag, obs)])
    assert np.array_equal(g["%s_%s_edges" % (tag, obs)], O.site_histogram_edges(0, 3 if obs == "distance" else 6,
                                                                                 30 if obs == "distance" else 24))


def test_reference_occupancy_cannot_run():
    """The Occupancy observable of the reference raises for any input (hop/siteanalysis.py:821-841:
    `_occupancy_histogram` returns one value fewer than `histogrammed_sites` has labels, and
    `SiteHistogramArray.add` indexes with both) -- recorded when the golden file was made, and shown live where
    the reference sources are present.  hop_b200.siteanalysis implements the documented meaning; its parity is
    unpinned."""
    g = dict(np.load(SA_FILE))
    assert "IndexError" in str(g["occupancy_error"])
    if not os.path.isdir("/root/reference/hop"):
        return
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import warnings

    import make_reference_golden as M

    try:
        SA = M.R.load_siteanalysis()
        hop = M.R.load_reference()
        xyz = M.synth(120, 20)
        u = M.R.FakeUniverse(xyz, selections={"name OH2": np.arange(120)})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = hop.density.DensityCreator(u, mode="solvent", delta=1.0, atomselection="name OH2", padding=2.0).create()["solvent"]
            d.convert_density("water")
            d.map_sites(math.e)
            group = u.select_atoms("name OH2")
            u.trajectory.rewind()
            ht = hop.trajectory.HoppingTrajectory(u.trajectory, group, d, verbosity=0)
            import tempfile

            with tempfile.TemporaryDirectory() as tmp:
                ht.write(os.path.join(tmp, "h"))
            ht.group = M.R.LegacySelection(ht.group)
            C = SA.Collection(selection=M.R.LegacySelection(group), hoptraj=ht, density=d)
            C.add_observable("occupancy", lo_bin=0, hi_bin=5)
            u.trajectory.rewind()
            with pytest.raises(IndexError):
                C.compute(stop=19)
    finally:
        M.R.unload_reference()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,obs,plane", SA_CASES)
def test_cuda_site_distance_histograms_match_reference(tag, obs, plane):
    """hopb_site_distance_hist on the reference's own hopping trajectory and site centres: the histograms of hop's
    Collection.compute, count for count."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hop_b200 import kernels as K

    g = dict(np.load(SA_FILE))
    want = g["%s_%s_histogram" % (tag, obs)]
    hist = torch.zeros(want.shape, dtype=torch.int64, device="cuda")
    K.site_distance_hist(torch.from_numpy(g["xyz"]).cuda(), torch.from_numpy(g[plane].astype(np.float32)).cuda(),
                         torch.from_numpy(g["site_center"]).cuda(), torch.from_numpy(_sa_rows(g, tag).astype(np.int32)).cuda(),
                         torch.from_numpy(g["%s_%s_edges" % (tag, obs)]).cuda(), want.shape[0], hist, start=0, stop=int(g["stop"]))
    assert np.array_equal(hist.cpu().numpy(), want)
    # a strided frame range and a permuted atom order (idcd2ihop) against the oracle
    perm = np.random.default_rng(1).permutation(g["xyz"].shape[1]).astype(np.int32)
    hist.zero_()
    K.site_distance_hist(torch.from_numpy(np.ascontiguousarray(g["xyz"][:, perm])).cuda(),
                         torch.from_numpy(g[plane].astype(np.float32)).cuda(), torch.from_numpy(g["site_center"]).cuda(),
                         torch.from_numpy(_sa_rows(g, tag).astype(np.int32)).cuda(), torch.from_numpy(g["%s_%s_edges" % (tag, obs)]).cuda(),
                         want.shape[0], hist, start=3, stop=50, step=4, atom_map=torch.from_numpy(perm).cuda())
    assert np.array_equal(hist.cpu().numpy(), O.site_distance_histogram(g["xyz"], g[plane], g["site_center"], _sa_rows(g, tag),
                                                                          g["%s_%s_edges" % (tag, obs)], 3, 50, 4))


@pytest.mark.gpu
def test_cuda_siteanalysis_collection():
    """The drop-in Collection end to end on its own pipeline products (its site numbering is the canonical one, the
    reference's is not, so the class-level comparison is against the oracle restatement, which the tests above pin
    to the reference): distance / orbit histograms, distribution and mean; occupancy as documented."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hop_b200.density import DensityCreator
    from hop_b200.siteanalysis import Collection
    from hop_b200.sitemap import Density
    from hop_b200.trajectory import HoppingTrajectory
    from hop_b200.universe import Universe

    g = dict(np.load(SA_FILE))
    xyz = g["xyz"]
    F, N = xyz.shape[:2]
    u = Universe(coordinates=xyz, names=["OH2"] * N, resnames=["TIP3"] * N, filename="mem.psf")
    dens = DensityCreator(u, mode="solvent", delta=1.0, atomselection="name OH2", padding=2.0).create()["solvent"]
    bulk = Density(grid=np.array(dens.grid), edges=[np.array(e) for e in dens.edges], parameters={"isDensity": True},
                   unit={"length": "Angstrom", "density": "Angstrom^{-3}"})
    dens.convert_density("water")
    dens.map_sites(math.e)
    bulk.convert_density("water")
    bulk.map_sites(math.exp(-0.5))
    dens.site_insert_bulk(bulk)
    group = u.select_atoms("name OH2")
    ht = HoppingTrajectory(u.trajectory, group, dens)
    hx, hy = ht.hop_arrays()
    assert np.array_equal(relabel(g["hop_x"], canonical_perm(g["map"])), hx.astype(np.int64))
    C = Collection(selection=group, hoptraj=ht, density=dens)
    C.add_observable("distance", lo_bin=0, hi_bin=3)
    C.add_observable("orbit", lo_bin=0, hi_bin=6, Nbins=24)
    C.add_observable("occupancy", lo_bin=0, hi_bin=5, trajectory=True)
    with pytest.raises(NotImplementedError):
        C.add_observable("distance", lo_bin=0, hi_bin=3, trajectory=True)
    with pytest.raises(ValueError):
        C.add_observable("nonsense", lo_bin=0, hi_bin=3)
    C.compute(stop=F - 1)
    centers = np.array([np.full(3, np.nan) if (c is None or np.ndim(c) == 0) else np.asarray(c, float)
                        for c in dens.site_properties.center])
    centers[0] = np.nan
    L = len(centers)
    for obs, plane in (("distance", hx), ("orbit", hy)):
        H = C.SiteHistograms[obs]
        s2i = np.full(L, H.Nsites, dtype=np.int64)
        s2i[H.sites] = np.arange(H.Nsites)
        want = O.site_distance_histogram(xyz, plane, centers, s2i, H.edges, 0, F - 1)
        assert np.array_equal(H._histogram, want)
        norm = want[:-1].sum(axis=1).astype(float)
        norm[norm == 0] = -1
        assert np.allclose(H.distribution, want[:-1, 1:-1] / (norm[:, None] * H.delta))
    # the same physical sites as the reference, whatever they are called: the multiset of per-site histograms agrees
    # for all sites that no outlier aliases onto (the last label differs between the two numberings)
    from collections import Counter

    ref_rows = Counter(map(tuple, g["default_distance_histogram"][:-1].tolist()))
    got_rows = Counter(map(tuple, C.SiteHistograms["distance"]._histogram[:-1].tolist()))
    assert sum((ref_rows - got_rows).values()) <= 2 and sum((got_rows - ref_rows).values()) <= 2
    occ = O.site_occupancy(hx[:F - 1], L)
    H = C.SiteHistograms["occupancy"]
    for s in H.sites[:20]:
        assert np.array_equal(H.site_histogram(int(s)), np.bincount(np.digitize(occ[:, s], H.edges), minlength=len(H.edges) + 1)[1:-1])
    assert np.array_equal(C.SiteTrajectories["occupancy"].site_trajectory(H.sites)[:, :F - 1], occ[:, H.sites].T)
