"""The reference-facing Python API (DensityCreator / Density / HoppingTrajectory / TransportNetwork)
against the oracle, written the way hop's own workflow calls it (hop/interactive.py:235-307,
417-464, scripts/hop-generate-*.py)."""
import math
import os
import pickle

import numpy as np
import pytest

from oracle import hop_oracle as O
from hop_b200.synth import SynthParams, generate

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def system():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hop_b200.universe import Universe

    p = SynthParams(n_atoms=800, n_frames=120)
    xyz = generate(p)
    # a mixed system: waters (OH2) interleaved with a few atoms of other names
    names = np.array(["OH2"] * 800, dtype=object)
    names[::97] = "SOD"
    u = Universe(coordinates=xyz, names=list(names), dt=2.0, filename="synthetic.psf")
    sel = np.nonzero(names == "OH2")[0]
    ref = O.pipeline(xyz[:, sel], delta=1.0, faithful=False, with_bulk=False)
    return u, xyz, sel, ref


def test_density_creator_solvent(system):
    from hop_b200.density import DensityCreator

    u, xyz, sel, ref = system
    DC = DensityCreator(u, mode="solvent", delta=1.0, atomselection="name OH2")
    dens = DC.create()["solvent"]
    for a, b in zip(dens.edges, ref["edges"]):
        assert np.array_equal(a, b)
    ref_grid = O.make_density(ref["counts"], ref["edges"], xyz.shape[0])
    assert np.array_equal(dens.grid, ref_grid)
    assert dens.P["isDensity"] and dens.unit["density"] == "Angstrom^{-3}"
    md = dens.metadata
    assert md["collector"] == "solvent" and md["collector_mode"] == "SOLVENT" and md["n_frames"] == 120
    assert md["atomselection"] == "name OH2" and md["dt"] == 2.0 and md["totaltime"] == 238.0
    assert np.allclose(np.diag(dens.delta), 1.0, atol=1e-6)
    dens.convert_density("water")
    assert np.array_equal(dens.grid, ref["grid"])
    with pytest.raises(ValueError):
        dens.grid[0, 0, 0] = 1.0   # mirrored arrays are read-only


def test_collector_per_frame_api_matches_block_path(system):
    from hop_b200.density import DensityCollector

    u, xyz, sel, ref = system
    c = DensityCollector("solvent", u, delta=1.0, atomselection="name OH2", soluteselection=None, cutoff=0)
    c.init_histogram()
    n = 0
    for ts in u.trajectory:
        n += c.collect()
    assert n == len(sel) * 120
    c.finish(u.trajectory.n_frames)
    d = c.Density()
    assert np.array_equal(d.grid, O.make_density(ref["counts"], ref["edges"], 120))
    assert np.array_equal(c.grid, ref["counts"] / 120.0)


def test_error_conventions(system):
    from hop_b200.density import DensityCollector, DensityCreator
    from hop_b200.exceptions import MissingDataError
    from hop_b200.sitemap import Density
    from hop_b200.trajectory import HoppingTrajectory

    u, xyz, sel, ref = system
    with pytest.raises(TypeError):
        DensityCollector("x", object())
    with pytest.raises(ValueError):
        DensityCreator(u, mode="nonsense")
    c = DensityCollector("solvent", u, soluteselection=None, cutoff=0)
    with pytest.raises(MissingDataError):
        c.Density()
    d = Density(grid=ref["grid"], edges=ref["edges"], parameters={"isDensity": True}, unit={"length": "Angstrom", "density": "water"})
    with pytest.raises(ValueError):
        d.map_sites()               # no threshold
    with pytest.raises(NotImplementedError):
        Density(grid=ref["grid"], edges=ref["edges"], parameters={"isDensity": True, "MINsite": 3}, unit={"density": "water"})
    with pytest.raises(ValueError):
        HoppingTrajectory(u.trajectory, u.select_atoms("name OH2"), d)   # map not computed
    d.map_sites(math.e)
    with pytest.raises(TypeError):
        HoppingTrajectory(u.trajectory, object(), d)
    with pytest.raises(ValueError):
        HoppingTrajectory(u.trajectory, u.select_atoms("name XYZ"), d)
    small = Density(grid=np.ones((3, 3, 3)), edges=[np.arange(4.0)] * 3, parameters={"isDensity": True}, unit={"density": "water"})
    small.map_sites(0.5)
    with pytest.raises(ValueError):
        d.site_insert_bulk(small)   # different grid
    with pytest.raises(ValueError):
        HoppingTrajectory()


def test_map_sites_bulk_and_properties(system):
    from hop_b200.sitemap import Density

    u, xyz, sel, ref = system
    unit = {"length": "Angstrom", "density": "water"}
    solvent = Density(grid=ref["grid"], edges=ref["edges"], parameters={"isDensity": True}, unit=unit)
    bulk = Density(grid=ref["grid"], edges=ref["edges"], parameters={"isDensity": True}, unit=unit)
    solvent.map_sites(math.e)
    assert solvent.map.dtype == np.int16 and np.array_equal(solvent.map, ref["map"])
    assert [list(s) for s in solvent.sites] == ref["sites"]
    assert solvent.P["threshold"] == math.e and not solvent.has_bulk()
    bulk.map_sites(math.exp(-0.5))
    solvent.site_insert_bulk(bulk)
    assert solvent.has_bulk() and solvent.P["bulk_site"] == 1 and solvent.P["bulk_threshold"] == math.exp(-0.5)
    full = O.site_insert_bulk(ref["sites"], O.label_sites_ndimage(ref["grid"], math.exp(-0.5)))
    assert np.array_equal(solvent.map, O.draw_map_from_sites(full, ref["grid"].shape))
    assert [list(s) for s in solvent.sites] == full
    props = O.site_properties(full, ref["grid"], ref["edges"], "water")
    sp = solvent.site_properties
    assert list(sp.label) == list(range(len(full)))
    np.testing.assert_allclose(sp.volume, props["volume"], rtol=1e-12)
    np.testing.assert_allclose(sp.occupancy_avg, props["occupancy_avg"], rtol=1e-9)
    np.testing.assert_allclose(sp.occupancy_std, props["occupancy_std"], rtol=1e-6, atol=1e-12)
    assert sp.center[0] is None
    np.testing.assert_allclose(np.stack(sp.center[1:]), props["center"][1:], rtol=1e-9)
    assert list(solvent.site_labels("sites")) == list(range(2, len(full)))
    assert list(solvent.site_labels()) == list(range(1, len(full)))
    labels, vol = solvent.site_volume(include=[1, 2])
    assert list(labels) == [1, 2]
    with pytest.raises(ValueError):
        solvent.site_insert_bulk(bulk)    # already has one
    with pytest.warns(UserWarning):
        solvent.map_sites(math.e)          # removes the bulk site again
    assert not solvent.has_bulk() and np.array_equal(solvent.map, ref["map"])
    solvent.site_insert_nobulk()
    assert solvent.has_bulk() and len(solvent.sites[1]) == 0
    assert np.array_equal(solvent.map, np.where(ref["map"] > 0, ref["map"] + 1, 0))


def test_threshold_scan_matches_oracle(system, tmp_path):
    """DensityScanner.scan (hop/analysis.py:165-195): per cut-off map_sites + site_insert_bulk + stats."""
    import warnings

    from hop_b200.analysis import DensityScanner
    from hop_b200.sitemap import Density

    u, xyz, sel, ref = system
    unit = {"length": "Angstrom", "density": "water"}
    solvent = Density(grid=ref["grid"], edges=ref["edges"], parameters={"isDensity": True}, unit=unit)
    bulk = Density(grid=ref["grid"], edges=ref["edges"], parameters={"isDensity": True}, unit=unit)
    bulk.map_sites(math.exp(-0.5))
    cutoffs = [1.5, 2.0, math.e, 3.5]
    scanner = DensityScanner(solvent, bulk)
    with pytest.raises(ValueError):
        DensityScanner(solvent).scan(cutoffs)          # bulk densities are required
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                 # "Removed bulk site in order to map sites"
        scanner.scan(cutoffs)
    want = O.scan_thresholds(ref["grid"], ref["edges"], O.label_sites_ndimage(ref["grid"], math.exp(-0.5)), cutoffs, "water")
    assert sorted(scanner.scanstats) == sorted(cutoffs)
    for c in cutoffs:
        got = scanner.scanstats[c][0]
        assert got["rho_cut"] == c and got["rho_cut_bulk"] == math.exp(-0.5)
        assert got["N_sites"] == want[c]["N_sites"] and got["N_equivalence_sites"] == 0
        for k in ("site_volume_avg", "site_volume_std", "site_occupancy_rho_avg", "site_occupancy_rho_std"):
            assert got[k] == pytest.approx(want[c][k], rel=1e-6), (c, k)
    arr = scanner.scanarrays[0]
    assert list(arr.rho_cut) == sorted(cutoffs) and list(arr.N_sites) == [want[c]["N_sites"] for c in sorted(cutoffs)]
    data = {}
    solvent.stats(data=data)
    assert len(data["site_volume"]) == scanner.scanstats[cutoffs[-1]][0]["N_sites"]
    scanner.save(str(tmp_path / "scan"))
    again = DensityScanner()
    again.load(str(tmp_path / "scan"))
    assert sorted(again.scanstats) == sorted(cutoffs) and list(again.scanarrays[0].N_sites) == list(arr.N_sites)


def test_density_pickle_schema(system, tmp_path):
    from hop_b200.sitemap import Density

    u, xyz, sel, ref = system
    d = Density(grid=ref["grid"], edges=ref["edges"], parameters={"isDensity": True},
                unit={"length": "Angstrom", "density": "water"}, metadata={"psf": "a.psf"})
    d.map_sites(math.e)
    fn = str(tmp_path / "water")
    d.save(fn)
    raw = pickle.load(open(fn + ".pickle", "rb"))
    assert sorted(raw.__getstate__().keys()) == sorted(
        ['grid', 'edges', 'P', 'unit', 'metadata', 'map', 'sites', 'site_properties', 'equivalent_sites_index'])
    d2 = Density(filename=fn)
    assert np.array_equal(d2.grid, ref["grid"]) and np.array_equal(d2.map, ref["map"]) and d2.map.dtype == np.int16
    assert [list(s) for s in d2.sites] == ref["sites"]
    assert d2.P["threshold"] == math.e and d2.metadata["psf"] == "a.psf"
    d2.export(str(tmp_path / "water"))
    assert os.path.exists(str(tmp_path / "water.dx"))


@pytest.fixture(scope="module")
def mapped(system):
    from hop_b200.density import DensityCreator

    u, xyz, sel, ref = system
    DC = DensityCreator(u, mode="solvent", delta=1.0, atomselection="name OH2")
    dens = DC.create()["solvent"]
    dens.convert_density("water")
    dens.map_sites(math.e)
    bulk = DensityCreator(u, mode="solvent", delta=1.0, atomselection="name OH2").create()["solvent"]
    bulk.convert_density("water")
    bulk.map_sites(math.exp(-0.5))
    dens.site_insert_bulk(bulk)
    refb = O.pipeline(xyz[:, sel], delta=1.0, faithful=False, with_bulk=True)
    assert np.array_equal(dens.map, refb["map"])
    return dens, refb


def test_hopping_trajectory_and_files(system, mapped, tmp_path):
    from hop_b200.trajectory import HoppingTrajectory

    u, xyz, sel, _ = system
    dens, ref = mapped
    group = u.select_atoms("name OH2")
    h = HoppingTrajectory(u.trajectory, group, dens)
    assert h.n_frames == 120 and h.ts.n_atoms == len(sel)
    assert np.array_equal(h.ts._pos, ref["hop"][0])
    frames = []
    for ts in h.map_dcd():
        frames.append(ts._pos.copy())
    assert np.array_equal(np.stack(frames), ref["hop"])
    # a pass that starts later re-initialises the orbit column (trajectory.py:392, 396-401)
    sub = O.coord2hop_vectorised(xyz[40:100:3][:, sel], ref["edges"], ref["map"])
    got = np.stack([ts._pos.copy() for ts in h.map_dcd(40, 100, 3)])
    assert np.array_equal(got, sub)
    h_seq = HoppingTrajectory(u.trajectory, group, dens)          # next(): sequential mapping (trajectory.py:230-252)
    assert np.array_equal(h_seq.next()._pos, ref["hop"][1]) and np.array_equal(h_seq.next()._pos, ref["hop"][2])
    fn = str(tmp_path / "hoptraj")
    h.write(fn)
    assert os.path.exists(fn + ".dcd") and os.path.exists(fn + ".psf")
    assert h.hoptraj is not None and h.n_frames == 120
    got = np.stack([ts._pos.copy() for ts in h])
    assert np.array_equal(got, ref["hop"])
    psf = open(fn + ".psf").read().splitlines()
    assert psf[0] == "PSF EXT" and "!NATOM" in psf[7] and int(psf[7].split()[0]) == len(sel)
    h2 = HoppingTrajectory(hopdcd=fn + ".dcd", hoppsf=fn + ".psf")
    assert h2.n_frames == 120 and abs(h2.hoptraj.dt - 2.0) < 1e-5
    assert np.allclose(h2.ts.dimensions[:3], [ref["map"].max() + 2, ref["map"].max() + 2, 1])


def test_transport_network(system, mapped, tmp_path):
    from hop_b200.graph import TransportNetwork
    from hop_b200.trajectory import HoppingTrajectory

    u, xyz, sel, _ = system
    dens, ref = mapped
    props_ref, edges_ref, nodes_ref = O.analyze_hops_faithful(ref["hop"][:, :, 0], dt=2.0)
    group = u.select_atoms("name OH2")
    h = HoppingTrajectory(u.trajectory, group, dens)
    fn = str(tmp_path / "hoptraj")
    h.write(fn)
    with pytest.raises(TypeError):
        TransportNetwork("not a hoptraj")
    with pytest.raises(ValueError):
        TransportNetwork(h, dens, sitelabels=[0, 1])._analyze_hops()
    for traj in (h, HoppingTrajectory(u.trajectory, group, dens)):   # from the file, and fused from coordinates
        tn = TransportNetwork(traj, dens)
        hg = tn.HoppingGraph()
        gp = tn.graph_properties
        # dt comes back from the DCD header as a float32 in AKMA units, like it does through MDAnalysis
        assert abs(tn.dt - 2.0) < 1e-6
        props_ref, edges_ref, nodes_ref = O.analyze_hops_faithful(ref["hop"][:, :, 0], dt=tn.dt)
        assert set(gp.keys()) == set(props_ref.keys())
        for k, d in props_ref.items():
            for f in ("tau", "frame", "iatom"):
                assert np.array_equal(gp[k][f], d[f]), (k, f)
            if isinstance(k, tuple):
                assert np.array_equal(gp[k]["tbarrier"], d["tbarrier"])
            else:
                assert all(v is None for v in gp[k]["tbarrier"])
        assert set(tn.graph.edges()) == set(edges_ref)
        import networkx as NX

        G = NX.DiGraph()
        for e in edges_ref:                               # the reference adds edges at first occurrence
            G.add_edge(*e)
        assert list(tn.graph.edges()) == list(G.edges()) and list(tn.graph.nodes())[:len(G)] == list(G.nodes())
        assert tn.transition_counts() == {e: len(props_ref[e]["tau"]) for e in edges_ref}
        assert hg.trjdata["time_unit"] == "ps" and abs(hg.trjdata["dt"] - 2.0) < 1e-6 and abs(hg.trjdata["totaltime"] - 238.0) < 1e-3
        assert hg.number_of_hops(*edges_ref[0]) == len(props_ref[edges_ref[0]]["tau"])
        for e in edges_ref[:6]:                           # rates from the GPU's events vs the oracle's fit
            k_ref, n_ref, _ = O.rate_faithful(props_ref[e]["tau"])
            assert hg.graph[e[0]][e[1]]["N"] == n_ref
            assert abs(hg.graph[e[0]][e[1]]["k"] - k_ref) <= 1e-6 * abs(k_ref)
    tn.graph_alltransitions()
    pairs = O.graph_alltransitions(ref["hop"][:, :, 0])
    assert set(tn.graph.edges()) == set(map(tuple, pairs.tolist()))
    hg.save(str(tmp_path / "hopgraph"))
    from hop_b200.graph import HoppingGraph

    hg2 = HoppingGraph(filename=str(tmp_path / "hopgraph"))
    assert set(hg2.graph.edges()) == set(hg.graph.edges())


def test_interactive_workflow(system, tmp_path):
    from hop_b200 import interactive

    u, xyz, sel, _ = system
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        dens = interactive.generate_densities(u, mode="solvent", delta=1.0, atomselection="name OH2")["solvent"]
        assert dens.has_bulk() and os.path.exists("water.pickle")
        hops = interactive.make_hoppingtraj(dens, "hoptraj", universe=u)
        hg = interactive.build_hoppinggraph(hops, dens)
        assert hg.graph.number_of_edges() > 0
    finally:
        os.chdir(cwd)


# ---- bulk density: '<solvent> not within <cutoff> of <solute>' (hop/density.py:82-87, 257-272, 306-356) ----
@pytest.fixture(scope="module")
def solvated():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hop_b200.universe import Universe

    p = SynthParams(n_atoms=900, n_frames=40)
    water = generate(p)
    rng = np.random.default_rng(12)
    c = p.box / 2
    prot0 = (c + rng.normal(0, 4.0, size=(60, 3))).astype(np.float32)
    prot = (prot0[None] + rng.normal(0, 0.3, size=(40, 60, 3))).astype(np.float32)   # a wobbling "protein"
    xyz = np.concatenate([prot, water], axis=1)
    names = ["CA"] * 30 + ["HA"] * 30 + ["OH2"] * 900
    resn = ["ALA"] * 60 + ["TIP3"] * 900
    u = Universe(coordinates=xyz, names=names, resnames=resn, dt=1.0)
    return u, xyz, water, prot[:, :30]     # heavy atoms only: 'protein and not name H*'


def test_notwithin_selection_and_fused_densities(solvated):
    from hop_b200.bulk import notwithin_coordinates_factory
    from hop_b200.density import DensityCollector, DensityCreator

    u, xyz, water, heavy = solvated
    cutoff = 3.5
    nw = notwithin_coordinates_factory(u, "name OH2", "protein and not name H*", cutoff)
    u.trajectory[7]
    ref_mask = O.within_cutoff(water[7], heavy[7], cutoff)
    assert 0 < ref_mask.sum() < len(ref_mask)
    assert np.array_equal(nw(), water[7][~ref_mask])
    assert np.array_equal(nw.within_mask(water[:5], heavy[:5]).astype(bool),
                          np.stack([O.within_cutoff(water[f], heavy[f], cutoff) for f in range(5)]))
    u.trajectory.rewind()

    DC = DensityCreator(u, mode="all", delta=1.0, atomselection="name OH2", cutoff=cutoff,
                        soluteselection="protein and not name H*")
    dens = DC.create()
    assert set(dens) == {"bulk", "solvent"}
    edges = dens["solvent"].edges
    for a, b in zip(edges, dens["bulk"].edges):
        assert np.array_equal(a, b)
    _, _, ref_edges = O.grid_geometry(water[0], delta=1.0)
    for a, b in zip(edges, ref_edges):
        assert np.array_equal(a, b)
    ref_solvent = O.make_density(O.histogram_vectorised(water, edges), edges, 40)
    ref_bulk = O.make_density(O.histogram_bulk(water, heavy, edges, cutoff), edges, 40)
    assert np.array_equal(dens["solvent"].grid, ref_solvent)
    assert np.array_equal(dens["bulk"].grid, ref_bulk)
    assert dens["bulk"].metadata["collector_mode"] == "BULK" and dens["bulk"].metadata["cutoff"] == cutoff
    assert dens["bulk"].metadata["soluteselection"] == "protein and not name H*"

    # the per-frame collector API gives the same histogram as the fused block path
    cb = DensityCollector("bulk", u, delta=1.0, atomselection="name OH2", cutoff=cutoff,
                          soluteselection="protein and not name H*")
    cb.init_histogram(smin=np.min(water[0], axis=0) - 2.0, smax=np.max(water[0], axis=0) + 2.0)
    for ts in u.trajectory:
        cb.collect()
    cb.finish(40)
    assert np.array_equal(cb.Density().grid, ref_bulk)


    # mode='bulk' alone
    only = DensityCreator(u, mode="bulk", delta=1.0, atomselection="name OH2", cutoff=cutoff,
                          soluteselection="protein and not name H*").create()["bulk"]
    assert np.array_equal(only.grid.sum(), only.grid.sum()) and only.grid.shape[0] > 0

    # DensityWithBulk: the standard workflow (hop/density.py:306-356)
    final = DC.DensityWithBulk(density_unit="water", solvent_threshold=math.e, bulk_threshold=math.exp(-0.5))
    gs = O.make_density(O.histogram_vectorised(water, edges), edges, 40, "water")
    gb = O.make_density(O.histogram_bulk(water, heavy, edges, cutoff), edges, 40, "water")
    sites = O.site_insert_bulk(O.label_sites_ndimage(gs, math.e), O.label_sites_ndimage(gb, math.exp(-0.5)))
    assert np.array_equal(final.map, O.draw_map_from_sites(sites, gs.shape))
    assert final.has_bulk() and [list(s) for s in final.sites] == sites


def test_large_solute_goes_through_in_tiles(solvated, monkeypatch):
    """A solute too large for the shared-memory cell list of hist_bulk_kernel is processed in tiles, the within
    flags carried from tile to tile (the kd-tree selection of the reference has no size limit, hop/density.py:82-87).
    The tile size is forced down so that the small test solute needs several tiles; also a really large one."""
    import torch

    from hop_b200 import kernels as K

    u, xyz, water, heavy = solvated
    cutoff = 3.5
    _, _, edges = O.grid_geometry(water[0], delta=1.0)
    g = K.GridTables(edges)
    w, p = torch.from_numpy(water).cuda(), torch.from_numpy(np.ascontiguousarray(heavy)).cuda()

    def run(with_flags):
        cs = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
        cb = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
        cells = torch.empty(w.shape[:2], dtype=torch.int32, device="cuda")
        flags = torch.full(w.shape[:2], 7, dtype=torch.uint8, device="cuda") if with_flags else None
        K.hist_bulk(g, w, p, cutoff, cb, counts_solvent=cs, cells=cells, within=flags)
        return cs, cb, cells, flags

    want = run(True)
    assert np.array_equal(want[1].cpu().numpy().astype(np.int64), O.histogram_bulk(water, heavy, edges, cutoff))
    monkeypatch.setenv("HOPB_BULK_TILE", "5")
    assert heavy.shape[1] > 10
    for with_flags in (True, False):
        got = run(with_flags)
        assert all(torch.equal(a, b) for a, b in zip(want[:3], got[:3]))
        if with_flags:
            assert torch.equal(want[3], got[3])
    monkeypatch.delenv("HOPB_BULK_TILE")
    # 40 000 solute atoms (more than one launch can stage): compared with the brute-force oracle on two frames
    rng = np.random.default_rng(2)
    big = (rng.random((2, 40000, 3)) * (water[0].max() * 0.3) + water[0].max() * 0.35).astype(np.float32)
    cb = torch.zeros(g.shape, dtype=torch.int32, device="cuda")
    flags = torch.empty((2, water.shape[1]), dtype=torch.uint8, device="cuda")
    K.hist_bulk(g, w[:2].contiguous(), torch.from_numpy(big).cuda(), cutoff, cb, within=flags)
    ref = np.stack([O.within_cutoff(water[f], big[f], cutoff) for f in range(2)])
    assert np.array_equal(flags.cpu().numpy().astype(bool), ref) and 0 < ref.sum() < ref.size
    assert np.array_equal(cb.cpu().numpy().astype(np.int64), O.histogram_bulk(water[:2], big, edges, cutoff))


@pytest.mark.parametrize("seed", range(8))
def test_workflow_random_systems(seed, tmp_path):
    """The whole class workflow on seeded random solvated systems (different sizes, grid spacings,
    cut-offs, solute sizes): DensityCreator(mode='all') -> DensityWithBulk -> HoppingTrajectory (file and
    fused) -> TransportNetwork, every product against the oracle."""
    import warnings

    from hop_b200.density import DensityCreator
    from hop_b200.graph import TransportNetwork
    from hop_b200.trajectory import HoppingTrajectory
    from hop_b200.universe import Universe

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(500 + seed)
    n_w = int(rng.choice([150, 333, 700, 1025]))
    n_f = int(rng.choice([3, 17, 40, 65]))
    n_p = int(rng.choice([1, 7, 40, 120]))
    delta = float(rng.choice([0.8, 1.0, 1.5]))
    cutoff = float(rng.choice([2.0, 3.5, 5.0]))
    p = SynthParams(n_atoms=n_w, n_frames=n_f, seed=900 + seed, site_stride=int(rng.choice([5, 20])))
    water = generate(p)
    c = p.box / 2
    prot0 = (c + rng.normal(0, 2.5, size=(n_p, 3))).astype(np.float32)
    prot = (prot0[None] + rng.normal(0, 0.3, size=(n_f, n_p, 3))).astype(np.float32)
    xyz = np.concatenate([prot, water], axis=1)
    u = Universe(coordinates=xyz, names=["CA"] * n_p + ["OH2"] * n_w, resnames=["ALA"] * n_p + ["TIP3"] * n_w, dt=1.5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        DC = DensityCreator(u, mode="all", delta=delta, atomselection="name OH2", cutoff=cutoff, soluteselection="protein")
        dens = DC.create()
        edges = dens["solvent"].edges
        _, _, ref_edges = O.grid_geometry(water[0], delta=delta)
        for a, b in zip(edges, ref_edges):
            assert np.array_equal(a, b)
        assert np.array_equal(dens["solvent"].grid, O.make_density(O.histogram_vectorised(water, edges), edges, n_f))
        assert np.array_equal(dens["bulk"].grid, O.make_density(O.histogram_bulk(water, prot, edges, cutoff), edges, n_f))
        final = DC.DensityWithBulk(density_unit="water", solvent_threshold=math.e, bulk_threshold=math.exp(-0.5))
        gs = O.make_density(O.histogram_vectorised(water, edges), edges, n_f, "water")
        gb = O.make_density(O.histogram_bulk(water, prot, edges, cutoff), edges, n_f, "water")
        bulk_sites = O.label_sites_ndimage(gb, math.exp(-0.5))
        if len(bulk_sites) > 1:
            sites = O.site_insert_bulk(O.label_sites_ndimage(gs, math.e), bulk_sites)
        else:
            sites = list(O.label_sites_ndimage(gs, math.e))
            sites.insert(1, [])
        site_map = O.draw_map_from_sites(sites, gs.shape)
        assert np.array_equal(final.map, site_map)
        group = u.select_atoms("name OH2")
        hop_ref = O.coord2hop_vectorised(water, edges, site_map)
        h = HoppingTrajectory(u.trajectory, group, final)
        h.write(str(tmp_path / "hoptraj"))
        got = np.stack([ts._pos.copy() for ts in HoppingTrajectory(filename=str(tmp_path / "hoptraj")).hoptraj])
        assert np.array_equal(got[:, :, 0], hop_ref[:, :, 0]) and np.array_equal(got[:, :, 1], hop_ref[:, :, 1])
        assert not got[:, :, 2].any()
        if n_f < 2 or len(sites) < 2:
            return
        props_ref, edges_ref, nodes_ref = O.analyze_hops_faithful(hop_ref[:, :, 0], dt=1.5)
        tn = TransportNetwork(h, final)
        tn.compute_residency_times()
        gp = tn.graph_properties
        assert set(gp.keys()) == set(props_ref.keys())
        for k, d in props_ref.items():
            for f in ("frame", "iatom"):
                assert np.array_equal(gp[k][f], d[f]), (k, f)
            np.testing.assert_allclose(gp[k]["tau"], d["tau"], rtol=1e-6)
        assert tn.transition_counts() == {e: len(props_ref[e]["tau"]) for e in edges_ref}


def test_gpu_rate_fit_follows_minpack():
    """hopb_rate_fit (survival functions + MINPACK's lmdif restated, one warp per edge) against scipy.optimize.leastsq,
    the reference's own call (hop/graph.py:2354-2419, 1677-1748): the survival functions are identical; the fits take
    the same optimiser path (nfev and info equal for the single exponentials, whose paths do not branch on rounding) and
    agree to 1e-6 for the well-determined ones -- lmdif differentiates by forward differences (h = 1.5e-8 |p|), so the
    last-bit differences of exp() reach any result at the 1e-8 .. 1e-6 level, and the parameters of a double
    exponential whose fast component decays within one mesh step are not determined by the data at all.
    HoppingGraph(rate_method='auto') therefore repeats the fits a perturbed second run moves with scipy itself."""
    import warnings

    import scipy.optimize

    from hop_b200 import _ratefit
    from hop_b200 import kernels as K
    from hop_b200.graph import _fit_edges

    rng = np.random.default_rng(0)
    taus = []
    for e in range(1200):
        n = int(rng.choice([1, 2, 5, 10, 11, 12, 20, 50, 200]))
        if rng.random() < 0.5:
            t = rng.exponential(rng.choice([2.0, 10.0, 50.0, 300.0]), n)
        else:
            t = np.where(rng.random(n) < 0.6, rng.exponential(3.0, n), rng.exponential(120.0, n))
        taus.append(np.floor(t) * float(rng.choice([1.0, 1.0, 2.5])))
    out, S, off_x = K.rate_fit(taus)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = [_ratefit.fit_one(t) for t in taus]
        single = [e for e, t in enumerate(taus) if len(t) <= 10]
        for e in single[:200]:
            x = np.arange(0, taus[e].max() + 2.0, 1.0)
            y = _ratefit.survival_values(taus[e], x)
            p, cov, info, msg, ier = scipy.optimize.leastsq(lambda p, x, y: np.exp(-p[0] * x) - y, [1.0], args=(x, y), full_output=True)
            assert info["nfev"] == int(out[e][6]) and ier == int(out[e][5])
    close = 0
    for e, (r, o) in enumerate(zip(ref, out)):
        x = np.arange(0, taus[e].max() + 2.0, 1.0)
        assert np.array_equal(S[off_x[e]:off_x[e + 1]], _ratefit.survival_values(taus[e], x))
        assert int(o[7]) == len(taus[e])
        same = (r[0] == 'exp2') == (o[0] == 1.0) and abs(o[4] - r[3]) <= 1e-6 * abs(r[3])
        close += same
        if len(taus[e]) <= 10:
            assert (o[0] == 0.0) and abs(o[4] - r[3]) <= 5e-6 * abs(r[3])
    assert close >= 0.9 * len(taus)
    # the default method of HoppingGraph: within 1e-6 of the reference's library call on (nearly) every edge
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        auto = _fit_edges(taus, "auto", workers=0)
    ok = sum(1 for a_, r in zip(auto, ref) if a_[0] == r[0] and abs(a_[3] - r[3]) <= 1e-6 * abs(r[3]))
    assert ok >= 0.995 * len(taus)
