"""Generates tests/golden/reference/*.npz by running the REFERENCE ITSELF (hop's own classes, loaded
from /root/reference by tests/golden/reference_py3.py) on seeded synthetic trajectories.

Run from the repo root, in the container that has /root/reference:

    python tests/golden/make_reference_golden.py

Every array in the files is an output of hop code: `DensityCreator(...).create()`
(hop/density.py:188-304), `Density.convert_density` / `map_sites` / `site_insert_bulk`
(hop/sitemap.py:236-264, 480-521, 877-918), `HoppingTrajectory.write` -> `map_dcd` -> `_coord2hop`
(hop/trajectory.py:255-297, 379-468), `TransportNetwork.compute_residency_times` ->
`_analyze_hops` + `HoppingGraph` (hop/graph.py:137-288, 1677-1748).  Nothing from oracle/ or hop_b200/
takes part except `hop_b200.synth.generate`, which only makes the input coordinates (stored in the
files).  The tests then hold the oracle (CPU, tests/test_reference_golden.py) and the CUDA path (GPU)
to these files.
"""
import hashlib
import math
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import reference_py3 as R  # noqa: E402
from hop_b200.synth import SynthParams, generate  # noqa: E402

OUT = os.path.join(HERE, "reference")


def run_reference(xyz_all, n_solvent, delta, density_unit, solvent_threshold, bulk_threshold,
                  cutoff=3.5, dt=1.0, rates=True, pickles_to=None):
    """The workflow of scripts/hop-generate-densities.py / -hoptraj.py / -hopgraph.py, through the
    reference's classes.  xyz_all float32 [F][n_solvent + n_solute][3]."""
    hop = R.load_reference()
    n_total = xyz_all.shape[1]
    sel = {"name OH2": np.arange(n_solvent), "protein": np.arange(n_solvent, n_total)}
    u = R.FakeUniverse(xyz_all, selections=sel, dt=dt)
    out = {}
    clock = {"t": time.perf_counter()}
    timing = {}

    def lap(stage):
        now = time.perf_counter()
        timing[stage] = timing.get(stage, 0.0) + now - clock["t"]
        clock["t"] = now
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if n_total > n_solvent:
            DC = hop.density.DensityCreator(u, mode="all", delta=delta, atomselection="name OH2",
                                            soluteselection="protein", cutoff=cutoff, padding=2.0)
            DC.create()
            out["bulk_grid_hist"] = DC.densities["bulk"].grid.copy()      # Angstrom^-3, before conversion
            density = DC.DensityWithBulk(density_unit=density_unit, solvent_threshold=solvent_threshold,
                                         bulk_threshold=bulk_threshold)
            out["bulk_map"] = DC.densities["bulk"].map.copy()
        else:
            DC = hop.density.DensityCreator(u, mode="solvent", delta=delta, atomselection="name OH2",
                                            padding=2.0)
            density = DC.create()["solvent"]
            lap("density")
            # no solute: the bulk density is the solvent histogram itself (what bench.py does)
            bulk = hop.sitemap.Density(grid=density.grid.copy(), edges=[e.copy() for e in density.edges],
                                       parameters={"isDensity": True},
                                       unit={"length": "Angstrom", "density": "Angstrom^{-3}"})
            density.convert_density(density_unit)
            density.map_sites(solvent_threshold)
            bulk.convert_density(density_unit)
            bulk.map_sites(bulk_threshold)
            density.site_insert_bulk(bulk)
            lap("map_sites")
            out["bulk_map"] = bulk.map.copy()
        out["edges"] = [np.asarray(e, dtype=np.float64) for e in density.edges]
        out["grid"] = np.asarray(density.grid, dtype=np.float64)
        out["map"] = np.asarray(density.map)
        assert out["map"].dtype == np.int16
        out["P"] = dict(density.P)
        sp = density.site_properties
        out["site_label"] = np.asarray(sp.label, dtype=np.int64)
        out["site_volume"] = np.asarray(sp.volume, dtype=np.float64)
        out["site_occupancy_avg"] = np.asarray(sp.occupancy_avg, dtype=np.float64)
        out["site_occupancy_std"] = np.asarray(sp.occupancy_std, dtype=np.float64)
        # the interstitial's centre is None (Python 2 / old numpy) or a scalar nan (numpy >= 1.x mean([]))
        out["site_center"] = np.array([np.full(3, np.nan) if (c is None or np.ndim(c) == 0)
                                       else np.asarray(c, dtype=np.float64) for c in sp.center])
        out["site_ncells"] = np.array([len(s) for s in density.sites], dtype=np.int64)

        group = u.select_atoms("name OH2")
        u.trajectory.rewind()
        lap("other")
        ht = hop.trajectory.HoppingTrajectory(u.trajectory, group, density, verbosity=0)
        with tempfile.TemporaryDirectory() as tmp:
            base = os.path.join(tmp, "refhop")
            ht.write(base)                       # map_dcd -> _coord2hop per frame -> (fake) DCD, reloaded
            hopframes = R.FakeUniverse._files[base + ".dcd"].xyz
            with open(base + ".psf") as fh:      # the PSF is written to disk by hop itself
                out["psf_text"] = fh.read()
        lap("hoptraj")
        out["hop_cell"] = np.asarray(ht._ts.dimensions if ht._ts is not None else [], dtype=np.float64) \
            if hasattr(ht, "_ts") else np.zeros(0)
        out["hop"] = hopframes.copy()            # float32 [F][N][3]: x=site, y=orbit, z=0
        tn = hop.graph.TransportNetwork(ht, density)
        if rates:
            tn.compute_residency_times(verbosity=0)       # _analyze_hops + HoppingGraph (rate fits)
        else:
            tn._analyze_hops()
            lap("analyze_hops")
        gp = tn.graph_properties
        ev, nodes = [], []
        for key in sorted((k for k in gp if isinstance(k, tuple)), key=lambda k: (int(k[0]), int(k[1]))):
            d = gp[key]
            for j in range(len(d["tau"])):
                ev.append((int(key[0]), int(key[1]), float(d["tau"][j]), float(d["tbarrier"][j]),
                           int(d["frame"][j]), int(d["iatom"][j])))
        for key in sorted(int(k) for k in gp if not isinstance(k, tuple)):
            d = gp[key]
            for j in range(len(d["tau"])):
                nodes.append((key, float(d["tau"][j]), int(d["frame"][j]), int(d["iatom"][j])))
        out["events"] = np.array(ev, dtype=np.float64).reshape(-1, 6)
        out["nodes"] = np.array(nodes, dtype=np.float64).reshape(-1, 4)
        out["graph_edges"] = np.array(sorted((int(a), int(b)) for a, b in tn.graph.edges()),
                                      dtype=np.int64).reshape(-1, 2)
        # graph_alltransitions (hop/graph.py:107-135)
        tn2 = hop.graph.TransportNetwork(ht, density)
        tn2.graph_alltransitions()
        out["alltransitions"] = np.array(sorted((int(a), int(b)) for a, b in tn2.graph.edges()),
                                         dtype=np.int64).reshape(-1, 2)
        if rates and pickles_to:
            # hop's own on-disk products: Saveable.save (hop/utilities.py:216-222), HoppingGraph.save
            # (hop/graph.py:760-767) -- pickles of the instances, protocol 2
            os.makedirs(pickles_to, exist_ok=True)
            density.save(os.path.join(pickles_to, "density"))
            # host-side selections on that very density, for a CPU-only check of the drop-in's site_labels / stats
            combos = {}
            for i, (inc, exc) in enumerate(SITE_LABEL_COMBOS):
                try:
                    combos["labels_%02d" % i] = np.asarray(density.site_labels(include=inc, exclude=exc), dtype=np.int64)
                except Exception as exc_:                      # noqa: BLE001 - record what hop raises
                    combos["labels_%02d" % i] = np.array([-999])
                    combos["error_%02d" % i] = np.array(type(exc_).__name__)
            st = density.stats()
            np.savez_compressed(os.path.join(pickles_to, "site_labels.npz"),
                                stats_keys=np.array(sorted(st)), stats_values=np.array([float(st[k]) for k in sorted(st)]),
                                **combos)
            tn.hopgraph.save(os.path.join(pickles_to, "hopgraph"))
        if rates:
            hg = tn.hopgraph
            rows = []
            for (a, b) in sorted(hg.graph.edges()):
                d = hg.graph[a][b]
                k = d.get("k", np.nan)
                rows.append((int(a), int(b), float(k), float(d.get("N", np.nan))))
            out["rates"] = np.array(rows, dtype=np.float64).reshape(-1, 4)
    out["timing_s"] = np.array([timing.get(k, 0.0) for k in ("density", "map_sites", "hoptraj", "analyze_hops")])
    return out


def save(name, xyz_all, n_solvent, store_xyz=True, extra=None, **kw):
    res = run_reference(xyz_all, n_solvent, **kw)
    res.update(extra or {})
    res["xyz_sha256"] = hashlib.sha256(np.ascontiguousarray(xyz_all).tobytes()).hexdigest()
    if not store_xyz:
        xyz_all = np.zeros((0, xyz_all.shape[1], 3), np.float32)
    P = res.pop("P")
    edges = res.pop("edges")
    os.makedirs(OUT, exist_ok=True)
    hop = res.pop("hop")
    assert np.array_equal(hop, np.round(hop)) and np.all(hop[:, :, 2] == 0)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), xyz=xyz_all, n_solvent=n_solvent,
        delta=kw["delta"], density_unit=kw["density_unit"], solvent_threshold=kw["solvent_threshold"],
        bulk_threshold=kw["bulk_threshold"], cutoff=kw.get("cutoff", 3.5), dt=kw.get("dt", 1.0),
        edges_x=edges[0], edges_y=edges[1], edges_z=edges[2],
        hop_x=hop[:, :, 0].astype(np.int16), hop_y=hop[:, :, 1].astype(np.int16),
        P_threshold=P["threshold"], P_bulk_threshold=P["bulk_threshold"], P_bulk_site=P["bulk_site"],
        numpy_version=np.__version__, **res)
    print("%-16s grid %s sites %d events %d nodes %d outlier atom-frames %d rate edges %d  reference seconds "
          "(density, map_sites, hoptraj, analyze_hops) %s" % (
              name, res["grid"].shape, len(res["site_ncells"]) - 1, len(res["events"]), len(res["nodes"]),
              int((hop[:, :, 0] == -1).sum()), len(res.get("rates", [])), np.round(res["timing_s"], 3)))


def sitemap_variants(name, xyz, delta, unit):
    """Secondary behaviours of the site map, all from hop's Density (hop/sitemap.py): MINsite = 2
    (:1526-1527), map_sites on a density that already carries a bulk site (:506-514),
    site_insert_nobulk (:933-943), site_labels filters (:629-760), site_volume / site_occupancy
    (:563-598), stats (:1301-1339)."""
    hop = R.load_reference()
    water = 1.0 / (1e-24 * 6.02214179e+23 * 0.997 / 18.016)
    t_solvent = math.e if unit == "water" else math.e / water
    t_bulk = math.exp(-0.5) if unit == "water" else math.exp(-0.5) / water
    u = R.FakeUniverse(xyz, selections={"name OH2": np.arange(xyz.shape[1])})
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        def fresh(**parameters):
            u.trajectory.rewind()                 # the grid is defined by the frame the universe is on
            DC = hop.density.DensityCreator(u, mode="solvent", delta=delta, atomselection="name OH2", padding=2.0,
                                            parameters=dict(parameters))
            d = DC.create()["solvent"]
            d.convert_density(unit)
            return d
        d2 = fresh(MINsite=2)
        d2.map_sites(t_solvent)
        out["minsite2_map"] = d2.map.copy()
        out["minsite2_ncells"] = np.array([len(x) for x in d2.sites])
        d1 = fresh()
        d1.map_sites(t_solvent)
        out["minsite1_map"] = d1.map.copy()
        bulk = fresh()
        bulk.map_sites(t_bulk)
        d1.site_insert_bulk(bulk)
        out["with_bulk_map"] = d1.map.copy()
        for key, kw in (("default", {}), ("sites", dict(include="sites")), ("all_none", dict(include="all", exclude=None)),
                        ("no_bulk", dict(exclude=["default", "bulk"])), ("explicit", dict(include=[2, 3, 5], exclude=None))):
            out["labels_" + key] = np.asarray(d1.site_labels(**kw), dtype=np.int64)
        labels, vol = d1.site_volume(include="sites")
        out["site_volume_labels"], out["site_volume"] = np.asarray(labels), np.asarray(vol, dtype=np.float64)
        labels, occ, occ_std = d1.site_occupancy(include="sites")
        out["site_occupancy"] = np.asarray(occ, dtype=np.float64)
        out["site_occupancy_std"] = np.asarray(occ_std, dtype=np.float64)
        st = d1.stats()
        out["stats_keys"] = np.array(sorted(st))
        out["stats_values"] = np.array([float(st[k]) for k in sorted(st)])
        # re-mapping at a higher threshold removes the bulk site first (and forgets it)
        d1.map_sites(1.5 * t_solvent)
        out["remap_map"] = d1.map.copy()
        out["remap_has_bulk"] = bool(d1.has_bulk())
        out["remap_P_keys"] = np.array(sorted(d1.P.keys()))
        d1.site_insert_nobulk()
        out["nobulk_map"] = d1.map.copy()
        out["nobulk_ncells"] = np.array([len(x) for x in d1.sites])
        out["nobulk_has_bulk"] = bool(d1.has_bulk())
        out["nobulk_bulk_threshold_is_none"] = d1.P["bulk_threshold"] is None
    os.makedirs(os.path.join(OUT, "variants"), exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "variants", name + ".npz"), xyz=xyz, delta=delta, density_unit=unit,
                        solvent_threshold=t_solvent, bulk_threshold=t_bulk, **out)
    print("%-16s MINsite=1 sites %d, MINsite=2 sites %d, remapped sites %d" % (
        name, out["minsite1_map"].max(), len(out["minsite2_ncells"]) - 1, len(out["nobulk_ncells"]) - 2))


#: (include, exclude) arguments of Density.site_labels (hop/sitemap.py:629-760) frozen for the CPU-only test
SITE_LABEL_COMBOS = [
    ("default", "default"), ("all", "default"), ("all", None), ("sites", "default"), ("sites", None),
    ("default", None), ("default", ["default", "bulk"]), ("default", "bulk"), ("default", "interstitial"),
    (3, "default"), ([2, 5, 7], "default"), ([0, 1, 2], None), ([0, 1, 2], "default"), ([1, 4], "bulk"),
    ("subsites", None), ("equivalencesites", None), (["sites", 1], None), ("all", ["bulk", "interstitial"]),
]


def synth(n_atoms, n_frames, seed=None):
    p = SynthParams(n_atoms=n_atoms, n_frames=n_frames) if seed is None else \
        SynthParams(n_atoms=n_atoms, n_frames=n_frames, seed=seed)
    return generate(p)


def siteanalysis_golden(name, xyz, delta=1.0):
    """hop.siteanalysis.Collection with the Distance and Orbit observables (hop/siteanalysis.py:722-802, 866-1107)
    on a trajectory with outliers.  compute() stops one frame short of the end: its frame iterator advances both
    readers once more after the last frame and relies on the IOError old MDAnalysis readers raised there (:1173-1183).
    Also records that the Occupancy observable cannot run (see tests/test_reference_golden.py)."""
    hop = R.load_reference()
    SA = R.load_siteanalysis()
    n = xyz.shape[1]
    u = R.FakeUniverse(xyz, selections={"name OH2": np.arange(n)})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        density = hop.density.DensityCreator(u, mode="solvent", delta=delta, atomselection="name OH2", padding=2.0).create()["solvent"]
        bulk = hop.sitemap.Density(grid=density.grid.copy(), edges=[e.copy() for e in density.edges],
                                   parameters={"isDensity": True}, unit={"length": "Angstrom", "density": "Angstrom^{-3}"})
        density.convert_density("water")
        density.map_sites(math.e)
        bulk.convert_density("water")
        bulk.map_sites(math.exp(-0.5))
        density.site_insert_bulk(bulk)
        group = u.select_atoms("name OH2")
        u.trajectory.rewind()
        ht = hop.trajectory.HoppingTrajectory(u.trajectory, group, density, verbosity=0)
        with tempfile.TemporaryDirectory() as tmp:
            ht.write(os.path.join(tmp, "refhop"))
            hopframes = R.FakeUniverse._files[os.path.join(tmp, "refhop") + ".dcd"].xyz.copy()
        ht.group = R.LegacySelection(ht.group)
        u.trajectory.rewind()
        out = {}
        for tag, kw in (("default", {}), ("with_bulk", {"exclude_sites": ["default"]})):
            C = SA.Collection(selection=R.LegacySelection(group), hoptraj=ht, density=density, **kw)
            C.add_observable("distance", lo_bin=0, hi_bin=3)
            C.add_observable("orbit", lo_bin=0, hi_bin=6, Nbins=24)
            u.trajectory.rewind()
            C.compute(stop=xyz.shape[0] - 1)
            for obs in ("distance", "orbit"):
                H = C.SiteHistograms[obs]
                out["%s_%s_histogram" % (tag, obs)] = np.asarray(H._histogram, dtype=np.int64)
                out["%s_%s_edges" % (tag, obs)] = np.asarray(H.edges, dtype=np.float64)
                out["%s_%s_distribution" % (tag, obs)] = np.asarray(H.distribution, dtype=np.float64)
                out["%s_%s_mean" % (tag, obs)] = np.asarray(H.mean(), dtype=np.float64)
            out["%s_sites" % tag] = np.asarray(C.siteselection, dtype=np.int64)
        C = SA.Collection(selection=R.LegacySelection(group), hoptraj=ht, density=density)
        C.add_observable("occupancy", lo_bin=0, hi_bin=5)
        try:
            u.trajectory.rewind()
            C.compute(stop=xyz.shape[0] - 1)
            out["occupancy_error"] = np.array("")
        except Exception as exc:                                   # noqa: BLE001 - record what hop raises
            out["occupancy_error"] = np.array("%s: %s" % (type(exc).__name__, exc))
    os.makedirs(os.path.join(OUT, "siteanalysis"), exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "siteanalysis", name + ".npz"), xyz=xyz, delta=delta, stop=xyz.shape[0] - 1,
                        map=np.asarray(density.map), hop_x=hopframes[:, :, 0].astype(np.int16),
                        hop_y=hopframes[:, :, 1].astype(np.int16),
                        site_center=np.array([np.full(3, np.nan) if (c is None or np.ndim(c) == 0) else np.asarray(c, float)
                                              for c in density.site_properties.center]), **out)
    print("%-16s sites %d  distance counts %d  orbit counts %d  occupancy: %s" % (
        name, len(density.site_properties) - 1, int(out["default_distance_histogram"].sum()),
        int(out["default_orbit_histogram"].sum()), str(out["occupancy_error"])[:70]))


def main():
    water = 1.0 / (1e-24 * 6.02214179e+23 * 0.997 / 18.016)      # Angstrom^-3 -> 'water' units
    # 1. clean box, thresholds in Angstrom^-3 (no unit table involved)
    xyz = synth(300, 60)
    save("ref_clean", xyz, 300, delta=1.0, density_unit="Angstrom^{-3}",
         solvent_threshold=math.e / water, bulk_threshold=math.exp(-0.5) / water,
         pickles_to=os.path.join(OUT, "pickles"))
    # 2. outliers (also in the first frame: the (-1, N) quirk), hop's default 'water' unit
    xyz = synth(300, 60)
    rng = np.random.default_rng(7)
    m = rng.random(xyz.shape[:2]) < 0.03
    m[0, :] = False
    xyz[m] += np.float32(50.0)
    xyz[0, 5:9] -= np.float32(60.0)
    save("ref_outliers", xyz, 300, delta=1.0, density_unit="water",
         solvent_threshold=math.e, bulk_threshold=math.exp(-0.5))
    # 3. half-Angstrom grid, dt != 1
    xyz = synth(400, 40)
    save("ref_half", xyz, 400, delta=0.5, density_unit="water",
         solvent_threshold=math.e, bulk_threshold=math.exp(-0.5), dt=2.5)
    # 4. solute present: DensityCreator(mode='all') with the 'not within cutoff of solute' bulk selection
    xyz = synth(300, 50, seed=20261023)
    L = float(xyz[0].max())
    rng = np.random.default_rng(11)
    solute0 = (0.5 * L + rng.normal(0.0, 3.0, size=(24, 3))).astype(np.float32)
    solute = solute0[None] + rng.normal(0.0, 0.2, size=(xyz.shape[0], 24, 3)).astype(np.float32)
    save("ref_solute", np.concatenate([xyz, solute], axis=1), 300, delta=1.0, density_unit="water",
         solvent_threshold=math.e, bulk_threshold=math.exp(-0.5), cutoff=3.5)
    # 5. BASELINE.json configs[0]: 1 000 waters x 200 frames, 1.0 A grid -- "reference runs as-is".  The
    # coordinates are the bench generator's (seed 20261019 + 1, as tools/config_sweep.py); only their
    # hash is stored, the tests regenerate them.
    p = SynthParams(n_atoms=1000, n_frames=200, seed=20261019 + 1)
    xyz = generate(p)
    save("ref_config1", xyz, 1000, store_xyz=False, rates=False, extra={"synth_seed": p.seed},
         delta=1.0, density_unit="water", solvent_threshold=math.e, bulk_threshold=math.exp(-0.5))


if __name__ == "__main__":
    main()
    sitemap_variants("variants_clean", synth(300, 60), 1.0, "water")
    xyz = synth(300, 60)
    m = np.random.default_rng(3).random(xyz.shape[:2]) < 0.02
    m[0, :] = False
    xyz[m] += np.float32(50.0)
    siteanalysis_golden("siteanalysis", xyz)
