"""CPU oracle for the hop hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module.  ``hop_b200`` never does; the product path is CUDA only.

This is a numpy / networkx restatement of the reference algorithm (Becksteinlab/hop 0.4.0-dev,
pure Python 2.7), each function citing the file:line it restates.  The reference ships no runnable
tests or golden vectors (hop/tests/README:1-4) and cannot be imported as it is (Python 2 syntax,
MDAnalysis absent), so the oracle is **pinned against outputs of the reference itself, run here**:
``tests/golden/reference_py3.py`` executes hop's own sources from /root/reference under Python 3
(mechanical in-memory syntax translation, Python-2 builtins and classic division restored, three
numpy-API one-liners, MDAnalysis replaced by in-memory fakes) and
``tests/golden/make_reference_golden.py`` froze what ``DensityCreator``, ``Density.map_sites`` /
``site_insert_bulk``, ``HoppingTrajectory.write``, ``TransportNetwork`` and ``HoppingGraph`` produce
on four seeded systems into ``tests/golden/reference/*.npz``.  ``tests/test_reference_golden.py``
holds both restatements below to those files: edges and densities bit-exact, site map / hoptraj /
events / node records / transition counts exact (up to the numbering of equal-volume sites, which
the reference leaves to set order), site properties and rate constants to 1e-6.  What that run does
NOT pin: the unit constants of MDAnalysis.units (the fakes carry the same values as here), the
kd-tree selection of MDAnalysis (restated), and the float32 top-edge rule of numpy 1.x (below).
Beyond that the oracle is cross-checked by (i) two independent restatements per stage (a "faithful"
one that repeats the reference's loops and a vectorised one) that must agree, and (ii)
`scipy.ndimage.label` with the 15-cell structuring element as a third opinion on the connected
components.

Third-party arithmetic the reference delegates to and that is restated here from the published
behaviour (the packages' sources are not under /root/reference):

* ``numpy.histogramdd`` (numpy>=1.12, setup.py:39): the *installed* numpy (2.3.x) is called
  directly in faithful mode, i.e. ``searchsorted(edges, x, 'right')`` plus the exact
  ``x == edges[-1]`` fix-up.
* ``numpy.digitize`` / ``numpy.around`` in ``_coord2hop``: the top-edge comparison is done in
  **float32**, which is what numpy 1.x value-based casting did for ``float32_array == float64_scalar``
  (the reference era).
* ``MDAnalysis.units`` (MDAnalysis>=0.18.0, setup.py:42): constants pinned below from the
  0.18 release (from memory -- not verifiable offline).

Conventions: labels outlier=-1, interstitial=0, bulk=1 (hop/constants.py:70).
"""
from __future__ import annotations

import math

import numpy as np

SITELABEL = dict(outlier=-1, interstitial=0, bulk=1)  # hop/constants.py:70

# --- MDAnalysis.units (0.18) restated; used via hop/constants.py:45,63 --------------------------
N_AVOGADRO = 6.02214179e+23
WATER = {"exp": 0.997, "SPC": 0.985, "TIP3P": 1.002, "TIP4P": 1.001, "MolarMass": 18.016}
DENSITY_UNIT_FACTOR = {
    "Angstrom^{-3}": 1 / 1.0,
    "A^{-3}": 1 / 1.0,
    "nm^{-3}": 1 / 1e-3,
    "nanometer^{-3}": 1 / 1e-3,
    "Molar": 1 / (1e-27 * N_AVOGADRO),
    "SPC": 1 / (1e-24 * N_AVOGADRO * WATER["SPC"] / WATER["MolarMass"]),
    "TIP3P": 1 / (1e-24 * N_AVOGADRO * WATER["TIP3P"] / WATER["MolarMass"]),
    "TIP4P": 1 / (1e-24 * N_AVOGADRO * WATER["TIP4P"] / WATER["MolarMass"]),
    "water": 1 / (1e-24 * N_AVOGADRO * WATER["exp"] / WATER["MolarMass"]),
}


def density_conversion_factor(u1, u2):
    """MDAnalysis.units.get_conversion_factor('density', u1, u2) = factor[u2]/factor[u1]."""
    return DENSITY_UNIT_FACTOR[u2] / DENSITY_UNIT_FACTOR[u1]


# the 7 forward offsets, hop/sitemap.py:415-422
DELTA_FIRST_OCTANT = np.array(
    [[0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 1, 1], [1, 0, 1], [1, 1, 0], [1, 1, 1]]
)


# =================================================================================================
# Stage 1: grid geometry and histogram
# =================================================================================================
def fixedwidth_bins(delta, xmin, xmax):
    """hop/utilities.py:373-389."""
    if not np.all(xmin < xmax):
        raise ValueError("Boundaries are not sane: should be xmin < xmax.")
    _delta = np.asarray(delta, dtype=np.float64)
    _xmin = np.asarray(xmin, dtype=np.float64)
    _xmax = np.asarray(xmax, dtype=np.float64)
    _length = _xmax - _xmin
    N = np.ceil(_length / _delta).astype(np.int64)
    dx = 0.5 * (N * _delta - _length)
    return {"Nbins": N, "delta": _delta, "min": _xmin - dx, "max": _xmax + dx}


def grid_geometry(coords0, delta=1.0, padding=2.0):
    """Grid from the frame-0 coordinates of ONE collector.

    hop/density.py:119-123 (min/max -/+ padding, evaluated in the coordinates' float32),
    271-278 (sort over collectors; with one collector it is the identity), 99-117
    (fixedwidth_bins, then an empty histogramdd to obtain the edges).
    Returns (bins int64[3], arange list[(lo,hi)], edges list[f64 array]).
    """
    coords0 = np.asarray(coords0)
    smin = np.min(coords0, axis=0) - padding  # float32 arithmetic when coords are float32
    smax = np.max(coords0, axis=0) + padding
    BINS = fixedwidth_bins(delta, smin, smax)
    arange = list(zip(BINS["min"], BINS["max"]))
    bins = BINS["Nbins"]
    _, edges = np.histogramdd(np.zeros((1, 3)), bins=bins, range=arange)
    return bins, arange, [np.asarray(e, dtype=np.float64) for e in edges]


def histogram_faithful(frames, bins, arange):
    """Per-frame numpy.histogramdd + `grid += h` (hop/density.py:125-131).

    `frames` is an iterable of (N,3) float32 arrays.  Returns (grid f64 counts, n_frames)."""
    grid = None
    n = 0
    for coord in frames:
        n += 1
        if len(coord) > 0:
            h, _ = np.histogramdd(coord, bins=bins, range=arange)
            grid = h if grid is None else grid + h
    if grid is None:
        grid = np.zeros(tuple(int(b) for b in bins))
    return grid, n


def bin_right(x, edges):
    """searchsorted(edges, float64(x), 'right') == numpy.digitize(x, edges) for increasing edges."""
    return np.searchsorted(edges, np.asarray(x, dtype=np.float64), side="right")


def histogram_vectorised(xyz, edges):
    """Same counts as histogram_faithful, all frames at once (int64 counts).

    numpy.histogramdd semantics (numpy >= 1.15): bin = searchsorted(edges, x, 'right');
    x == edges[-1] goes to the last bin; anything outside is dropped."""
    pts = np.asarray(xyz).reshape(-1, 3)
    shape = tuple(len(e) - 1 for e in edges)
    idx = []
    ok = np.ones(len(pts), dtype=bool)
    for d in range(3):
        x = pts[:, d].astype(np.float64)
        b = np.searchsorted(edges[d], x, side="right")
        b[x == edges[d][-1]] -= 1
        ok &= (b >= 1) & (b <= shape[d])
        idx.append(b - 1)
    lin = np.ravel_multi_index([i[ok] for i in idx], shape)
    return np.bincount(lin, minlength=int(np.prod(shape))).reshape(shape).astype(np.int64)


def within_cutoff(solvent, solute, cutoff):
    """Boolean [Nw]: a solute atom lies within `cutoff` of the solvent atom (one frame).

    Restates the selection '<sel1> not within <cutoff> of <sel2>' that hop builds with MDAnalysis'
    kd-tree (`notwithin_coordinates_factory`, hop/density.py:82-87; MDAnalysis is absent, so this is
    unpinned): plain Euclidean distance, no periodic images, boundary included.  The squared
    distance is evaluated in float32 as (dx*dx + dy*dy) + dz*dz -- the order the kernel uses."""
    w = np.asarray(solvent, dtype=np.float32)
    p = np.asarray(solute, dtype=np.float32)
    c2 = np.float32(cutoff) * np.float32(cutoff)
    out = np.zeros(len(w), dtype=bool)
    for j in range(len(p)):   # chunk-free and exact; Np is small in the tests
        d = w - p[j]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        out |= d2 <= c2
    return out


def histogram_bulk(xyz_solvent, xyz_solute, edges, cutoff):
    """Counts of the solvent atoms that are NOT within `cutoff` of the solute, frame by frame
    (DensityCollector in BULK mode, hop/density.py:82-87 + 125-131)."""
    shape = tuple(len(e) - 1 for e in edges)
    total = np.zeros(shape, np.int64)
    for w, p in zip(xyz_solvent, xyz_solute):
        keep = ~within_cutoff(w, p, cutoff)
        if keep.any():
            total += histogram_vectorised(w[keep][None], edges)
    return total


def make_density(counts, edges, n_frames, unit=None):
    """counts -> density, reproducing the reference's operation order exactly.

    hop/density.py:133-137 (`grid /= float(n_frames)`), hop/sitemap.py:196-216 (`grid /=
    diff(edges[d])` for d = 0,1,2 in turn), hop/sitemap.py:236-264 (`grid *= factor`)."""
    grid = np.array(counts, dtype=np.float64)
    grid /= float(n_frames)
    for d in range(3):
        shape = np.ones(3, int)
        de = np.diff(edges[d])
        shape[d] = len(de)
        grid /= de.reshape(shape)
    if unit is not None and unit != "Angstrom^{-3}":
        grid *= density_conversion_factor("Angstrom^{-3}", unit)
    return grid


# =================================================================================================
# Stage 2: site labelling
# =================================================================================================
def _rank_components(components):
    """Size-ranked canonical numbering (hop/sitemap.py:1560-1569 + SURVEY.md F4).

    The reference sorts `(volume, list(component))` ascending and reverses; list(component) is
    in CPython set order, so ties are not deterministic there.  Canonical choice: component
    lists sorted ascending, which makes the comparison (volume, min cell) -> after the reverse,
    volume descending then minimum C-order cell descending."""
    sites = [sorted(c) for c in components]
    vs = sorted(zip([len(s) for s in sites], sites))
    vs.reverse()
    sites = [s for _, s in vs]
    sites.insert(SITELABEL["interstitial"], [])
    return sites


def label_sites_faithful(grid, threshold, minsite=1):
    """Density.map_sites as written: networkx graph built with Python loops.

    hop/sitemap.py:518-521 (mask), 1496-1527 (_make_graph), 1530-1541 (_shell),
    1543-1571 (_label_connected_graphs).  Returns `sites` (list of lists of (i,j,k))."""
    import networkx as NX

    grid = np.asarray(grid)
    mask_map = np.where(grid >= threshold, 1, SITELABEL["interstitial"]).astype(np.int16)
    graph = NX.Graph()
    sidx = list(map(tuple, np.array(np.where(grid >= threshold)).T.tolist()))
    sidx.sort()
    for site in sidx:
        for neighbour in map(tuple, (np.asarray(site) + DELTA_FIRST_OCTANT).tolist()):
            try:
                if mask_map[neighbour]:
                    graph.add_edge(site, neighbour)
            except IndexError:
                pass
    if minsite == 1:
        graph.add_nodes_from(sidx)
    return _rank_components(NX.connected_components(graph))


def label_sites_ndimage(grid, threshold, minsite=1):
    """Independent opinion: scipy.ndimage.label with the 15-cell structuring element
    (centre + the 7 forward offsets + their negatives); same canonical numbering."""
    from scipy import ndimage

    grid = np.asarray(grid)
    mask = grid >= threshold
    st = np.zeros((3, 3, 3), dtype=bool)
    st[1, 1, 1] = True
    for o in DELTA_FIRST_OCTANT:
        st[tuple(1 + o)] = True
        st[tuple(1 - o)] = True
    lab, n = ndimage.label(mask, structure=st)
    flat = lab.ravel()
    order = np.argsort(flat, kind="stable")
    counts = np.bincount(flat, minlength=n + 1)
    starts = np.cumsum(counts) - counts
    comps = []
    for l in range(1, n + 1):
        cells = order[starts[l]:starts[l] + counts[l]]
        if minsite == 2 and len(cells) < 2:
            continue
        comps.append(list(map(tuple, np.array(np.unravel_index(cells, grid.shape)).T.tolist())))
    return _rank_components(comps)


def draw_map_from_sites(sites, shape):
    """hop/sitemap.py:1574-1587: later labels overwrite earlier ones; int16."""
    m = SITELABEL["interstitial"] * np.ones(shape, dtype=np.int16)
    for label, sidx in enumerate(sites):
        if label > np.iinfo(np.int16).max:
            raise OverflowError("more than 32767 sites cannot be represented in the int16 map")
        if len(sidx):
            a = np.asarray(sidx)
            m[a[:, 0], a[:, 1], a[:, 2]] = label
    return m


def site_insert_bulk(sites, bulk_sites, bulklabel=SITELABEL["bulk"]):
    """hop/sitemap.py:914: `sites.insert(1, bulkdensity.sites[bulklabel])`."""
    out = list(sites)
    out.insert(SITELABEL["bulk"], list(bulk_sites[bulklabel]))
    return out


def site_properties(sites, grid, edges, density_unit="Angstrom^{-3}"):
    """_annotate_sites: volume, occupancy avg/std, centre (hop/sitemap.py:1589-1612, 601-627).

    Returns dict of arrays indexed by label; centre is NaN for empty sites (None there)."""
    delta = np.diag([(e[-1] - e[0]) / (len(e) - 1) for e in edges])  # sitemap.py:161-162
    midpoints = [0.5 * (e[:-1] + e[1:]) for e in edges]
    det = np.linalg.det(delta)
    factor = density_conversion_factor(density_unit, "Angstrom^{-3}")
    Vcell = factor * det
    n = len(sites)
    volume = np.zeros(n)
    occ_avg = np.zeros(n)
    occ_std = np.zeros(n)
    center = np.full((n, 3), np.nan)
    for l, site in enumerate(sites):
        volume[l] = det * len(site)
        V = len(site) * Vcell
        if V == 0:
            continue
        a = np.asarray(site)
        dens = np.array([grid[tuple(s)] for s in site]) if len(site) < 64 else grid[a[:, 0], a[:, 1], a[:, 2]]
        occ_avg[l] = np.average(dens) * V
        occ_std[l] = np.std(dens) * V
        center[l] = np.mean(
            np.stack([midpoints[0][a[:, 0]], midpoints[1][a[:, 1]], midpoints[2][a[:, 2]]], axis=1), axis=0
        )
    return dict(label=np.arange(n), volume=volume, occupancy_avg=occ_avg, occupancy_std=occ_std, center=center)


def map_sites(grid, threshold, minsite=1, faithful=False):
    """Convenience: (sites, map int16) for one density."""
    sites = (label_sites_faithful if faithful else label_sites_ndimage)(grid, threshold, minsite)
    return sites, draw_map_from_sites(sites, np.asarray(grid).shape)


# =================================================================================================
# Stage 3: hopping trajectory
# =================================================================================================
def top_edge_decimals(edges):
    """`int(-log10(dedges.min())) + 6` per axis (hop/trajectory.py:435)."""
    return [int(-np.log10(np.diff(e).min())) + 6 for e in edges]


def _around32(x32, decimals):
    """numpy.around on a float32 array: multiply, rint, divide, all in float32."""
    return np.around(np.asarray(x32, dtype=np.float32), decimals)


def digitize_hop(coords, edges):
    """Per-axis bin index in [0, N_d+1] with the reference's top-edge rule applied.

    hop/trajectory.py:428 (numpy.digitize) and 433-441 (rounded equality, evaluated in float32 as
    numpy 1.x did for float32_array == float64_scalar)."""
    coords = np.asarray(coords)
    decs = top_edge_decimals(edges)
    out = []
    for d in range(3):
        x = coords[:, d]
        idx = np.digitize(x, edges[d])
        top = np.float32(np.around(edges[d][-1], decs[d]))
        on_edge = _around32(x, decs[d]) == top
        idx = idx.copy()
        idx[on_edge] -= 1
        out.append(idx)
    return out


def coord2hop_faithful(frames, edges, site_map):
    """HoppingTrajectory._coord2hop over a whole pass (hop/trajectory.py:396-468).

    Per frame the map is copied into the -1 padded buffer and looked up per atom with a Python
    list comprehension, exactly like the reference.  Returns float32 [F][N][3] (x=site, y=orbit,
    z=0)."""
    site_map = np.asarray(site_map)
    buffered_map = SITELABEL["outlier"] * np.ones(tuple(np.asarray(site_map.shape) + 2))
    out = []
    sites_last = None
    for coords in frames:
        coords = np.asarray(coords)
        N = coords.shape[0]
        if sites_last is None:
            sites_last = SITELABEL["interstitial"] * np.ones(N)  # trajectory.py:399
        indices = digitize_hop(coords, edges)
        buffered_map[1:-1, 1:-1, 1:-1] = site_map
        pos = np.empty((N, 3), dtype=np.float32)
        pos[:, 0] = [buffered_map[indices[0][i], indices[1][i], indices[2][i]] for i in range(N)]
        s = pos[:, 0]
        offsites = (s == SITELABEL["interstitial"]) | (s == SITELABEL["outlier"])
        pos[:, 1] = s
        pos[offsites, 1] = sites_last[offsites]
        pos[:, 2] = 0
        sites_last[:] = pos[:, 1]
        out.append(pos)
    return np.stack(out) if out else np.zeros((0, 0, 3), np.float32)


def coord2hop_vectorised(xyz, edges, site_map):
    """Same result as coord2hop_faithful, vectorised over atoms (frames still sequential for the
    orbit column)."""
    xyz = np.asarray(xyz)
    F, N, _ = xyz.shape
    site_map = np.asarray(site_map)
    buffered = np.full(tuple(np.asarray(site_map.shape) + 2), SITELABEL["outlier"], dtype=np.int32)
    buffered[1:-1, 1:-1, 1:-1] = site_map
    out = np.zeros((F, N, 3), dtype=np.float32)
    last = np.zeros(N, dtype=np.int32)
    for f in range(F):
        ix, iy, iz = digitize_hop(xyz[f], edges)
        s = buffered[ix, iy, iz]
        off = (s == 0) | (s == -1)
        y = np.where(off, last, s)
        out[f, :, 0] = s
        out[f, :, 1] = y
        last = y
    return out


# =================================================================================================
# Stage 4: transition scan
# =================================================================================================
def analyze_hops_faithful(hop_x, dt=1.0):
    """Transcription of TransportNetwork._analyze_hops (hop/graph.py:190-285).

    `hop_x` is the x column of the hopping trajectory, [F][N] (float32 or int).  Frame numbers
    are 0-based; frame 0 initialises the per-atom state, the loop covers frames 1..F-1.
    Returns (graph_properties, edges, nodes): graph_properties maps an edge tuple (i,j) or a node
    int to {'tau','tbarrier','frame','iatom'} arrays in emission order (frame-major, then atom)."""
    hop_x = np.asarray(hop_x)
    F, n_atoms = hop_x.shape
    props = {}
    edges = []
    nodes = []
    s = hop_x[0].astype(int)
    slast = s.copy()
    state = [{"s0": int(s[i]), "t0": 0, "t1": None} for i in range(n_atoms)]
    frame = 0
    for frame in range(1, F):
        s = hop_x[frame].astype(int)
        for iatom in range(n_atoms):
            if s[iatom] == SITELABEL["outlier"]:
                s[iatom] = slast[iatom]
            st = state[iatom]
            if s[iatom] != st["s0"]:
                if slast[iatom] == st["s0"]:
                    st["t1"] = frame
                if s[iatom] != SITELABEL["interstitial"]:
                    if st["s0"] != SITELABEL["interstitial"]:
                        edge = (st["s0"], int(s[iatom]))
                        tau = st["t1"] - st["t0"]
                        tbarrier = frame - st["t1"]
                        if edge not in props:
                            edges.append(edge)
                            props[edge] = {"tau": [], "tbarrier": [], "frame": [], "iatom": []}
                        props[edge]["tau"].append(tau)
                        props[edge]["tbarrier"].append(tbarrier)
                        props[edge]["frame"].append(frame)
                        props[edge]["iatom"].append(iatom)
                    state[iatom] = {"s0": int(s[iatom]), "t0": frame, "t1": None}
        slast = s.copy()
    last_frame = F - 1
    for iatom in range(n_atoms):
        if state[iatom]["s0"] == SITELABEL["interstitial"]:
            continue
        node = state[iatom]["s0"]
        tau = last_frame - state[iatom]["t0"]
        if node not in props:
            nodes.append(node)
            props[node] = {"tau": [], "tbarrier": [], "frame": [], "iatom": []}
        props[node]["tau"].append(tau)
        props[node]["tbarrier"].append(None)
        props[node]["frame"].append(last_frame)
        props[node]["iatom"].append(iatom)
    for k, d in props.items():
        d["frame"] = np.array(d["frame"])
        d["iatom"] = np.array(d["iatom"])
        d["tau"] = dt * np.array(d["tau"])
        try:
            d["tbarrier"] = dt * np.array(d["tbarrier"])
        except TypeError:
            d["tbarrier"] = np.array(d["tbarrier"])
    return props, edges, nodes


def analyze_hops_vectorised(hop_x):
    """Second restatement of A.4, vectorised over atoms.  Returns integer event arrays
    (frame, iatom, s0, s, tau, tbarrier) sorted frame-major then atom, and node records
    (iatom, s0, tau) sorted by atom; times in frames."""
    hop_x = np.asarray(hop_x)
    F, N = hop_x.shape
    s0 = hop_x[0].astype(np.int64)
    slast = s0.copy()
    t0 = np.zeros(N, np.int64)
    t1 = np.full(N, -1, np.int64)
    ev = []
    for t in range(1, F):
        s = hop_x[t].astype(np.int64)
        s = np.where(s == -1, slast, s)
        differs = s != s0
        leave = differs & (slast == s0)
        t1 = np.where(leave, t, t1)
        enter = differs & (s != 0)
        emit = enter & (s0 != 0)
        ia = np.nonzero(emit)[0]
        if len(ia):
            ev.append(np.stack([np.full(len(ia), t), ia, s0[ia], s[ia], t1[ia] - t0[ia], t - t1[ia]], axis=1))
        s0 = np.where(enter, s, s0)
        t0 = np.where(enter, t, t0)
        t1 = np.where(enter, -1, t1)
        slast = s
    events = np.concatenate(ev) if ev else np.zeros((0, 6), np.int64)
    ia = np.nonzero(s0 != 0)[0]
    nodes = np.stack([ia, s0[ia], (F - 1) - t0[ia]], axis=1) if len(ia) else np.zeros((0, 3), np.int64)
    return events, nodes


def events_to_counts(events):
    """Transition-count matrix in COO form: sorted unique (s0, s) pairs and len(tau) per pair."""
    if len(events) == 0:
        return np.zeros((0, 2), np.int64), np.zeros(0, np.int64)
    pairs, counts = np.unique(events[:, 2:4], axis=0, return_counts=True)
    return pairs, counts


def graph_alltransitions(hop_x):
    """TransportNetwork.graph_alltransitions (hop/graph.py:107-135): the set of (slast -> s)
    pairs over consecutive frames, interstitial and outliers included.  Sorted (K,2) array."""
    hop_x = np.asarray(hop_x).astype(np.int64)
    if hop_x.shape[0] < 2:
        return np.zeros((0, 2), np.int64)
    a = hop_x[:-1].ravel()
    b = hop_x[1:].ravel()
    return np.unique(np.stack([a, b], axis=1), axis=0)


# =================================================================================================
# Threshold scans (row f.4 of the scope table)
# =================================================================================================
def density_stats(sites, props, bulk_label=None):
    """Density.stats (hop/sitemap.py:1301-1339) on the sites without interstitial and bulk."""
    nodes = [l for l in range(len(sites)) if l != SITELABEL["interstitial"] and l != bulk_label]
    vol = props["volume"][nodes]
    occ = props["occupancy_avg"][nodes]
    return dict(N_sites=len(nodes), site_volume_avg=np.average(vol), site_volume_std=np.std(vol),
                site_occupancy_rho_avg=np.average(occ), site_occupancy_rho_std=np.std(occ))


def scan_thresholds(grid, edges, bulk_sites, cutoffs, density_unit="Angstrom^{-3}", minsite=1):
    """DensityScanner.scan for one density (hop/analysis.py:165-195): per cut-off map_sites,
    site_insert_bulk, stats."""
    out = {}
    for rhocut in cutoffs:
        sites = label_sites_ndimage(grid, rhocut, minsite)
        sites = site_insert_bulk(sites, bulk_sites)
        props = site_properties(sites, grid, edges, density_unit)
        out[rhocut] = density_stats(sites, props, bulk_label=SITELABEL["bulk"])
    return out


# =================================================================================================
# Rates from waiting times (row f.2 of the scope table)
# =================================================================================================
def survival_faithful(taus, t):
    """S(t) = <Theta(t0 - t)> over the waiting times t0, the way hop.graph.survivalfunction
    (hop/graph.py:2433-2499) builds it: one step function per waiting time, Theta(0) = 1/2
    (Unitstep, hop/graph.py:2501-2518), summed and divided by N."""
    taus = np.asarray(taus, dtype=float)
    t = np.asarray(t, dtype=float)
    total = np.zeros(t.shape)
    for t0 in taus:
        total += 0.5 * (1 + np.sign(t0 - t))
    return total / len(taus)


def rate_faithful(taus):
    """HoppingGraph._rate (hop/graph.py:1677-1748): least-squares fit (scipy.optimize.leastsq, as in
    fit_func, hop/graph.py:2354-2397) of a*exp(-k1 t) + (1-a)*exp(-k2 t) to S(t) on t = 0, 1, ...,
    max(tau)+1 when there are more than 10 events; exp(-k t) when there are not, when the double
    exponential is asymmetric (|0.5 - a| > 0.49) or has a negative rate.  Returns (k, N, parameters)."""
    import scipy.optimize

    taus = np.asarray(taus)
    x = np.arange(0, np.max(taus) + 1.0 + 1, 1.0)
    y = survival_faithful(taus, x)

    def exp2(p, x):
        return p[0] * np.exp(-p[1] * x) + (1 - p[0]) * np.exp(-p[2] * x)

    def exp1(p, x):
        return np.exp(-p[0] * x)

    def lsq(f, p0):
        p, _ = scipy.optimize.leastsq(lambda p, x, y: f(p, x) - y, p0, args=(x, y))
        return np.atleast_1d(p)

    single = len(taus) <= 10
    if not single:
        p = lsq(exp2, [0.5, 0.1, 1e-4])
        a, k1, k2 = p
        single = abs(0.5 - a) > 0.49 or k1 < 0 or k2 < 0
        if not single:
            k = max(kk for kk in (k1, k2) if kk > 0)
    if single:
        p = lsq(exp1, [1.0])
        k = p[0]
    return k, len(taus), p


# =================================================================================================
# Whole pipeline (used by tests, smoke and the CPU baseline)
# =================================================================================================
def pipeline(xyz, delta=1.0, padding=2.0, solvent_threshold=math.e, bulk_threshold=math.exp(-0.5),
             density_unit="water", minsite=1, faithful=False, with_bulk=True):
    """density -> map_sites (+ bulk site from the same histogram at bulk_threshold, the
    no-solute case of DensityCreator(mode='all'), hop/density.py:257-272, 306-356) -> hoptraj ->
    transitions.  Returns a dict of every intermediate product."""
    xyz = np.asarray(xyz, dtype=np.float32)
    F = xyz.shape[0]
    bins, arange, edges = grid_geometry(xyz[0], delta, padding)
    if faithful:
        counts, _ = histogram_faithful(iter(xyz), bins, arange)
    else:
        counts = histogram_vectorised(xyz, edges)
    grid = make_density(counts, edges, F, density_unit)
    label = label_sites_faithful if faithful else label_sites_ndimage
    sites = label(grid, solvent_threshold, minsite)
    if with_bulk:
        bulk_sites = label(grid, bulk_threshold, minsite)
        if len(bulk_sites) > 1:
            sites = site_insert_bulk(sites, bulk_sites)
        else:
            sites = list(sites)
            sites.insert(SITELABEL["bulk"], [])
    site_map = draw_map_from_sites(sites, grid.shape)
    if faithful:
        hop = coord2hop_faithful(iter(xyz), edges, site_map)
        props, e, n = analyze_hops_faithful(hop[:, :, 0])
        events, nodes = None, None
    else:
        hop = coord2hop_vectorised(xyz, edges, site_map)
        events, nodes = analyze_hops_vectorised(hop[:, :, 0])
        props = None
    return dict(edges=edges, counts=np.asarray(counts), grid=grid, sites=sites, map=site_map,
                hop=hop, events=events, nodes=nodes, props=props)


# -------------------------------------------------------------------------------------------------
# Per-site observables (hop/siteanalysis.py) -- SURVEY.md 8f rank 3
# -------------------------------------------------------------------------------------------------
def site_histogram_edges(lo_bin, hi_bin, n_bins):
    """SiteHistogramArray.__init__ (hop/siteanalysis.py:503-507): bins = linspace(lo, hi, Nbins, endpoint=False), one
    more edge appended at the same spacing."""
    bins = np.linspace(lo_bin, hi_bin, n_bins, endpoint=False)
    return np.concatenate((bins, [2 * bins[-1] - bins[-2]]))


def site_distance_histogram(xyz, labels, centers, site2indx, edges, start=0, stop=None, step=1):
    """Distance._compute_distance + SiteHistogramArray.add, frame by frame as the reference does it
    (hop/siteanalysis.py:753-776, 531-544).  xyz [F][N][3] float32, labels [F][N] (hoptraj x or y, integral),
    centers [L][3] float64 with NaN for the interstitial, site2indx [L] -> row.  The label indexes the arrays
    directly (so -1 reads the last entries), numpy.digitize sends NaN to the top, and the fancy-index `+= 1`
    counts a (row, bin) pair once per frame.  Returns _histogram [max(site2indx) + 1][len(edges) + 1]."""
    F = xyz.shape[0]
    stop = F if stop is None else stop
    centers = np.asarray(centers, dtype=np.float64)
    site2indx = np.asarray(site2indx)
    hist = np.zeros((int(site2indx.max()) + 1, len(edges) + 1), dtype=np.int64)
    for f in range(start, stop, step):
        sites = np.asarray(labels[f]).astype(int)
        distvec = centers[sites] - xyz[f]
        dist = np.sqrt(np.sum(distvec * distvec, axis=1))
        bins = [np.digitize([x], edges)[0] for x in dist]
        hist[site2indx[sites], bins] += 1
    return hist


def site_occupancy(labels, n_labels):
    """Atoms per site and frame: numpy.histogram of the frame's labels with unit bins
    (Occupancy._compute_occupancy, hop/siteanalysis.py:821-841); outliers (-1) fall outside."""
    labels = np.asarray(labels).astype(np.int64)
    occ = np.zeros((labels.shape[0], n_labels), dtype=np.int64)
    for f in range(labels.shape[0]):
        l = labels[f]
        occ[f] = np.bincount(l[l >= 0], minlength=n_labels)[:n_labels]
    return occ
