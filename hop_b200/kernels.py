"""Thin functional layer over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function here enqueues hand-written sm_100a kernels from libhopb200.so on torch's current
stream.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import SUMMARY_INTS, STATE_INTS, HOPB_MAX_SITES

EVENT_DTYPE = np.dtype([("frame", "<i4"), ("iatom", "<i4"), ("s0", "<i4"), ("s", "<i4"),
                        ("tau", "<i4"), ("tbarrier", "<i4")])


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _h(t=None):
    return _lib.handle(t.device.index if t is not None else None)


def _rows(t, dtype, name):
    """A 2-d CUDA tensor whose rows are contiguous (a column range of a wider contiguous array is fine)."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.dim() == 2 and (t.shape[1] <= 1 or t.stride(1) == 1)):
        raise TypeError("%s must be a CUDA tensor of dtype %s with contiguous rows" % (name, dtype))
    return t


def _ld(t):
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def _cuda(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA tensor of dtype %s" % (name, dtype))
    return t


class GridTables:
    """Device-side binning tables for one set of bin edges (hopb_grid)."""

    def __init__(self, edges, device=None):
        self.edges = [np.ascontiguousarray(e, dtype=np.float64) for e in edges]
        if len(self.edges) != 3:
            raise ValueError("need three edge arrays")
        self.shape = tuple(len(e) - 1 for e in self.edges)
        self.n_cells = int(np.prod(self.shape))
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._handle = _lib.handle(self.device.index)
        # `decimal` of hop/trajectory.py:435, computed with numpy exactly like the reference
        decs = np.array([int(-np.log10(np.diff(e).min())) + 6 for e in self.edges], dtype=np.int32)
        out = ctypes.c_void_p()
        e = self.edges
        rc = _lib.lib().hopb_grid_create(self._handle, e[0].ctypes.data, self.shape[0], e[1].ctypes.data, self.shape[1],
                                         e[2].ctypes.data, self.shape[2], decs.ctypes.data, ctypes.byref(out))
        _lib.check(self._handle, rc)
        self._grid = out
        # number of distinct (dx, dy, dz) bin-width triples; 0: irregular edges, no count thresholds
        self.width_classes = int(_lib.lib().hopb_grid_width_classes(out))

    @property
    def ptr(self):
        return self._grid

    def set_periodic_box(self, origin=None, length=None):
        """Wrap every coordinate into [origin, origin + length) before it is binned (stages 1 and 3).  Off by
        default -- hop never wraps (hop/density.py:21-28) -- and switched off again with no arguments."""
        if origin is None or length is None:
            rc = _lib.lib().hopb_grid_set_periodic_box(self._grid, None, None)
            self.periodic_box = None
        else:
            o = np.ascontiguousarray(origin, dtype=np.float64).reshape(3)
            l = np.ascontiguousarray(length, dtype=np.float64).reshape(3)
            rc = _lib.lib().hopb_grid_set_periodic_box(self._grid, o.ctypes.data, l.ctypes.data)
            self.periodic_box = (o, l)
        _lib.check(self._handle, rc)

    def __del__(self):
        try:
            if getattr(self, "_grid", None):
                _lib.lib().hopb_grid_destroy(self._grid)
                self._grid = None
        except Exception:
            pass


def hist_accumulate(grid: GridTables, xyz: torch.Tensor, counts: torch.Tensor, cells=None):
    """counts[cell] += points per cell (hop/density.py:125-131).  xyz: float32 [..., 3].
    `cells` (int32, one per point) optionally receives the packed hopping-trajectory cell."""
    _cuda(xyz, torch.float32, "xyz")
    _cuda(counts, torch.int32, "counts")
    if counts.numel() != grid.n_cells:
        raise ValueError("counts has the wrong size")
    n_points = xyz.numel() // 3
    if cells is not None:
        _cuda(cells, torch.int32, "cells")
        if cells.numel() != n_points:
            raise ValueError("cells has the wrong size")
    if n_points == 0:
        return counts       # a frame with no selected atoms contributes nothing (hop/density.py:128)
    h = _h(xyz)
    _lib.check(h, _lib.lib().hopb_hist_accumulate_cells(h, _stream(), grid.ptr, _p(xyz), n_points, _p(counts), _p(cells)))
    return counts


def hist_accumulate_multi(grids, species: torch.Tensor, xyz: torch.Tensor, counts, cells=None):
    """One pass over xyz (float32 [F][N][3]) for several selections: atom i belongs to selection species[i] (uint8) and is
    binned on grids[species[i]] into counts[species[i]]; `cells` (int32 [F][N]) receives its brick index in that grid
    (hopb_hist_accumulate_multi; hop reads the trajectory once per selection, scripts/hop-generate-many-densities.py:98-125)."""
    _cuda(xyz, torch.float32, "xyz")
    _cuda(species, torch.uint8, "species")
    F, N = int(xyz.shape[0]), int(xyz.shape[1])
    if species.numel() != N or len(grids) != len(counts):
        raise ValueError("one species id per atom, one count array per grid")
    for g, c in zip(grids, counts):
        _cuda(c, torch.int32, "counts")
        if c.numel() != g.n_cells:
            raise ValueError("counts has the wrong size")
    if cells is not None:
        _cuda(cells, torch.int32, "cells")
        if tuple(cells.shape) != (F, N):
            raise ValueError("cells has the wrong shape")
    K = len(grids)
    gp = (ctypes.c_void_p * K)(*[g.ptr for g in grids])
    cp = (ctypes.c_void_p * K)(*[c.data_ptr() for c in counts])
    h = _h(xyz)
    _lib.check(h, _lib.lib().hopb_hist_accumulate_multi(h, _stream(), K, gp, _p(species), _p(xyz), F, N, cp, _p(cells)))
    return counts


def hist_bulk(grid: GridTables, solvent: torch.Tensor, solute: torch.Tensor, cutoff: float, counts_bulk: torch.Tensor,
              counts_solvent=None, cells=None, within=None):
    """Solvent histogram + bulk histogram ('solvent not within cutoff of solute', hop/density.py:82-87)
    from one pass.  solvent float32 [F][Nw][3], solute float32 [F][Np][3]."""
    _cuda(solvent, torch.float32, "solvent")
    _cuda(counts_bulk, torch.int32, "counts_bulk")
    F, Nw = int(solvent.shape[0]), int(solvent.shape[1])
    Np = 0
    if solute is not None and solute.numel():
        _cuda(solute, torch.float32, "solute")
        if solute.shape[0] != F:
            raise ValueError("solvent and solute must cover the same frames")
        Np = int(solute.shape[1])
    if F == 0 or Nw == 0:
        return counts_bulk
    h = _h(solvent)
    rc = _lib.lib().hopb_hist_bulk(h, _stream(), grid.ptr, _p(solvent), _p(solute if Np else None), F, Nw, Np, float(cutoff),
                                   _p(counts_solvent), _p(counts_bulk), _p(cells), _p(within))
    _lib.check(h, rc)
    return counts_bulk


def set_hist_l2_budget(nbytes: int, device=None):
    """Largest part of the count grid one histogram pass reduces into (slabbed passes beyond it)."""
    h = _lib.handle(device)
    _lib.check(h, _lib.lib().hopb_set_hist_l2_budget(h, int(nbytes)))


def brick_shape(shape):
    """(n_bricks, words of the 2-bit block map) for a grid shape."""
    nb = 1
    for n in shape:
        nb *= -(-int(n) // 4)
    return nb, -(-nb // 16)


def fine_map_bytes(shape):
    """Size of the fine block map buffer of brick_map: two 1-bit maps + the bulk-cell counter."""
    nb = brick_shape(shape)[0]
    return 2 * ((nb + 3) // 4 * 4) + 16


class BrickMap(object):
    """The site map as stage 3 reads it: `brick` (int16, 4x4x4 brick order), `coarse` (2 bits per
    brick), `fine` (1 bit per 2x2x2 sub-brick, one map for "all cells interstitial" and one for "all
    cells bulk"), `cell2` (2 bits per cell, for grids whose block maps do not fit into shared memory).
    Unpacks as (brick, coarse)."""

    def __init__(self, brick, coarse, fine, cell2=None):
        self.brick, self.coarse, self.fine, self.cell2 = brick, coarse, fine, cell2

    def __iter__(self):
        return iter((self.brick, self.coarse))


def brick_map(map16: torch.Tensor, brick=None, coarse=None, fine=None, cell2=None):
    """Site map in 4x4x4 brick order (what the hoptraj kernel gathers from) and its block maps, which
    let the kernel skip the gather where a whole (sub-)brick carries one label: `coarse`, 2 bits per
    brick (uniformly interstitial / bulk / mixed), and `fine`, 1 bit per 2x2x2 sub-brick (uniformly
    interstitial, uniformly bulk: two maps, the kernel picks the one most atoms fall under).  `cell2`
    (2 bits per cell: interstitial / bulk / other) is the small gather target for grids too large for
    either."""
    _cuda(map16, torch.int16, "map16")
    nx, ny, nz = map16.shape
    nb, words = brick_shape(map16.shape)
    if brick is None:
        brick = torch.empty(nb * 64, dtype=torch.int16, device=map16.device)
    if coarse is None:
        coarse = torch.empty(words, dtype=torch.int32, device=map16.device)
    if fine is None:
        fine = torch.zeros(fine_map_bytes(map16.shape), dtype=torch.uint8, device=map16.device)
    if cell2 is None:
        cell2 = torch.empty(nb * 4, dtype=torch.int32, device=map16.device)
    h = _h(map16)
    _lib.check(h, _lib.lib().hopb_brick_map(h, _stream(), nx, ny, nz, _p(map16), _p(brick), _p(coarse), _p(fine),
                                            _p(cell2)))
    return BrickMap(brick, coarse, fine, cell2)


def density_from_counts(grid: GridTables, counts: torch.Tensor, n_frames: int, factor: float = 1.0, out=None,
                        ctas_per_sm: int = 0):
    """counts -> density, bit-identical to DensityCollector.finish + Grid.make_density + convert_density.
    ctas_per_sm > 0 limits the grid (for running beside other kernels on another stream)."""
    _cuda(counts, torch.int32, "counts")
    if out is None:
        out = torch.empty(grid.shape, dtype=torch.float64, device=counts.device)
    h = _h(counts)
    _lib.check(h, _lib.lib().hopb_density_from_counts_ex(h, _stream(), grid.ptr, _p(counts), int(n_frames), float(factor),
                                                         _p(out), int(ctas_per_sm)))
    return out


def label_sites(density: torch.Tensor, threshold: float, minsite: int = 1):
    """(labels int32 [nx,ny,nz], n_sites, volumes int32[n_sites+1]) -- Density.map_sites."""
    _cuda(density, torch.float64, "density")
    if density.dim() != 3:
        raise ValueError("density must be 3-dimensional")
    nx, ny, nz = density.shape
    labels = torch.empty(density.shape, dtype=torch.int32, device=density.device)
    volumes = torch.zeros(HOPB_MAX_SITES + 1, dtype=torch.int32, device=density.device)
    n_sites = ctypes.c_int32(0)
    h = _h(density)
    rc = _lib.lib().hopb_label_sites(h, _stream(), nx, ny, nz, _p(density), float(threshold), int(minsite), _p(labels),
                                     ctypes.byref(n_sites), _p(volumes))
    _lib.check(h, rc)
    return labels, int(n_sites.value), volumes[: n_sites.value + 1]


def label_sites_async(density: torch.Tensor, threshold: float, minsite: int, labels: torch.Tensor,
                      n_sites_dev: torch.Tensor, volumes=None, lane: int = 0):
    """Density.map_sites without a host round trip: `n_sites_dev` (CUDA int32[1]) receives the
    number of sites; a value above HOPB_MAX_SITES means the int16 map cannot hold them."""
    _cuda(density, torch.float64, "density")
    _cuda(labels, torch.int32, "labels")
    _cuda(n_sites_dev, torch.int32, "n_sites_dev")
    nx, ny, nz = density.shape
    h = _lib.handle(density.device.index, lane)   # lane > 0: own scratch, safe on a concurrent stream
    rc = _lib.lib().hopb_label_sites_async(h, _stream(), nx, ny, nz, _p(density), float(threshold), int(minsite),
                                           _p(labels), _p(n_sites_dev), _p(volumes))
    _lib.check(h, rc)
    return labels


def insert_bulk(own: torch.Tensor, bulk_labels, bulklabel: int = 1, member=None, map16=None):
    """In place on `own`; returns (member uint8, map16 int16)."""
    _cuda(own, torch.int32, "own")
    if bulk_labels is not None:
        _cuda(bulk_labels, torch.int32, "bulk_labels")
        if bulk_labels.shape != own.shape:
            raise ValueError("The bulk density was defined on a different grid than this density.")
    if member is None:
        member = torch.empty(own.shape, dtype=torch.uint8, device=own.device)
    if map16 is None:
        map16 = torch.empty(own.shape, dtype=torch.int16, device=own.device)
    h = _h(own)
    _lib.check(h, _lib.lib().hopb_insert_bulk(h, _stream(), own.numel(), _p(own), _p(bulk_labels), int(bulklabel), _p(member), _p(map16)))
    return member, map16


def labels_to_map16(labels: torch.Tensor, map16=None):
    _cuda(labels, torch.int32, "labels")
    if map16 is None:
        map16 = torch.empty(labels.shape, dtype=torch.int16, device=labels.device)
    h = _h(labels)
    _lib.check(h, _lib.lib().hopb_labels_to_map16(h, _stream(), labels.numel(), _p(labels), _p(map16)))
    return map16


def label_sites_counts_async(grid: GridTables, counts: torch.Tensor, n_frames: int, factor: float, threshold: float,
                             minsite: int, out: torch.Tensor, n_sites_dev: torch.Tensor, volumes=None, lane: int = 0):
    """label_sites_async on density_from_counts(counts, n_frames, factor) without that density: the
    threshold is translated into a count threshold per bin-width class on the device (identical labels).
    Needs grid.width_classes > 0."""
    _cuda(counts, torch.int32, "counts")
    if tuple(counts.shape) != grid.shape:
        raise ValueError("counts must have the shape of the grid")
    h = _lib.handle(counts.device.index, lane)
    rc = _lib.lib().hopb_label_sites_counts_async(h, _stream(), grid.ptr, _p(counts), int(n_frames), float(factor),
                                                  float(threshold), int(minsite), _p(out), _p(n_sites_dev), _p(volumes))
    _lib.check(h, rc)
    return out


def site_map_fused_ok(grid: GridTables):
    """Whether site_map_fused handles this grid (regular bin widths, z extent within its shared-memory tile)."""
    return 0 < grid.width_classes <= 512 and grid.shape[2] <= 3072


def site_map_fused(grid: GridTables, counts: torch.Tensor, n_frames: int, factor: float, solvent_threshold: float,
                   bulk_threshold: float, with_bulk: bool, minsite: int, density, own, member, map16, brick: "BrickMap",
                   n_sites_dev: torch.Tensor, props=None, volumes=None):
    """Stage 2 in one launch (hopb_site_map_fused): counts -> density, hydration-site labels, bulk site,
    int16 map, bricked map + block maps and (props = (count, mean, std, center) buffers) the annotation."""
    _cuda(counts, torch.int32, "counts")
    if tuple(counts.shape) != grid.shape:
        raise ValueError("counts must have the shape of the grid")
    _cuda(density, torch.float64, "density")
    _cuda(own, torch.int32, "own")
    _cuda(map16, torch.int16, "map16")
    _cuda(n_sites_dev, torch.int32, "n_sites_dev")
    if with_bulk:
        _cuda(member, torch.uint8, "member")
    n_cap, pc, pm, ps, pcen = 0, None, None, None, None
    if props is not None:
        pc, pm, ps, pcen = props
        n_cap = int(pc.numel())
    h = _h(counts)
    rc = _lib.lib().hopb_site_map_fused(h, _stream(), grid.ptr, _p(counts), int(n_frames), float(factor),
                                        float(solvent_threshold), float(bulk_threshold), 1 if with_bulk else 0, int(minsite),
                                        _p(density), _p(own), _p(member if with_bulk else None), _p(map16), _p(brick.brick),
                                        _p(brick.coarse), _p(brick.fine), _p(brick.cell2), _p(n_sites_dev), _p(volumes),
                                        n_cap, _p(pc), _p(pm), _p(ps), _p(pcen))
    _lib.check(h, rc)
    return density, own, member, map16


def site_properties(grid: GridTables, density: torch.Tensor, own: torch.Tensor, member, n_labels: int, out=None):
    """(count int64[n], mean f64[n], std f64[n], center f64[n,3]) per site label.  `n_labels` may be
    an upper bound (rows past the real number of sites come back empty)."""
    _cuda(density, torch.float64, "density")
    _cuda(own, torch.int32, "own")
    dev = density.device
    if out is not None:
        count, mean, std, center = out
    else:
        count = torch.empty(n_labels, dtype=torch.int64, device=dev)
        mean = torch.empty(n_labels, dtype=torch.float64, device=dev)
        std = torch.empty(n_labels, dtype=torch.float64, device=dev)
        center = torch.empty((n_labels, 3), dtype=torch.float64, device=dev)
    h = _h(density)
    _lib.check(h, _lib.lib().hopb_site_properties(h, _stream(), grid.ptr, _p(density), _p(own), _p(member), int(n_labels),
                                                  _p(count), _p(mean), _p(std), _p(center)))
    return count, mean, std, center


def rate_fit(taus_list, device=None, perturb=0.0, noise=0):
    """HoppingGraph._rate for many edges at once (hopb_rate_fit): taus_list = the waiting times of every edge.
    Returns (out float64 [E][8] = kind, p0, p1, p2, k, info, nfev, N; S float64 concatenated; off_x int64 [E + 1])."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    E = len(taus_list)
    lens = np.array([len(t) for t in taus_list], dtype=np.int64)
    if E == 0 or lens.min() == 0:
        if E == 0:
            return np.zeros((0, 8)), np.zeros(0), np.zeros(1, np.int64)
        raise ValueError("every edge needs at least one waiting time")
    off_tau = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tau = np.concatenate([np.asarray(t, dtype=np.float64) for t in taus_list])
    tmax = np.maximum.reduceat(tau, off_tau[:-1])
    m = np.array([len(np.arange(0, t + 2.0, 1.0)) for t in tmax], dtype=np.int64)      # x = arange(0, tmax + 1, dt), tmax = max + dt
    off_x = np.concatenate([[0], np.cumsum(m)]).astype(np.int64)
    total = int(off_x[-1])
    d_tau = torch.from_numpy(tau).to(dev)
    d_ot = torch.from_numpy(off_tau).to(dev)
    d_ox = torch.from_numpy(off_x).to(dev)
    S = torch.empty(total, dtype=torch.float64, device=dev)
    out = torch.empty((E, 8), dtype=torch.float64, device=dev)
    h = _lib.handle(dev.index)
    _lib.check(h, _lib.lib().hopb_rate_fit(h, _stream(), _p(d_tau), _p(d_ot), _p(d_ox), E, total, float(perturb), int(noise), _p(S), _p(out)))
    return out.cpu().numpy(), S.cpu().numpy(), off_x


def site_occupancy(labels: torch.Tensor, n_labels: int, out=None):
    """occ int32 [F][n_labels]: atoms per site and frame (hop/siteanalysis.py:837-841); labels = a hoptraj plane."""
    _cuda(labels, torch.float32, "labels")
    F, N = labels.shape
    if out is None:
        out = torch.empty((F, n_labels), dtype=torch.int32, device=labels.device)
    h = _h(labels)
    _lib.check(h, _lib.lib().hopb_site_occupancy(h, _stream(), _p(labels), int(F), int(N), int(n_labels), _p(out)))
    return out


def site_distance_hist(xyz: torch.Tensor, labels: torch.Tensor, centers: torch.Tensor, site2indx: torch.Tensor, edges: torch.Tensor,
                       n_rows: int, hist: torch.Tensor, start=0, stop=None, step=1, atom_map=None):
    """hist[row][bin] += per-frame (site, bin) hits of the atom-to-site-centre distance (hop/siteanalysis.py:753-776,
    531-544); see hopb_site_distance_hist for the reference behaviours it keeps."""
    _cuda(xyz, torch.float32, "xyz")
    _cuda(labels, torch.float32, "labels")
    _cuda(centers, torch.float64, "centers")
    _cuda(site2indx, torch.int32, "site2indx")
    _cuda(edges, torch.float64, "edges")
    _cuda(hist, torch.int64, "hist")
    F, N = int(xyz.shape[0]), int(xyz.shape[1])
    if labels.shape[0] != F:
        raise ValueError("Data trajectory and hopping trajectory differ in length: %d vs %d frames" % (F, labels.shape[0]))
    if atom_map is not None:
        _cuda(atom_map, torch.int32, "atom_map")
    stop = F if stop is None else int(stop)
    n_bins = int(edges.numel()) - 1
    if tuple(hist.shape) != (int(n_rows), n_bins + 2):
        raise ValueError("hist must have shape (n_rows, n_bins + 2)")
    h = _h(xyz)
    rc = _lib.lib().hopb_site_distance_hist(h, _stream(), _p(xyz), _p(labels), _p(atom_map), N, int(labels.shape[1]), int(start),
                                            stop, int(step), _p(centers), _p(site2indx), int(centers.shape[0]), _p(edges),
                                            n_bins, int(n_rows), _p(hist))
    _lib.check(h, rc)
    return hist


def site_cells(own: torch.Tensor, member, n_labels: int):
    """CSR cell lists: (cells int32[M], offsets int64[n_labels+1])."""
    _cuda(own, torch.int32, "own")
    dev = own.device
    offsets = torch.empty(n_labels + 1, dtype=torch.int64, device=dev)
    n_cells = ctypes.c_int64(0)
    h = _h(own)
    capacity = 0
    cells = None
    for _ in range(2):
        rc = _lib.lib().hopb_site_cells(h, _stream(), own.numel(), _p(own), _p(member), int(n_labels), _p(cells), capacity,
                                        _p(offsets), ctypes.byref(n_cells))
        if rc == _lib.HOPB_E_CAPACITY:
            capacity = int(n_cells.value)
            cells = torch.empty(max(capacity, 1), dtype=torch.int32, device=dev)
            continue
        _lib.check(h, rc)
        break
    if cells is None:
        cells = torch.empty(0, dtype=torch.int32, device=dev)
    return cells[: n_cells.value], offsets


def plan_chunks(n_frames: int, n_atoms: int, num_sms: int = 148, min_chunk: int = 32, warps_per_sm: int = 32,
                max_waves: int = 4):
    """(n_chunks, chunk_len) for the fused hoptraj kernel: one thread per (atom, chunk), dealt out as
    work items of (32 consecutive atoms, chunk) to the warps of the grid.

    The kernel is latency-bound per thread, so the items should fill a whole number of waves of
    resident warps (num_sms * warps_per_sm; the kernel holds 1024 threads per SM): pick the chunk
    count with the best wave efficiency, preferring fewer chunks (less stitching) when efficiencies
    are within 2 %."""
    items_per_chunk = -(-n_atoms // 32)
    capacity = num_sms * warps_per_sm
    max_chunks = max(1, min(n_frames // min_chunk, 65535, (max_waves * capacity) // max(items_per_chunk, 1)))
    best, best_eff = 1, 0.0
    for c in range(1, max_chunks + 1):
        items = c * items_per_chunk
        eff = items / (-(-items // capacity) * capacity)
        if eff > best_eff + 0.02:
            best, best_eff = c, eff
    chunk_len = -(-n_frames // best)
    n_chunks = -(-n_frames // chunk_len)
    return int(n_chunks), int(chunk_len)


class EventBuffer:
    """Device buffer of hopb_event records plus the running device-side count."""

    def __init__(self, capacity: int, device):
        self.capacity = int(capacity)
        self.data = torch.empty((max(self.capacity, 1), 6), dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)

    def reset(self):
        self.count.zero_()

    def n(self):
        return int(self.count.item())


def hoptraj(grid: GridTables, xyz, brick, frame0: int, n_chunks: int, chunk_len: int,
            events: EventBuffer, hop_x=None, hop_y=None, summaries=None, cells=None, coarse=None, fine=None,
            fine_value=-1, block_map="auto", cell2=None):
    """Fused _coord2hop + transition scan over one slab (hop/trajectory.py:403-468, hop/graph.py:212-254).
    Input is either `xyz` (float32 [F][N][3]) or the brick indices `cells` (int32 [F][N]) from
    hist_accumulate.  `brick` is the BrickMap of brick_map(Density.map) -- block_map = "auto" (the
    fine block map when it fits into shared memory, else the coarse one, else the 2-bit cell map),
    "fine", "coarse", "cell2" or "none"
    -- or the bricked int16 tensor with `coarse` / `fine` given separately.  fine_value picks the fine
    map (0: interstitial, 1: bulk, -1: decided on the device from the size of the bulk site)."""
    if isinstance(brick, BrickMap):
        bm = brick
        brick = bm.brick
        coarse = bm.coarse if block_map in ("auto", "coarse") else None
        fine = bm.fine if block_map in ("auto", "fine") else None
        cell2 = bm.cell2 if block_map in ("auto", "cell2") else None
    _cuda(brick, torch.int16, "brick")
    map16 = brick
    # Inputs and outputs may be column ranges of wider arrays (several selections binned in one pass,
    # hist_accumulate_multi): rows `ld` atoms apart, atoms adjacent.
    if cells is not None:
        _rows(cells, torch.int32, "cells")
        F, N = cells.shape[0], cells.shape[1]
        dev = cells.device
        ld = _ld(cells)
    else:
        if not (isinstance(xyz, torch.Tensor) and xyz.is_cuda and xyz.dtype == torch.float32 and xyz.dim() == 3
                and xyz.stride(2) == 1 and xyz.stride(1) == 3 and xyz.stride(0) % 3 == 0):
            raise TypeError("xyz must be a CUDA float32 [F][N][3] tensor (or a column range of one)")
        F, N = xyz.shape[0], xyz.shape[1]
        dev = xyz.device
        ld = xyz.stride(0) // 3 if F > 1 else N
    # planes: float32 (the DCD record type) or int16 (the same labels, half the bytes; both planes then)
    planes16 = hop_x is not None and hop_x.dtype == torch.int16
    if planes16 and (hop_y is None or hop_y.dtype != torch.int16):
        raise TypeError("int16 planes: hop_x and hop_y must both be int16")
    for name, t in (("hop_x", hop_x), ("hop_y", hop_y)):
        if t is not None:
            _rows(t, torch.int16 if planes16 else torch.float32, name)
            if tuple(t.shape) != (F, N) or (F > 1 and _ld(t) != ld):
                raise ValueError("%s must have the shape and row stride of the input" % name)
    nb = brick_shape(grid.shape)[0]
    if map16.numel() != nb * 64:
        raise ValueError("bricked map does not match the grid (use kernels.brick_map)")
    if coarse is not None:
        _cuda(coarse, torch.int32, "coarse")
    if fine is not None:
        _cuda(fine, torch.uint8, "fine")
        if fine.numel() != fine_map_bytes(grid.shape):
            raise ValueError("fine block map does not match the grid (use kernels.brick_map)")
    if cell2 is not None:
        _cuda(cell2, torch.int32, "cell2")
        if cell2.numel() != nb * 4:
            raise ValueError("2-bit cell map does not match the grid (use kernels.brick_map)")
    if summaries is None:
        summaries = torch.empty((n_chunks, SUMMARY_INTS, N), dtype=torch.int32, device=dev)
    h = _h(map16)
    fn = _lib.lib().hopb_hoptraj_i16 if planes16 else _lib.lib().hopb_hoptraj_strided
    rc = fn(h, _stream(), grid.ptr, _p(None if cells is not None else xyz), _p(cells), F, N, int(ld),
            int(frame0), _p(map16), _p(coarse), _p(fine), int(fine_value), _p(cell2), _p(hop_x),
            _p(hop_y), int(n_chunks), int(chunk_len), _p(summaries), _p(events.data),
            _p(events.count), events.capacity)
    _lib.check(h, rc)
    return summaries


def hopscan_labels(hop_x: torch.Tensor, frame0: int, n_chunks: int, chunk_len: int, events: EventBuffer, summaries=None):
    _cuda(hop_x, torch.float32, "hop_x")
    F, N = hop_x.shape
    if summaries is None:
        summaries = torch.empty((n_chunks, SUMMARY_INTS, N), dtype=torch.int32, device=hop_x.device)
    h = _h(hop_x)
    rc = _lib.lib().hopb_hopscan_labels(h, _stream(), _p(hop_x), F, N, int(frame0), int(n_chunks), int(chunk_len),
                                        _p(summaries), _p(events.data), _p(events.count), events.capacity)
    _lib.check(h, rc)
    return summaries


def hop_combine(summaries: torch.Tensor, chunk_len: int, n_frames: int):
    n_chunks, _, N = summaries.shape
    slab = torch.empty((SUMMARY_INTS, N), dtype=torch.int32, device=summaries.device)
    h = _h(summaries)
    _lib.check(h, _lib.lib().hopb_hop_combine(h, _stream(), N, n_chunks, int(chunk_len), int(n_frames), _p(summaries), _p(slab)))
    return slab


def hop_advance(summaries, state: torch.Tensor, init: bool):
    N = state.shape[1]
    n = 0 if summaries is None else summaries.shape[0]
    h = _h(state)
    _lib.check(h, _lib.lib().hopb_hop_advance(h, _stream(), N, n, _p(summaries), 1 if init else 0, _p(state)))
    return state


def hop_fixup(summaries: torch.Tensor, chunk_len: int, n_frames: int, state: torch.Tensor, hop_y, events: EventBuffer):
    n_chunks, _, N = summaries.shape
    h = _h(summaries)
    y16 = hop_y is not None and hop_y.dtype == torch.int16
    ld = N if hop_y is None else _ld(_rows(hop_y, torch.int16 if y16 else torch.float32, "hop_y"))
    fn = _lib.lib().hopb_hop_fixup_i16 if y16 else _lib.lib().hopb_hop_fixup_strided
    rc = fn(h, _stream(), N, ld, n_chunks, int(chunk_len), int(n_frames), _p(summaries), _p(state),
            _p(hop_y), _p(events.data), _p(events.count), events.capacity)
    _lib.check(h, rc)
    return state


def new_state(n_atoms: int, device):
    return torch.zeros((STATE_INTS, n_atoms), dtype=torch.int32, device=device)


def events_finalize(events: EventBuffer, n_events: int, n_atoms: int, n_frames_total: int, n_labels: int,
                    count_matrix=None, out=None):
    """Sort events into reference order; returns (sorted [n,6] int32, edge_s0, edge_s, edge_start, edge_count).
    `count_matrix` (zeroed CUDA int32 [(n_labels+1)^2]) optionally receives the dense transition counts;
    `out` may supply the five output buffers (sized for n events) to avoid allocations."""
    dev = events.data.device
    n = int(n_events)
    cap_edges = max(1, min(n, (n_labels + 1) ** 2))
    if out is not None:
        sorted_ev, e_s0, e_s, e_start, e_count = out
    else:
        sorted_ev = torch.empty((max(n, 1), 6), dtype=torch.int32, device=dev)
        e_s0 = torch.empty(cap_edges, dtype=torch.int32, device=dev)
        e_s = torch.empty(cap_edges, dtype=torch.int32, device=dev)
        e_start = torch.empty(cap_edges, dtype=torch.int64, device=dev)
        e_count = torch.empty(cap_edges, dtype=torch.int64, device=dev)
    n_edges = ctypes.c_int64(0)
    h = _h(events.data)
    rc = _lib.lib().hopb_events_finalize_ex(h, _stream(), _p(events.data), n, int(n_atoms), int(n_frames_total),
                                            int(n_labels), _p(sorted_ev), _p(e_s0), _p(e_s), _p(e_start), _p(e_count),
                                            _p(count_matrix), ctypes.byref(n_edges))
    _lib.check(h, rc)
    k = int(n_edges.value)
    return sorted_ev[:n], e_s0[:k], e_s[:k], e_start[:k], e_count[:k]


FIN_OVERFLOW, FIN_TOO_MANY_SITES, FIN_MATRIX_SKIPPED = 1, 2, 4


def events_finalize_dev(events: EventBuffer, n_atoms: int, n_frames_total: int, n_sites_dev: torch.Tensor,
                        label_offset: int, out, count_matrix, result_dev: torch.Tensor):
    """events_finalize without any host round trip: the event count (events.count) and the number of
    sites (n_sites_dev, int32) are read on the device.  `out` = five capacity-sized output buffers,
    `count_matrix` = square int32 buffer (or None) filled with its own row stride, `result_dev` = int64
    [4] receiving (n_events, n_edges, n_labels, status bits FIN_*)."""
    sorted_ev, e_s0, e_s, e_start, e_count = out
    stride = 0 if count_matrix is None else int(count_matrix.shape[0])
    h = _h(events.data)
    rc = _lib.lib().hopb_events_finalize_dev(h, _stream(), _p(events.data), _p(events.count), events.capacity,
                                             int(n_atoms), int(n_frames_total), _p(n_sites_dev), int(label_offset),
                                             _p(sorted_ev), _p(e_s0), _p(e_s), _p(e_start), _p(e_count),
                                             _p(count_matrix), stride, _p(result_dev))
    _lib.check(h, rc)


def nvls_allreduce_u32(t: torch.Tensor, mc_ptr: int, rank: int, world: int):
    """In-place sum over the ranks of a 32-bit integer tensor in symmetric memory (csrc/exchange.cu); mc_ptr is the
    multicast address of `t`.  The caller brackets it with cross-rank barriers (hop_b200/exchange.py)."""
    if not (t.is_cuda and t.is_contiguous() and t.element_size() == 4):
        raise TypeError("a contiguous CUDA tensor of 32-bit integers is needed")
    h = _h(t)
    _lib.check(h, _lib.lib().hopb_nvls_allreduce_u32(h, _stream(), ctypes.c_void_p(int(mc_ptr)), t.numel(), int(rank), int(world)))
    return t


def p2p_allreduce_u32(t: torch.Tensor, peer_ptrs, mc_ptr: int, rank: int, world: int):
    """nvls_allreduce_u32 with the additions on the SM: peer_ptrs = the ranks' copies of `t` as mapped into this process."""
    if not (t.is_cuda and t.is_contiguous() and t.element_size() == 4):
        raise TypeError("a contiguous CUDA tensor of 32-bit integers is needed")
    arr = (ctypes.c_void_p * world)(*[int(p) for p in peer_ptrs])
    h = _h(t)
    _lib.check(h, _lib.lib().hopb_p2p_allreduce_u32(h, _stream(), arr, ctypes.c_void_p(int(mc_ptr)), t.numel(), int(rank), int(world)))
    return t


def nvls_allgather(src: torch.Tensor, out: torch.Tensor, mc_out_ptr: int, rank: int, world: int):
    """src -> slot `rank` of every rank's copy of out ([world] x src) through the switch (csrc/exchange.cu)."""
    n_bytes = src.numel() * src.element_size()
    if not (src.is_cuda and src.is_contiguous() and out.is_contiguous() and out.numel() * out.element_size() == world * n_bytes):
        raise TypeError("out must be [world] x src, both contiguous CUDA tensors")
    h = _h(src)
    _lib.check(h, _lib.lib().hopb_nvls_allgather(h, _stream(), _p(src), ctypes.c_void_p(int(mc_out_ptr)), n_bytes, int(rank), int(world)))
    return out


def labels_to_int16(labels: torch.Tensor, out=None):
    """float32 hoptraj plane -> int16 (lossless), for the download (hopb_labels_to_int16)."""
    _cuda(labels, torch.float32, "labels")
    if out is None:
        out = torch.empty(labels.shape, dtype=torch.int16, device=labels.device)
    h = _h(labels)
    _lib.check(h, _lib.lib().hopb_labels_to_int16(h, _stream(), _p(labels), labels.numel(), _p(out)))
    return out


def all_transitions(hop_x: torch.Tensor, n_labels: int):
    """Sorted (K,2) int64 array of (label[t-1], label[t]) pairs (hop/graph.py:107-135)."""
    x16 = hop_x.dtype == torch.int16
    _cuda(hop_x, torch.int16 if x16 else torch.float32, "hop_x")
    F, N = hop_x.shape
    W = n_labels + 1
    words = (W * W + 31) // 32
    bitmap = torch.zeros(words, dtype=torch.int32, device=hop_x.device)
    h = _h(hop_x)
    fn = _lib.lib().hopb_all_transitions_i16 if x16 else _lib.lib().hopb_all_transitions
    _lib.check(h, fn(h, _stream(), _p(hop_x), F, N, int(n_labels), _p(bitmap)))
    bits = np.unpackbits(bitmap.cpu().numpy().view(np.uint8), bitorder="little")[: W * W]
    idx = np.nonzero(bits)[0]
    return np.stack([idx // W - 1, idx % W - 1], axis=1).astype(np.int64)


class SynthStream(object):
    """The synthetic trajectory chunk by chunk, each chunk continuing where the previous one stopped (a source for
    HopPipeline.run_streamed).  chunk(f0, f1) must be called for consecutive ranges; a call with f0 == 0 restarts."""

    def __init__(self, params, device=None):
        self.params = params
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        centres, half = params.lattice()
        self._c = _lib.SynthParamsC(seed=params.seed, n_atoms=params.n_atoms, n_frames=params.n_frames,
                                    site_stride=params.site_stride, n_sites=len(centres), box=np.float32(params.box),
                                    cube_half=np.float32(half), sigma_site=np.float32(params.sigma_site),
                                    sigma_bulk=np.float32(params.sigma_bulk), p_hop=np.float32(params.p_hop),
                                    p_rebind=np.float32(params.p_rebind))
        self._centres = torch.from_numpy(np.ascontiguousarray(centres)).to(self.device)
        self._walker = torch.zeros((params.n_atoms, 4), dtype=torch.float32, device=self.device)
        self._next = 0
        self._bufs = {}

    def chunk(self, f0, f1):
        if f0 != 0 and f0 != self._next:
            raise ValueError("SynthStream: chunks must be consecutive (expected frame %d, got %d)" % (self._next, f0))
        n = int(f1 - f0)
        out = self._bufs.get(n)
        if out is None:
            out = self._bufs[n] = torch.empty((n, self.params.n_atoms, 3), dtype=torch.float32, device=self.device)
        h = _lib.handle(self.device.index)
        rc = _lib.lib().hopb_synth_generate_chunk(h, _stream(), ctypes.byref(self._c), _p(self._centres), int(f0), n, _p(out),
                                                  _p(self._walker), 1 if f0 > 0 else 0)
        _lib.check(h, rc)
        self._next = int(f1)
        return out

    __call__ = chunk


def synth_generate(params, frame_start=0, n_out=None, device=None, out=None):
    """Device twin of hop_b200.synth.generate: float32 [n_out][N][3] in HBM."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    centres, half = params.lattice()
    if n_out is None:
        n_out = params.n_frames - frame_start
    c = _lib.SynthParamsC(seed=params.seed, n_atoms=params.n_atoms, n_frames=params.n_frames,
                          site_stride=params.site_stride, n_sites=len(centres), box=np.float32(params.box),
                          cube_half=np.float32(half), sigma_site=np.float32(params.sigma_site),
                          sigma_bulk=np.float32(params.sigma_bulk), p_hop=np.float32(params.p_hop),
                          p_rebind=np.float32(params.p_rebind))
    d_centres = torch.from_numpy(np.ascontiguousarray(centres)).to(dev)
    if out is None:
        out = torch.empty((n_out, params.n_atoms, 3), dtype=torch.float32, device=dev)
    h = _lib.handle(dev.index)
    rc = _lib.lib().hopb_synth_generate(h, _stream(), ctypes.byref(c), _p(d_centres), int(frame_start), int(n_out), _p(out))
    _lib.check(h, rc)
    torch.cuda.current_stream().synchronize()  # d_centres must outlive the kernel
    return out


def events_to_numpy(ev: torch.Tensor):
    """[n,6] int32 device tensor -> structured numpy array (EVENT_DTYPE)."""
    return np.ascontiguousarray(ev.cpu().numpy()).view(EVENT_DTYPE).reshape(-1)
