"""Bulk-solvent selection '<atomsel> not within <cutoff> of <solutesel>' (hop/density.py:82-87).

The reference delegates to MDAnalysis' `notwithin_coordinates_factory` (a kd-tree neighbour search
per frame).  Here the test runs on the GPU, fused with the histogram (csrc/hist.cu,
`hist_bulk_kernel`): per frame the solute atoms go into a shared-memory cell list and every solvent
atom checks the 27 cells around it.  "Within" is plain Euclidean distance <= cutoff, no periodic
images, evaluated in float32.
"""
from __future__ import annotations

import numpy


def _torch():
    import torch

    return torch


class NotWithinCoordinates:
    """Callable returning the coordinates of the solvent atoms of the current frame that are not
    within `cutoff` of any solute atom (what `notwithin_coordinates_factory` returns)."""

    def __init__(self, universe, sel1, sel2, cutoff, not_within=True):
        self.universe = universe
        self.solvent = universe.select_atoms(sel1)
        self.solute = universe.select_atoms(sel2)
        self.cutoff = float(cutoff)
        self.not_within = bool(not_within)
        self._grid = None

    def _dummy_grid(self):
        from . import kernels as K

        if self._grid is None:
            big = numpy.array([-1.0e30, 1.0e30])
            self._grid = K.GridTables([big, big, big])
            self._counts = _torch().zeros(1, dtype=_torch().int32, device="cuda")
        return self._grid

    def within_mask(self, solvent_xyz, solute_xyz):
        """uint8 numpy [n][Nw]: 1 where a solute atom is within the cutoff (one or more frames)."""
        from . import kernels as K

        t = _torch()
        w = t.from_numpy(numpy.ascontiguousarray(solvent_xyz, dtype=numpy.float32)).cuda()
        p = t.from_numpy(numpy.ascontiguousarray(solute_xyz, dtype=numpy.float32)).cuda()
        if w.dim() == 2:
            w, p = w[None], p[None]
        within = t.empty(w.shape[:2], dtype=t.uint8, device="cuda")
        K.hist_bulk(self._dummy_grid(), w, p, self.cutoff, self._counts, within=within)
        return within.cpu().numpy()

    def __call__(self, cutoff=None):
        if cutoff is not None and float(cutoff) != self.cutoff:
            self.cutoff = float(cutoff)
        w = numpy.asarray(self.solvent.positions)
        if len(w) == 0:
            return w
        mask = self.within_mask(w, numpy.asarray(self.solute.positions))[0].astype(bool)
        return w[~mask] if self.not_within else w[mask]


def notwithin_coordinates_factory(universe, sel1, sel2, cutoff, not_within=True, use_kdtree=True):
    """Generate the optimized selection '*sel1* not within *cutoff* of *sel2*' (MDAnalysis'
    notwithin_coordinates_factory, called at hop/density.py:84-85).  `use_kdtree` is accepted for
    signature compatibility; the search always runs on the GPU."""
    return NotWithinCoordinates(universe, sel1, sel2, cutoff, not_within=not_within)
