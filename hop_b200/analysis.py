"""Density statistics as a function of the threshold -- drop-in for the scanning part of
hop.analysis (DensityAnalysis.scan / DensityScanner, hop/analysis.py:128-226).

For every cut-off the reference redraws the site map of each density (``map_sites``), puts the bulk
site back (``site_insert_bulk`` -- redrawing the map erases it, hop/analysis.py:168-171) and collects
``Density.stats()``.  Each of those ``map_sites`` calls is the Python cell loop + networkx of stage 2
(minutes per call at 10^6 site cells); here it is the GPU labelling (csrc/label.cu, < 1 ms for 8 M
cells), so a scan over dozens of thresholds costs what one map cost before.

Equivalence sites (``find_equivalence_sites_with``, hop/sitemap.py:1082-1279) are outside this
package; the scan therefore reports every density (and the optional reference) on its own.
"""
from __future__ import annotations

import logging

import numpy

from . import sitemap, utilities

logger = logging.getLogger("MDAnalysis.app.hop.analysis")


class DensityScanner(utilities.Saveable):
    """Scan density statistics over a range of thresholds.

        S = DensityScanner(densities, bulkdensities, reference=None, refbulkdensity=None)
        S.scan([1.5, 2.0, 2.72, 3.0])
        S.scanstats[2.72][0]['N_sites'];  S.scanarrays[0].N_sites

    ``densities`` / ``bulkdensities``: Density objects (or lists of them, or file names) on the same
    grid pairwise; the bulk densities must carry their own site map (thresholded at the bulk
    threshold) because site 1 of it is re-inserted after every remapping."""

    _saved_attributes = ['scanstats']

    def __init__(self, densities=None, bulkdensities=None, reference=None, refbulkdensity=None, with_densities=True):
        self.scanstats = {}
        self.with_densities = with_densities and densities is not None
        if self.with_densities:
            self.densities = [utilities.easy_load(d, sitemap.Density, 'stats') for d in utilities.asiterable(densities)]
            self.bulkdensities = [utilities.easy_load(d, sitemap.Density, 'stats')
                                  for d in utilities.asiterable(bulkdensities)] if bulkdensities is not None else None
            self.reference = utilities.easy_load(reference, sitemap.Density, 'stats') if reference is not None else None
            self.refbulkdensity = (utilities.easy_load(refbulkdensity, sitemap.Density, 'stats')
                                   if refbulkdensity is not None else None)

    def scan(self, cutoffs, merge=True):
        """Recompute the site maps and statistics for each threshold in `cutoffs`
        (hop/analysis.py:165-195).  merge=False erases earlier results."""
        if not self.with_densities or not self.bulkdensities:
            # redrawing the map at a new cut-off erases the manually inserted bulk site
            raise ValueError("Densities (including bulk) are required for a scan.")
        if len(self.bulkdensities) != len(self.densities):
            raise ValueError("One bulk density per density is required.")
        if (self.reference is None) != (self.refbulkdensity is None):
            raise ValueError("A reference density needs its bulk density (and vice versa).")
        for d, b in zip(self.densities, self.bulkdensities):
            if d.map is not None and b.map is not None and tuple(d.map.shape) != tuple(b.map.shape):
                raise ValueError("density and bulk density must be defined on the same grid")
        if not merge:
            self.scanstats = {}
        try:
            iter(cutoffs)
        except TypeError:
            cutoffs = [cutoffs]
        for rhocut in cutoffs:
            stats = {}
            if self.reference is not None:
                self.reference.map_sites(rhocut)
                self.reference.site_insert_bulk(self.refbulkdensity)
                stats['reference'] = self.reference.stats()
            for k, (d, b) in enumerate(zip(self.densities, self.bulkdensities)):
                d.map_sites(rhocut)
                d.site_insert_bulk(b)
                stats[k] = d.stats()
            self.scanstats[rhocut] = stats
        self._scan2recarray()
        logger.info("Scan complete. See scanstats and scanarrays attributes.")

    def load(self, fn=None, merge=True):
        super(DensityScanner, self).load(fn, merge=merge)
        self._scan2recarray()

    def _scan2recarray(self):
        """scanstats -> one record array per density, one row per threshold (hop/analysis.py:211-224)."""
        stats = self.scanstats
        self.scanarrays = {}
        if not stats:
            return
        xkeys = sorted(stats.keys())
        ids = sorted(stats[xkeys[0]].keys(), key=str)
        fieldnames = sorted(stats[xkeys[0]][ids[0]].keys())
        for id_ in ids:
            cols = [[stats[x][id_][y] for x in xkeys] for y in fieldnames]
            self.scanarrays[id_] = numpy.rec.fromarrays(cols, names=fieldnames)
