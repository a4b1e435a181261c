"""Bulk-solvent selection '<atomsel> not within <cutoff> of <solutesel>' (hop/density.py:82-87).

The reference delegates to MDAnalysis' kd-tree (`notwithin_coordinates_factory`).  This is a
"next" row of the hot-path scope table and is not built yet.
"""


def notwithin_coordinates_factory(universe, sel1, sel2, cutoff, not_within=True, use_kdtree=True):
    raise NotImplementedError(
        "bulk selection (atomselection not within cutoff of soluteselection) is not implemented yet; "
        "use mode='solvent' or pass soluteselection=None")
