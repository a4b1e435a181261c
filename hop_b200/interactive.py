"""Workflow glue with the reference's call order and defaults (hop/interactive.py:235-307,
417-464).  Thin by design: it only strings the four stage classes together."""
from __future__ import annotations

import numpy

from .density import DensityCreator
from .graph import TransportNetwork
from .trajectory import HoppingTrajectory
from .universe import as_Universe


def generate_densities(*args, **kwargs):
    """Solvent and bulk density from (topology, trajectory) or a Universe; returns the dict of
    densities with the site map of the solvent density computed and the bulk site inserted.

    Defaults as in hop/interactive.py:285-289: solvent_threshold = e, bulk_threshold = exp(-0.5),
    density_unit 'water', files water.pickle / bulk.pickle."""
    density_unit = kwargs.pop('density_unit', 'water')
    solvent_threshold = kwargs.pop('solvent_threshold', numpy.e)
    bulk_threshold = kwargs.pop('bulk_threshold', numpy.exp(-0.5))
    solvent_density = kwargs.pop('solvent_density', 'water')
    bulk_density = kwargs.pop('bulk_density', 'bulk')
    save = kwargs.pop('save', True)
    DC = DensityCreator(*args, **kwargs)
    densities = DC.create()
    if save:
        densities['solvent'].save(solvent_density)
        if 'bulk' in densities:
            densities['bulk'].save(bulk_density)
    if 'bulk' in densities and 'solvent' in densities:
        DC.DensityWithBulk(density_unit=density_unit, solvent_threshold=solvent_threshold,
                           bulk_threshold=bulk_threshold)
    else:
        d = densities['solvent']
        d.convert_density(density_unit)
        d.map_sites(solvent_threshold)
        d.site_insert_nobulk()
    if save:
        densities['solvent'].save(solvent_density)
    return densities


def make_hoppingtraj(density, filename, universe=None, atomselection=None, **hopargs):
    """Hopping trajectory for `density` (hop/interactive.py:417-447); topology/trajectory and the
    selection default to what the density's metadata recorded."""
    if universe is None:
        universe = as_Universe(density.metadata['psf'], density.metadata['dcd'])
    if atomselection is None:
        atomselection = density.metadata['atomselection']
    group = universe.select_atoms(atomselection)
    hops = HoppingTrajectory(universe.trajectory, group, density, **hopargs)
    hops.write(filename)
    return hops


def build_hoppinggraph(hoppingtrajectory, density):
    """Compute the hopping graph from a hopping trajectory and the hydration sites
    (hop/interactive.py:449-464)."""
    tgraph = TransportNetwork(hoppingtrajectory, density)
    return tgraph.HoppingGraph()
