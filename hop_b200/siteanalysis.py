"""Per-site observables on the GPU: drop-in for the part of hop.siteanalysis that accumulates histograms of an
observable per hydration site over a trajectory (hop/siteanalysis.py:866-1107 ``Collection``, :722-802
``Distance`` / ``Orbit``, :804-861 ``Occupancy`` / ``OrbitOccupancy``, :440-599 ``SiteHistogramArray``).

    C = Collection(selection=universe.select_atoms('name OH2'),
                   hoptraj=HoppingTrajectory(filename='hoptraj'),
                   density=Density(filename='water.pickle'))
    C.add_observable('distance', lo_bin=0, hi_bin=3)
    C.compute()
    C.SiteHistograms['distance'].site_histogram([2, 3, 5])

The frame loop of the reference (one Python iteration per frame and observable, a list comprehension of
numpy.digitize calls per atom) is one kernel launch per observable here (csrc/siteanalysis.cu).

Parity.  ``distance`` and ``orbit`` are pinned by the reference itself: tests/golden/reference/siteanalysis/siteanalysis.npz
was produced by hop's own Collection / Distance / Orbit (tests/golden/make_reference_golden.py) and the
histograms here equal it count for count, including three behaviours one would not guess -- a (site, bin)
pair hit by several atoms of one frame counts once (fancy-index ``+=``, :544), outliers (label -1) index the
arrays from the end and so count as the highest site, and the interstitial's NaN distance lands in the upper
outlier column.  ``occupancy`` / ``orbitoccupancy`` CANNOT run in the reference: ``Occupancy._occupancy_histogram``
drops the last bin of a histogram whose bins already end one short of ``histogrammed_sites``, so
``SiteHistogramArray.add`` gets index arrays of lengths N and N - 1 and raises
(hop/siteanalysis.py:821-841; shown by tests/test_reference_golden.py::test_reference_occupancy_cannot_run).
They are implemented as documented -- "number of particles on the site", one value per site and frame,
histogrammed with the same SiteHistogramArray rules -- and their parity is unpinned.
Site trajectories (``trajectory=True``) exist for the occupancies; the reference raises NotImplementedError
for the distances' (:745-749) and so does this.  Plotting is out of scope.
"""
from __future__ import annotations

import numpy

from .constants import SITELABEL
from .exceptions import MissingDataError
from .utilities import Saveable


class SiteHistogramArray(Saveable):
    """One histogram per collected site: ``_histogram`` int64 [Nsites + 1][Nbins + 2] -- the extra row takes every
    site that is not collected, the first / last column values below / above the range (hop/siteanalysis.py:440-599)."""

    _saved_attributes = 'all'

    def __init__(self, name, sites, maxsites, parameters, properties=None):
        self.name = name
        self.sites = numpy.asarray(sites, dtype=numpy.int64)
        self.maxsites = int(maxsites)
        self.parameters = dict(parameters)
        self.properties = dict(properties or {})
        self.Nsites = len(self.sites)
        # label -> row; labels that are not collected share the extra row Nsites
        self.site2indx = numpy.full(self.maxsites + 1, self.Nsites, dtype=numpy.int64)
        self.site2indx[self.sites] = numpy.arange(self.Nsites)
        self.indx2site = self.sites
        try:
            Nbins, lo_bin, hi_bin = parameters['Nbins'], parameters['lo_bin'], parameters['hi_bin']
        except KeyError:
            raise KeyError('parameters must be a dict containing Nbins, lo_bin, hi_bin')
        self.bins = numpy.linspace(lo_bin, hi_bin, Nbins, endpoint=False)
        self.edges = numpy.concatenate((self.bins, [2 * self.bins[-1] - self.bins[-2]]))
        self.midpoints = 0.5 * (self.edges[1:] + self.edges[:-1])
        self.delta = self.edges[1:] - self.edges[:-1]
        self._histogram = numpy.zeros((self.Nsites + 1, Nbins + 2), dtype=numpy.int64)
        self.distribution = numpy.zeros((self.Nsites, Nbins))

    def reset(self):
        self._histogram[:] = 0
        self.distribution[:] = 0

    def normalize(self):
        """Normalisation (outlier columns included) and distribution (hop/siteanalysis.py:546-550)."""
        self.normalization = numpy.sum(self._histogram[:-1], axis=1)
        self.normalization[self.normalization == 0] = -1
        self.distribution = self.histogram / (self.normalization[:, numpy.newaxis] * self.delta)

    @property
    def histogram(self):
        return self._histogram[:-1, 1:-1]

    def site_histogram(self, sites):
        return self.histogram[self.site2indx[sites]]

    def site_distribution(self, sites):
        return self.distribution[self.site2indx[sites]]

    data = histogram
    site_data = site_histogram

    def mean(self):
        return numpy.sum(self.distribution * self.bins, axis=1)

    def std(self):
        EX = self.mean()
        EXX = numpy.sum(self.distribution * self.bins ** 2, axis=1)
        return numpy.sqrt(EXX - EX * EX)

    def mean_std(self):
        return self.sites, self.mean(), self.std()


class SiteTrajectoryArray(object):
    """Value of an observable per collected site and frame: ``trajectory`` [Nsites][n_frames]."""

    def __init__(self, name, sites, maxsites, n_frames):
        self.name = name
        self.sites = numpy.asarray(sites, dtype=numpy.int64)
        self.site2indx = numpy.full(int(maxsites) + 1, len(self.sites), dtype=numpy.int64)
        self.site2indx[self.sites] = numpy.arange(len(self.sites))
        self.trajectory = numpy.zeros((len(self.sites), int(n_frames)))

    def site_trajectory(self, sites):
        return self.trajectory[self.site2indx[sites]]


_OBSERVABLES = {
    # name: (hoptraj plane, kind, default bin width)
    'distance': (0, 'distance', 0.1),
    'orbit': (1, 'distance', 0.2),
    'occupancy': (0, 'occupancy', 1),
    'orbitoccupancy': (1, 'occupancy', 1),
}
ObservablesRegistry = dict((name, name) for name in _OBSERVABLES)


class Collection(Saveable):
    """Site-resolved histograms of observables over a trajectory (hop/siteanalysis.py:866-1107)."""

    _saved_attributes = ['siteselection', 'site_properties', '_all_sites', 'SiteHistograms', 'SiteTrajectories']

    def __init__(self, selection=None, hoptraj=None, density=None, include_sites='default',
                 exclude_sites=['default', 'bulk'], filename=None, verbosity=None):
        if not (selection is None and hoptraj is None and density is None):
            if selection is None or hoptraj is None or density is None:
                raise ValueError("Not enough data in argument list. Either supply selection, "
                                 "hoptraj, and density or load from a pickled file <filename>.pickle")
            self.selection = selection
            self.universe = selection.universe
            self.hoptraj = hoptraj
            self.density = density
            self.siteselection = numpy.asarray(density.site_labels(include=include_sites, exclude=exclude_sites))
            self.site_properties = density.site_properties
            self._all_sites = numpy.asarray(density.site_labels('default'))
            self.SiteHistograms = {}
            self.SiteTrajectories = {}
            self._observables = {}
            if filename is not None:
                self.filename(filename, set_default=True)
            n_data = len(self.universe.trajectory)
            if n_data != hoptraj.n_frames:
                raise ValueError("Data trajectory and hopping trajectory differ in length: "
                                 "%d vs %d frames" % (n_data, hoptraj.n_frames))
            # hoptraj column of every selected atom: one hoptraj atom per (segment, residue)
            hop_group = hoptraj.group if getattr(hoptraj, 'group', None) is not None else hoptraj.tgroup
            by_residue = {}
            for ihop, atom in enumerate(hop_group):
                key = (atom.segid, atom.resid)
                if key in by_residue:
                    raise ValueError("Each residue should only have a single atom in the HoppingTrajectory. "
                                     "Detected a collision: atoms %d and %d" % (by_residue[key], ihop))
                by_residue[key] = ihop
            try:
                self.idcd2ihop = numpy.array([by_residue[(atom.segid, atom.resid)] for atom in selection], dtype=numpy.int32)
            except KeyError as err:
                raise KeyError("An atom of the selection belongs to a residue %r that has no atom in the hop "
                               "trajectory. Change the selection appropriately." % (err.args[0],))
        elif filename is not None:
            self.load(filename)
        else:
            raise ValueError("Not enough data in argument list. Either supply selection, "
                             "hoptraj, and density or load from a pickled file <filename>.pickle")

    def add_observable(self, name, lo_bin, hi_bin, Nbins=None, trajectory=False, **kwargs):
        """Histogram observable `name` for the collected sites when compute() runs (hop/siteanalysis.py:1001-1030)."""
        if name not in _OBSERVABLES:
            raise ValueError('Observable ' + str(name) + ' not implemented. Choose one of ' + str(list(_OBSERVABLES)))
        if hi_bin <= lo_bin:
            raise ValueError("histogram parameters must be hi_bin > lo_bin, but lo_bin=%g and hi_bin=%g." % (lo_bin, hi_bin))
        plane, kind, width = _OBSERVABLES[name]
        if not Nbins:
            Nbins = int((hi_bin - lo_bin) / width)
        if trajectory and kind == 'distance':
            raise NotImplementedError('Current implementation of distance site traj is broken.')   # as hop, :745-749
        parameters = {'Nbins': Nbins, 'lo_bin': lo_bin, 'hi_bin': hi_bin}
        H = SiteHistogramArray(name, self.siteselection, self._all_sites.max(), parameters)
        self._observables[name] = dict(plane=plane, kind=kind, histograms=H, trajectory=bool(trajectory))

    def del_observable(self, names=None):
        for name in self._names(names):
            del self._observables[name]

    def _names(self, names):
        if names is None:
            return list(self._observables)
        if isinstance(names, str):
            names = [names]
        missing = [n for n in names if n not in self._observables]
        if missing:
            raise ValueError("Only observables defined in %r can be used as <observables>." % list(self._observables))
        return list(names)

    # -- data on the device -------------------------------------------------------------------------------
    def _device_inputs(self):
        import torch

        if getattr(self, '_dev', None) is None:
            x, y = self.hoptraj.hop_arrays()
            planes = [torch.from_numpy(numpy.ascontiguousarray(p, dtype=numpy.float32)).cuda() for p in (x, y)]
            traj = self.universe.trajectory
            idx = numpy.asarray(self.selection.indices)
            if hasattr(traj, 'coordinate_array'):
                block = traj.coordinate_array(0, None, 1)
                if not (len(idx) == block.shape[1] and numpy.array_equal(idx, numpy.arange(len(idx)))):
                    block = numpy.take(block, idx, axis=1)
            else:
                block = numpy.stack([numpy.array(self.selection.positions, dtype=numpy.float32) for ts in traj])
            xyz = torch.from_numpy(numpy.ascontiguousarray(block, dtype=numpy.float32)).cuda()
            self._dev = (planes, xyz)
        return self._dev

    def compute(self, observables=None, start=None, stop=None, step=1, progress_interval=10):
        """Fill SiteHistograms (and SiteTrajectories) for frames start, start + step, ... < stop."""
        import torch

        from . import kernels as K

        try:
            n_frames = self.hoptraj.n_frames
            centers = self.density.site_properties.center
        except AttributeError:
            raise MissingDataError("Cannot compute() histograms because trajectory data is missing.")
        if start is None:
            start = 0
        elif start < 0:
            start = n_frames + start
        if stop is None:
            stop = n_frames
        elif stop < 0:
            stop = n_frames + stop + 1
        names = self._names(observables)
        planes, xyz = self._device_inputs()
        L = len(self.density.site_properties)
        for name in names:
            o = self._observables[name]
            H = o['histograms']
            H.reset()
            labels = planes[o['plane']]
            if o['kind'] == 'distance':
                c = numpy.array([numpy.full(3, numpy.nan) if (v is None or numpy.ndim(v) == 0) else numpy.asarray(v, float)
                                 for v in centers], dtype=numpy.float64)
                c[SITELABEL['interstitial']] = numpy.nan
                # row of every label; the arrays of the reference are indexed by the label itself, so labels beyond
                # maxsites cannot occur for a consistent density
                s2i = numpy.full(L, H.Nsites, dtype=numpy.int32)
                n = min(L, len(H.site2indx))
                s2i[:n] = H.site2indx[:n]
                hist = torch.zeros(H._histogram.shape, dtype=torch.int64, device='cuda')
                identity = numpy.array_equal(self.idcd2ihop, numpy.arange(len(self.idcd2ihop)))
                amap = None if identity else torch.from_numpy(self.idcd2ihop).cuda()
                K.site_distance_hist(xyz, labels, torch.from_numpy(c).cuda(), torch.from_numpy(s2i).cuda(),
                                     torch.from_numpy(numpy.ascontiguousarray(H.edges)).cuda(), H.Nsites + 1, hist,
                                     start=start, stop=stop, step=step, atom_map=amap)
                H._histogram[:] = hist.cpu().numpy()
            else:
                occ = K.site_occupancy(labels, L)[start:stop:step].cpu().numpy()        # [frames][labels]
                self._add_occupancy(o, occ, n_frames, start, stop, step)
            H.normalize()
        self.SiteHistograms.update((name, self._observables[name]['histograms']) for name in names)

    def _add_occupancy(self, o, occ, n_frames, start, stop, step):
        """One value per site and frame into the site histograms: SiteHistogramArray.add(nu, sites) per frame for the
        labels hmin .. hmax of the proper sites (hop/siteanalysis.py:821-841 as documented; see the module docstring)."""
        H = o['histograms']
        hmin, hmax = int(self._all_sites.min()), int(self._all_sites.max())
        labels = numpy.arange(hmin, hmax + 1)
        nu = occ[:, hmin:hmax + 1]
        bins = numpy.digitize(nu.ravel(), H.edges).reshape(nu.shape)        # 0 .. Nbins + 1
        rows = H.site2indx[labels]
        collected = rows < H.Nsites
        width = H._histogram.shape[1]
        # collected sites: one value per frame each
        for r, col in zip(rows[collected], numpy.nonzero(collected)[0]):
            H._histogram[r] += numpy.bincount(bins[:, col], minlength=width)
        # all other sites share one row, where the fancy-index += counts a bin once per frame
        if (~collected).any():
            hit = numpy.zeros((bins.shape[0], width), dtype=bool)
            fr = numpy.repeat(numpy.arange(bins.shape[0]), int((~collected).sum()))
            hit[fr, bins[:, ~collected].ravel()] = True
            H._histogram[H.Nsites] += hit.sum(axis=0)
        if o['trajectory']:
            T = SiteTrajectoryArray(H.name, H.sites, self._all_sites.max(), n_frames)
            T.trajectory[:, start:stop:step] = occ[:, H.sites].T
            self.SiteTrajectories[H.name] = T
