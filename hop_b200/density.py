"""Stage 1: densities from trajectories -- the drop-in for hop.density.

Same surface as hop/density.py: ``DensityCollector`` (59-178), ``DensityCreator`` (180-356),
``density_from_Universe`` (358-369), ``density_from_trajectory`` (372-430).  The per-frame
``numpy.histogramdd`` + ``grid += h`` of ``DensityCollector.collect`` is replaced by the CUDA
histogram kernel (csrc/hist.cu): coordinates are staged to HBM in large frame blocks through pinned
memory and binned there; the counts never leave the device until someone reads ``Density.grid``.
"""
from __future__ import annotations

import logging

import numpy

from .exceptions import MissingDataError
from .sitemap import Density
from .universe import as_Universe
from .utilities import fixedwidth_bins

logger = logging.getLogger("MDAnalysis.app.hop.density")

#: bytes of coordinates staged per host->device block
STAGE_BYTES = 256 << 20


def _torch():
    import torch

    return torch


class _FrameStager:
    """Collects float32 [n][N][3] coordinate blocks in pinned memory and ships them to the GPU on a
    copy stream, double buffered, so the histogram kernel of block k overlaps the copy of k+1."""

    def __init__(self, n_atoms, keep_on_device=False, n_frames_hint=None):
        t = _torch()
        if not t.cuda.is_available():
            raise RuntimeError("hop_b200 needs a CUDA device (no CPU fallback for the density histogram)")
        self.t = t
        self.n_atoms = int(n_atoms)
        per_frame = max(self.n_atoms * 12, 1)
        self.block_frames = max(1, min(STAGE_BYTES // per_frame, n_frames_hint or (1 << 30)))
        self.pinned = [t.empty((self.block_frames, self.n_atoms, 3), dtype=t.float32).pin_memory() for _ in range(2)]
        self.dev = [t.empty((self.block_frames, self.n_atoms, 3), dtype=t.float32, device="cuda") for _ in range(2)]
        self.copy_stream = t.cuda.Stream()
        self.done = [None, None]    # event: kernel finished reading dev[i]
        self.copied = [None, None]  # event: copy out of pinned[i] finished
        self.cur = 0
        self.fill = 0
        self.keep = [] if keep_on_device else None

    def host_block(self):
        """numpy view of the pinned block being filled (wait until its previous copy has been issued
        and consumed)."""
        return self.pinned[self.cur].numpy()

    def flush(self, n_frames, consume):
        """Ship the first n_frames of the current block and call consume(device_tensor)."""
        t = self.t
        i = self.cur
        main = t.cuda.current_stream()
        with t.cuda.stream(self.copy_stream):
            if self.done[i] is not None:
                self.copy_stream.wait_event(self.done[i])      # the kernel that read dev[i] is done
            if self.keep is not None:
                d = t.empty((n_frames, self.n_atoms, 3), dtype=t.float32, device="cuda")
            else:
                d = self.dev[i][:n_frames]
            d.copy_(self.pinned[i][:n_frames], non_blocking=True)
            self.copied[i] = t.cuda.Event()
            self.copied[i].record(self.copy_stream)
        main.wait_event(self.copied[i])
        consume(d)
        self.done[i] = t.cuda.Event()
        self.done[i].record(main)
        if self.keep is not None:
            d.record_stream(main)
            self.keep.append(d)
        self.cur ^= 1
        if self.copied[self.cur] is not None:
            self.copied[self.cur].synchronize()                 # pinned[cur] may be refilled now


class DensityCollector(object):
    """Collect subsequent coordinate frames to build up a :class:`Density` (hop/density.py:59-178)."""

    use_kdtree = True

    def __init__(self, name, universe, **kwargs):
        self.name = name
        try:
            universe.select_atoms('all')
            universe.trajectory.ts
        except AttributeError:
            errmsg = "DensityCollector: The universe must be a proper MDAnalysis.Universe instance."
            logger.fatal(errmsg)
            raise TypeError(errmsg)
        self.universe = u = universe
        self.delta = kwargs.pop('delta', 1.0)
        self.atomselection = kwargs.pop('atomselection', 'name OH2')
        self.cutoff = kwargs.pop('cutoff', 3.5)
        self.soluteselection = kwargs.pop('soluteselection', None)
        self.padding = kwargs.pop('padding', 2.0)
        self.metadata = kwargs.pop('metadata', {})
        self.parameters = kwargs.pop('parameters', {})
        if self.cutoff > 0 and self.soluteselection is not None:
            from .bulk import notwithin_coordinates_factory

            self.current_coordinates = notwithin_coordinates_factory(
                u, self.atomselection, self.soluteselection, self.cutoff, use_kdtree=self.use_kdtree)
            self.group = self.current_coordinates.solvent
            self.solute_group = self.current_coordinates.solute
            self.mode = "BULK"
        else:
            group = u.select_atoms(self.atomselection)
            self.group = group
            self.current_coordinates = lambda: group.positions
            self.mode = "SOLVENT"
        coord = self.current_coordinates()
        logger.info("%-10s: Selected %d atoms out of %d atoms (%s) from %d total.",
                    self.name, coord.shape[0], len(u.select_atoms(self.atomselection)),
                    self.atomselection, len(u.atoms))
        self._complete = False
        self._counts = None
        self._tables = None
        self._pending = None
        self._pending_n = 0

    def init_histogram(self, **kwargs):
        """Grid geometry from the padded extent of the current (first) frame; bin edges exactly as
        numpy.histogramdd lays them out (hop/density.py:99-117)."""
        from . import kernels as K

        t = _torch()
        smin = kwargs.pop("smin", None)
        smax = kwargs.pop("smax", None)
        if smin is None:
            smin = self.min_coordinates(padding=self.padding)
        if smax is None:
            smax = self.max_coordinates(padding=self.padding)
        BINS = fixedwidth_bins(self.delta, smin, smax)
        self.arange = list(zip(BINS['min'], BINS['max']))
        self.bins = BINS['Nbins']
        self.edges = [numpy.linspace(lo, hi, int(n) + 1) for (lo, hi), n in zip(self.arange, self.bins)]
        self._tables = K.GridTables(self.edges)
        self._counts = t.zeros(self._tables.shape, dtype=t.int32, device="cuda")
        self._n_collected = 0

    def min_coordinates(self, **kwargs):
        return numpy.min(self.current_coordinates(), axis=0) - kwargs.pop('padding', self.padding)

    def max_coordinates(self, **kwargs):
        return numpy.max(self.current_coordinates(), axis=0) + kwargs.pop('padding', self.padding)

    @property
    def grid(self):
        """Accumulated histogram (float64 copy of the device counts; divided by n_frames after
        finish(), like the reference attribute)."""
        if self._counts is None:
            raise AttributeError("grid")
        self._flush_pending()
        g = self._counts.cpu().numpy().astype(numpy.uint32).astype(numpy.float64)
        if self._complete:
            g /= float(self._n_frames)
        return g

    def collect(self):
        """Bin the current frame's selected coordinates (hop/density.py:125-131).  Frames are
        batched in pinned memory and histogrammed on the GPU a block at a time."""
        assert self._counts is not None, "init_histogram() must be called first"
        coord = self.current_coordinates()
        n = len(coord)
        if n > 0:
            self._push_points(numpy.asarray(coord, dtype=numpy.float32))
        return n

    def _push_points(self, pts):
        t = _torch()
        if self._pending is None:
            self._pending = t.empty((max(STAGE_BYTES // 12, len(pts)), 3), dtype=t.float32).pin_memory()
            self._pending_np = self._pending.numpy()
        if self._pending_n + len(pts) > self._pending_np.shape[0]:
            self._flush_pending()
        if len(pts) > self._pending_np.shape[0]:
            self._pending = t.empty((len(pts), 3), dtype=t.float32).pin_memory()
            self._pending_np = self._pending.numpy()
        self._pending_np[self._pending_n:self._pending_n + len(pts)] = pts
        self._pending_n += len(pts)

    def _flush_pending(self):
        from . import kernels as K

        if self._pending is not None and self._pending_n > 0:
            d = self._pending[:self._pending_n].cuda(non_blocking=True)
            K.hist_accumulate(self._tables, d, self._counts)
            _torch().cuda.current_stream().synchronize()
            self._pending_n = 0

    def collect_block(self, xyz_dev, cells=None):
        """Bin a whole block of frames already in HBM: CUDA float32 [n][N][3] of the selected atoms.
        `cells` (CUDA int32 [n][N]) optionally receives the brick index of every atom-frame, which
        HoppingTrajectory reads instead of the coordinates."""
        from . import kernels as K

        K.hist_accumulate(self._tables, xyz_dev, self._counts, cells)

    def finish(self, n_frames):
        if self.isComplete():
            return
        self._flush_pending()
        self._n_frames = int(n_frames)
        self._complete = True

    def Density(self):
        if self._counts is None:
            errmsg = "DensityCollector.Density(): No data for density available. Run collect() first."
            logger.error(errmsg)
            raise MissingDataError(errmsg)
        u = self.universe
        metadata = self.metadata
        metadata['collector'] = self.name
        metadata['collector_mode'] = self.mode
        metadata['psf'] = u.filename
        metadata['dcd'] = u.trajectory.filename
        metadata['atomselection'] = self.atomselection
        metadata['n_frames'] = u.trajectory.n_frames
        metadata['dt'] = u.trajectory.dt
        metadata['totaltime'] = u.trajectory.totaltime
        if self.mode == 'BULK':
            metadata['soluteselection'] = self.soluteselection
            metadata['cutoff'] = self.cutoff
        parameters = self.parameters
        parameters['isDensity'] = False  # the histogram is turned into a density on the device
        n_frames = self._n_frames if self._complete else 1
        g = Density._from_device_counts(self._counts, self.edges, n_frames, parameters=parameters,
                                        unit=dict(length='Angstrom'), metadata=metadata)
        logger.info("%-10s: Histogram completed (initial density in %s**-3)", self.name, 'Angstrom')
        return g

    def isComplete(self):
        return self._complete

    def __repr__(self):
        if self.mode == "BULK":
            return "<DensityCollector %(name)r, delta=%(delta).1f A: " \
                   "'%(atomselection)s and not around %(cutoff).1f (%(soluteselection)s)'>" % vars(self)
        return "<DensityCollector %(name)r, delta=%(delta).1f A: %(atomselection)r>" % vars(self)


class DensityCreator(object):
    """Create density grids from a trajectory (hop/density.py:180-356).

        DensityCreator(universe | (topology, trajectory), mode='all'|'bulk'|'solvent', delta=1.0,
                       atomselection='name OH2', padding=2.0, soluteselection=..., cutoff=3.5)
    """
    modes = ("all", "bulk", "solvent")
    defaults = {'cutoff': 3.5, 'soluteselection': "protein and not name H*", 'delta': 1.0,
                'atomselection': "name OH2", 'padding': 2.0}

    def __init__(self, *args, **kwargs):
        _kwargs = self.defaults.copy()
        _kwargs.update(kwargs)
        kwargs = _kwargs
        universe_kwargs = {'permissive': kwargs.pop('permissive', False)}
        self.universe = as_Universe(*args, **universe_kwargs)
        self.mode = kwargs.pop("mode", "all")
        if self.mode not in self.modes:
            errmsg = "DensityCreator: mode must be one of %r, not %r" % (self.modes, self.mode)
            logger.fatal(errmsg)
            raise ValueError(errmsg)
        self.keep_coordinates = kwargs.pop("keep_coordinates", True)
        wanted = list(self.modes[1:]) if self.mode == "all" else [self.mode]
        # one collector per density; the plain solvent density ignores the solute (hop/density.py:259-270)
        overrides = {"solvent": {'soluteselection': None, 'cutoff': 0}}
        self.collectors = [DensityCollector(m, self.universe, **dict(kwargs, **overrides.get(m, {}))) for m in wanted]
        # all densities of a run share ONE grid: the box around every collector's frame-0 extent (hop/density.py:271-278)
        lower = numpy.min([c.min_coordinates() for c in self.collectors], axis=0)
        upper = numpy.max([c.max_coordinates() for c in self.collectors], axis=0)
        for c in self.collectors:
            c.init_histogram(smin=lower, smax=upper)
        self.densities = {}

    def create(self):
        """Loop through the trajectory and build all densities (hop/density.py:281-304)."""
        u = self.universe
        traj = u.trajectory
        fast = [c for c in self.collectors if hasattr(traj, "coordinate_array")]
        slow = [c for c in self.collectors if c not in fast]
        if fast:
            self._create_blocks(fast)
        if slow:
            for ts in traj:
                for c in slow:
                    c.collect()
        self.densities = {}
        for c in self.collectors:
            c.finish(traj.n_frames)
            self.densities[c.name] = c.Density()
        return self.densities

    def _create_blocks(self, collectors):
        """Whole frame blocks: pinned staging -> HBM -> histogram kernel, copy overlapped with compute.
        A bulk collector is fused with the solvent collector of the same atom selection: one read of
        the coordinates feeds both histograms (csrc/hist.cu, hist_bulk_kernel).  The device copies are
        kept (memory permitting) so HoppingTrajectory can reuse them."""
        from . import kernels as K

        t = _torch()
        traj = self.universe.trajectory
        F = traj.n_frames
        done = set()
        for c in collectors:
            if id(c) in done or len(c.group.indices) == 0:
                continue
            idx = c.group.indices
            N = len(idx)
            # partner collector on the same atoms (mode='all': 'bulk' + 'solvent')
            partner = None
            for o in collectors:
                if o is not c and id(o) not in done and o.mode != c.mode and numpy.array_equal(o.group.indices, idx):
                    partner = o
            bulk = c if c.mode == "BULK" else (partner if partner is not None and partner.mode == "BULK" else None)
            solv = c if c.mode == "SOLVENT" else (partner if partner is not None and partner.mode == "SOLVENT" else None)
            done.add(id(c))
            if partner is not None:
                done.add(id(partner))
            total_bytes = F * N * 16
            free, _ = t.cuda.mem_get_info()
            keep = self.keep_coordinates and total_bytes < 0.4 * free
            stager = _FrameStager(N, keep_on_device=keep, n_frames_hint=F)
            cells_blocks = []
            pidx = bulk.solute_group.indices if bulk is not None else None
            state = {"f": 0}

            def consume(d, bulk=bulk, solv=solv, pidx=pidx, state=state):
                cells = t.empty(d.shape[:2], dtype=t.int32, device=d.device) if keep else None
                if bulk is None:
                    solv.collect_block(d, cells)
                else:
                    f0 = state["f"]
                    pblock = numpy.ascontiguousarray(numpy.take(traj.coordinate_array(f0, f0 + d.shape[0]), pidx, axis=1),
                                                     dtype=numpy.float32)
                    pd = t.from_numpy(pblock).cuda()
                    K.hist_bulk(bulk._tables, d, pd, bulk.cutoff, bulk._counts,
                                counts_solvent=None if solv is None else solv._counts, cells=cells)
                if keep:
                    cells_blocks.append(cells)

            identity = self.universe._is_identity(idx) if hasattr(self.universe, "_is_identity") else False
            f = 0
            while f < F:
                n = min(stager.block_frames, F - f)
                block = traj.coordinate_array(f, f + n)
                host = stager.host_block()
                if identity:
                    host[:n] = block
                else:
                    numpy.take(block, idx, axis=1, out=host[:n])
                state["f"] = f
                stager.flush(n, consume)
                f += n
            t.cuda.current_stream().synchronize()
            if keep:
                cache = getattr(traj, "_hopb_device_cache", None)
                if cache is None:
                    cache = traj._hopb_device_cache = {}
                edges = (solv or bulk).edges
                cache[idx.tobytes()] = {"xyz": stager.keep, "cells": cells_blocks,
                                        "edges": [numpy.array(e) for e in edges]}

    def DensityWithBulk(self, density_unit='water', solvent_threshold=numpy.e, bulk_threshold=0.6):
        """Solvent density with the bulk site inserted (hop/density.py:306-356)."""
        problem = None
        if len(self.densities) != 2:
            problem = "DensityCreator.DensityWithBulk(): Need exactly two densities ('solvent' and 'bulk')."
        elif not {'solvent', 'bulk'} <= set(self.densities):
            problem = "Need a 'solvent' and a 'bulk' density in %s.densities" % self.__class__.__name__
        if problem:
            logger.fatal(problem)
            raise MissingDataError(problem)
        solvent, bulk = self.densities['solvent'], self.densities['bulk']
        for dens, threshold in ((solvent, solvent_threshold), (bulk, bulk_threshold)):
            dens.convert_density(density_unit)
            dens.map_sites(threshold)
        solvent.site_insert_bulk(bulk)
        return solvent


def density_from_Universe(universe, delta=1.0, atomselection='name OH2', metadata=None, padding=2.0,
                          cutoff=0, soluteselection=None, use_kdtree=True, **kwargs):
    """Create a :class:`hop_b200.sitemap.Density` from a Universe (hop/density.py:358-369; the
    arithmetic is that of DensityCollector)."""
    mode = "bulk" if (cutoff > 0 and soluteselection is not None) else "solvent"
    DC = DensityCreator(universe, mode=mode, delta=delta, atomselection=atomselection, padding=padding,
                        cutoff=cutoff, soluteselection=soluteselection, metadata=metadata or {}, **kwargs)
    return DC.create()[mode]


def density_from_trajectory(*args, **kwargs):
    """density_from_trajectory(PSF, DCD, delta=1.0, atomselection='name OH2', ...) --> density
    (hop/density.py:372-430)."""
    kwargs.pop('verbosity', None)
    u = as_Universe(*args, permissive=kwargs.pop('permissive', False))
    return density_from_Universe(u, **kwargs)
