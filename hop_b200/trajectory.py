"""Stage 3: the hopping trajectory -- drop-in for hop.trajectory.HoppingTrajectory.

Same surface as hop/trajectory.py:50-480: construct from ``(trajectory, group, density)`` or load
with ``hopdcd=/hoppsf=`` or ``filename=``; ``map_dcd(start, stop, step)``, iteration and ``next()``
yield Timesteps whose x column is the site label, y the orbit site and z = 0; ``write(filename)``
produces ``<name>.dcd`` + ``<name>.psf``.

The per-frame work of ``_coord2hop`` (three ``numpy.digitize`` calls, the rounded top-edge rule, a
copy of the whole map into a -1 padded buffer and a per-atom Python list comprehension,
hop/trajectory.py:403-468) runs as one CUDA kernel over whole frame blocks (csrc/hoptraj.cu); the
same pass also runs the transition state machine, so a TransportNetwork built from this object
needs no second pass over the frames.
"""
from __future__ import annotations

import logging

import numpy

from . import utilities
from .constants import SITELABEL
from .dcd import DCDWriter
from .universe import Timestep, Universe

logger = logging.getLogger("MDAnalysis.app.hop.trajectory")

#: frames mapped per kernel launch when streaming through map_dcd()
BLOCK_BYTES = 256 << 20


def _torch():
    import torch

    return torch


class HoppingTrajectory(object):
    """Time-sequence of sites visited by individual molecules (hop/trajectory.py:50-89).

    :Attributes: ``ts`` (Timestep), ``n_frames``, ``group``/``tgroup``, ``totaltime``, ``hoptraj``.
    """

    def __init__(self, trajectory=None, group=None, density=None, filename=None, hopdcd=None, hoppsf=None,
                 fixtrajectory=None, verbosity=3):
        self.verbosity = verbosity
        self._events_cache = None
        if not (trajectory is None or group is None or density is None):
            self.traj = trajectory
            try:
                self.tgroup = group.atoms
            except AttributeError:
                raise TypeError('group must be a <AtomGroup>, eg MDAnalyis.Universe.select_atoms().')
            if len(self.tgroup) == 0:
                raise ValueError("Group contains 0 particles, should be >0")
            if isinstance(fixtrajectory, dict):
                for attr, val in fixtrajectory.items():
                    if not hasattr(trajectory, attr):
                        raise AttributeError('fixtrajectory: dcd object does not have attribute "' + str(attr) + '"')
                    trajectory.__dict__[attr] = val
            self.totaltime = trajectory.totaltime
            self.traj.rewind()
            self._GD = density
            self.map = self._GD.map
            self.edges = self._GD.edges
            self.dedges = [numpy.diff(e) for e in self.edges]
            try:
                if not self._GD.grid.shape == self.map.shape:
                    raise ValueError
            except (AttributeError, ValueError):
                raise ValueError("The density object must have its site map computed.")
            coord = numpy.asarray(self.tgroup.positions)
            Natoms, D = coord.shape
            if not D == numpy.ndim(self.map):
                raise ValueError("Coordinates and map have different dimensions.")
            self._ts = Timestep(Natoms)
            self._ts.frame = self.traj.ts.frame
            numlabels = float(self.map.max() - self.map.min() + 2)
            self._ts.dimensions = numpy.array([numlabels, numlabels, 1, 90, 90, 90], dtype=numpy.float32)
            self.n_frames = self.traj.n_frames
            self.hoptraj = None
            self._tables = None
            self._map16 = None
            self._stream_state = None
            self._first = self._map_frames(0, 1)
            self._set_ts(0, self._first[0][0], self._first[1][0])
        elif not (hopdcd is None or hoppsf is None) or filename is not None:
            for a in ("traj", "tgroup", "map", "edges", "dedges"):
                if not hasattr(self, a):
                    setattr(self, a, None)
            if filename is not None:
                hoppsf = self.filename(filename, 'psf')
                hopdcd = self.filename(filename, 'dcd')
            u = Universe(hoppsf, hopdcd)
            if u.atoms.n_atoms == 0:
                raise ValueError("Hop trajectory contains 0 particles.")
            self.group = u.atoms
            self.hoptraj = u.trajectory
            self.n_frames = self.hoptraj.n_frames
            self.totaltime = self.hoptraj.totaltime
        else:
            raise ValueError('Not sufficient data to create a hopping trajectory.')

    filename = utilities.filename_function

    # ---- Timestep access ------------------------------------------------------------------------------
    @property
    def ts(self):
        """Timestep of the hoptraj"""
        if self.hoptraj:
            return self.hoptraj.ts
        return self._ts

    @ts.setter
    def ts(self, x):
        if self.hoptraj:
            raise AttributeError("Cannot modify the loaded hoptraj.ts")
        self._ts = x

    def _set_ts(self, frame, x, y):
        pos = self._ts._pos
        if not pos.flags.writeable:
            pos = self._ts._pos = numpy.zeros((len(x), 3), dtype=numpy.float32)
        pos[:, 0] = x
        pos[:, 1] = y
        pos[:, 2] = 0
        self._ts.frame = frame
        return self._ts

    def next(self):
        """Next time step of the hopping trajectory: read from the hoptraj file if one is loaded,
        otherwise mapped from the coordinate trajectory."""
        if self.hoptraj:
            return self.hoptraj.next()
        return self._map_next_timestep()

    __next__ = next

    def _map_next_timestep(self):
        nxt = self._ts.frame + 1
        if nxt >= self.n_frames:
            raise StopIteration
        if self._stream_state is None or self._stream_state[0] != nxt:
            # sequential access continues the orbit column from the frames mapped so far
            x, y = self._map_frames(0, nxt + 1)
            self._stream_state = (nxt + 1, y[-1].copy())
            return self._set_ts(nxt, x[-1], y[-1])
        x, y = self._map_frames(nxt, nxt + 1, orbit0=self._stream_state[1])
        self._stream_state = (nxt + 1, y[-1].copy())
        return self._set_ts(nxt, x[0], y[0])

    def _read_next_timestep(self):
        return self.hoptraj.next()

    # ---- GPU mapping ------------------------------------------------------------------------------------
    def _device_setup(self):
        from . import kernels as K

        if self._tables is None:
            self._tables = K.GridTables(self.edges)
            self._map16 = self._GD._device_map()
            if self._map16.numel() != self._tables.n_cells:
                raise ValueError("The density object must have its site map computed.")
            self._brick = K.brick_map(self._map16)

    @staticmethod
    def _slice_blocks(blocks, start, stop):
        t = _torch()
        out, f0 = [], 0
        for b in blocks:
            f1 = f0 + b.shape[0]
            lo, hi = max(start, f0), min(stop, f1)
            if lo < hi:
                out.append(b[lo - f0: hi - f0])
            f0 = f1
        return out[0] if len(out) == 1 else t.cat(out)

    def _cached(self):
        """Device copies DensityCreator left behind for this atom group (coordinates, cell indices)."""
        cache = getattr(self.traj, "_hopb_device_cache", None)
        if cache is None:
            return None
        return cache.get(numpy.asarray(self.tgroup.indices).tobytes())

    def _cells(self, start, stop, step=1):
        """Brick indices [n][N] written by the density pass, if they were computed on this very grid."""
        entry = self._cached()
        if entry is None or step != 1 or not entry["cells"]:
            return None
        if not all(numpy.array_equal(a, b) for a, b in zip(entry["edges"], self.edges)):
            return None
        return self._slice_blocks(entry["cells"], start, stop)

    def _coordinates(self, start, stop, step=1):
        """float32 [n][N][3] coordinates of the tracked group for a frame range, as CUDA tensor."""
        t = _torch()
        idx = numpy.asarray(self.tgroup.indices)
        entry = self._cached()
        if entry is not None and step == 1:
            return self._slice_blocks(entry["xyz"], start, stop)
        if hasattr(self.traj, "coordinate_array"):
            block = self.traj.coordinate_array(start, stop, step)
        else:
            frames = []
            for ts in self.traj[start:stop:step]:
                frames.append(numpy.array(self.tgroup.positions, dtype=numpy.float32))
            return t.from_numpy(numpy.stack(frames)).cuda()
        u = self.tgroup.universe
        identity = hasattr(u, "_is_identity") and u._is_identity(idx)
        if not identity:
            block = numpy.take(block, idx, axis=1)
        pinned = t.from_numpy(numpy.ascontiguousarray(block, dtype=numpy.float32)).pin_memory()
        return pinned.cuda(non_blocking=True)

    def _map_block(self, xyz_dev, frame0, state, events, hop_x=None, hop_y=None, cells_dev=None):
        """One slab through the fused kernel + fix-up.  `state` is carried from slab to slab."""
        from . import kernels as K

        src = cells_dev if cells_dev is not None else xyz_dev
        F, N = int(src.shape[0]), int(src.shape[1])
        t = _torch()
        if hop_x is None:
            hop_x = t.empty((F, N), dtype=t.float32, device=src.device)
            hop_y = t.empty((F, N), dtype=t.float32, device=src.device)
        n_chunks, chunk_len = K.plan_chunks(F, N)
        s = K.hoptraj(self._tables, xyz_dev, self._brick, frame0, n_chunks, chunk_len, events, hop_x, hop_y,
                      cells=None if cells_dev is None else cells_dev.contiguous())
        K.hop_fixup(s, chunk_len, F, state, hop_y, events)
        return hop_x, hop_y

    def _map_frames(self, start, stop, step=1, orbit0=None):
        """numpy (x, y) arrays [n][N] for frames start:stop:step of one pass that begins at `start`
        with orbit sites `orbit0` (default: interstitial, as _init_coord2hop does)."""
        from . import kernels as K

        self._device_setup()
        t = _torch()
        N = len(self.tgroup)
        frames = range(start, stop, step)
        n = len(frames)
        x = numpy.empty((n, N), dtype=numpy.float32)
        y = numpy.empty((n, N), dtype=numpy.float32)
        if n == 0:
            return x, y
        block = max(1, BLOCK_BYTES // (N * 12))
        state = K.new_state(N, "cuda")
        K.hop_advance(None, state, init=True)
        if orbit0 is not None:
            state[0].copy_(t.from_numpy(numpy.asarray(orbit0, dtype=numpy.int32)).cuda())
        events = K.EventBuffer(0, "cuda")
        done = 0
        while done < n:
            m = min(block, n - done)
            f0 = start + done * step
            cells = self._cells(f0, f0 + (m - 1) * step + 1, step)
            xyz = None if cells is not None else self._coordinates(f0, f0 + (m - 1) * step + 1, step)
            # frame numbers only matter to the (discarded) transition events here; starting at 1
            # keeps the start-of-trajectory rule for outliers out of a pass that begins mid-way
            hx, hy = self._map_block(xyz, done + 1, state, events, cells_dev=cells)
            x[done:done + m] = hx.cpu().numpy()
            y[done:done + m] = hy.cpu().numpy()
            done += m
        return x, y

    def map_dcd(self, start=None, stop=None, step=None):
        """Generator over the trajectory from start to stop that maps positions to sites
        (hop/trajectory.py:379-394).  The orbit column starts from interstitial at `start`."""
        s, e, st = slice(start, stop, step).indices(self.n_frames)
        N = len(self.tgroup)
        block = max(1, BLOCK_BYTES // (N * 12))
        frames = list(range(s, e, st))
        orbit = None
        for b0 in range(0, len(frames), block):
            sub = frames[b0:b0 + block]
            x, y = self._map_frames(sub[0], sub[-1] + 1, st, orbit0=orbit)
            orbit = y[-1].copy()
            for k, f in enumerate(sub):
                yield self._set_ts(f, x[k], y[k])

    def hop_arrays(self, with_events=False):
        """The whole hopping trajectory as numpy arrays (x, y), each [n_frames][N] float32.  With
        ``with_events`` the transition events of the same pass are kept for TransportNetwork."""
        from . import kernels as K

        if self.hoptraj:
            return self.hoptraj.plane_array(0), self.hoptraj.plane_array(1)
        self._device_setup()
        t = _torch()
        N = len(self.tgroup)
        F = self.n_frames
        x = t.empty((F, N), dtype=t.float32).pin_memory()
        y = t.empty((F, N), dtype=t.float32).pin_memory()
        block = max(1, BLOCK_BYTES // (N * 12))
        state = K.new_state(N, "cuda")
        K.hop_advance(None, state, init=True)
        cap = max(1 << 16, (F * N) // 8) if with_events else 0
        events = K.EventBuffer(cap, "cuda")
        for f0 in range(0, F, block):
            m = min(block, F - f0)
            cells = self._cells(f0, f0 + m)
            xyz = None if cells is not None else self._coordinates(f0, f0 + m)
            hx, hy = self._map_block(xyz, f0, state, events, cells_dev=cells)
            x[f0:f0 + m].copy_(hx, non_blocking=True)
            y[f0:f0 + m].copy_(hy, non_blocking=True)
        t.cuda.current_stream().synchronize()
        if with_events:
            n_ev = events.n()
            if n_ev > events.capacity:
                # rare: rerun with room for every event
                events = K.EventBuffer(n_ev, "cuda")
                state = K.new_state(N, "cuda")
                K.hop_advance(None, state, init=True)
                for f0 in range(0, F, block):
                    m = min(block, F - f0)
                    cells = self._cells(f0, f0 + m)
                    self._map_block(None if cells is not None else self._coordinates(f0, f0 + m), f0, state, events,
                                    cells_dev=cells)
                n_ev = events.n()
            self._events_cache = (events, n_ev, state)
        return x.numpy(), y.numpy()

    def write(self, filename, start=None, step=None, delta=None, load=True):
        """Write hopping trajectory as standard dcd file, together with a minimal psf
        (hop/trajectory.py:255-297).  start, step and delta are accepted and ignored exactly as in the
        reference, whose docstring says so (hop/trajectory.py:264-266: "Ignore the other options and leave
        them at the defaults. Currently, only the whole trajectory is written") and whose body never reads them."""
        psfname = self.filename(filename, 'psf')
        dcdname = self.filename(filename, 'dcd')
        x, y = self.hop_arrays()
        with DCDWriter(dcdname, len(self.tgroup), dt=self.traj.dt,
                       remarks='Hopping trajectory: x=site y=orbit_site z=0') as w:
            dims = self._ts.dimensions
            for f in range(x.shape[0]):
                w.write_planes(x[f], y[f], None, dims)
        logger.info("HoppingTrajectory.write(): wrote hoptraj %r.", dcdname)
        self.write_psf(psfname)
        logger.info("HoppingTrajectory.write(): wrote hoppsf %r.", psfname)
        if load is True:
            self.__init__(filename=filename, verbosity=self.verbosity)

    def write_psf(self, filename):
        """Write a dummy psf (EXT, header + !NATOM section) for the atoms in the selected group
        (hop/trajectory.py:299-376)."""
        fmt = '%(iatom)10d %(segid)8s %(resid)-8d %(resname)8s %(name)-8s %(type)4s %(charge)-14.6f%(mass)-14.4f%(imove)8d\n'
        with open(filename, 'w') as psf:
            psf.write('PSF EXT\n\n')
            psf.write('%7d !NTITLE\n' % 3)
            psf.write('* Hopping trajectory written by hop.trajectory.HoppingTrajectory.write()\n'
                      '* See https://github.com/Becksteinlab/hop\n'
                      '* This is NOT a fully functional psf but should work for visualization.\n')
            psf.write('\n')
            psf.write('%10d !NATOM\n' % len(self.tgroup))
            for atom in self.tgroup:
                if atom.resid >= 10 ** 8 or atom.index + 1 >= 10 ** 10:
                    raise NotImplementedError("Sorry, too many atoms (%d) or resids (%d) for the standard PSF format."
                                              % (atom.index + 1, atom.resid))
                psf.write(fmt % {'iatom': atom.index + 1, 'segid': str(atom.segid)[:8], 'resid': atom.resid,
                                 'resname': str(atom.resname)[:8], 'name': str(atom.name)[:8], 'type': atom.type,
                                 'charge': atom.charge, 'mass': atom.mass, 'imove': 0})

    def iterator(self):
        return self.__iter__()

    def __iter__(self):
        if self.hoptraj:
            for ts in self.hoptraj:
                yield ts
        else:
            for ts in self.map_dcd():
                yield ts
