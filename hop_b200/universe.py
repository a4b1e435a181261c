"""A minimal, duck-typed stand-in for the parts of MDAnalysis the hop hot path touches.

hop takes an ``MDAnalysis.Universe`` (hop/density.py:251, hop/trajectory.py:91-211).  MDAnalysis
is not a dependency of this package; a real Universe works wherever the attributes used below
exist (``select_atoms``, ``atoms``, ``trajectory.ts / n_frames / dt / totaltime / filename``,
``AtomGroup.positions``).  For trajectories that live in memory (or in a DCD read through
hop_b200.dcd) the classes here additionally expose ``trajectory.coordinate_array()`` so whole frame
ranges can be staged to the GPU without a Python loop over frames.
"""
from __future__ import annotations

import fnmatch
import re

import numpy as np

from .dcd import DCDFile, read_psf_atoms
from .exceptions import SelectionError


class Timestep:
    """Coordinates of one frame (x, y, z columns of a float32 [N,3] array)."""

    def __init__(self, n_atoms):
        self.n_atoms = int(n_atoms)
        self.frame = 0
        self._pos = np.zeros((self.n_atoms, 3), dtype=np.float32)
        self.dimensions = np.zeros(6, dtype=np.float32)

    @property
    def positions(self):
        return self._pos

    @positions.setter
    def positions(self, value):
        self._pos[:] = value

    @property
    def _x(self):
        return self._pos[:, 0]

    @property
    def _y(self):
        return self._pos[:, 1]

    @property
    def _z(self):
        return self._pos[:, 2]


class _TrajectoryBase:
    """Iteration protocol of an MDAnalysis reader (>= 0.11): iterating re-starts from frame 0 and
    updates ``ts`` in place."""

    n_frames = 0
    n_atoms = 0
    dt = 1.0
    filename = None

    def _load(self, frame):
        raise NotImplementedError

    @property
    def totaltime(self):
        return (self.n_frames - 1) * self.dt

    def rewind(self):
        self._load(0)
        return self.ts

    def next(self):
        nxt = self.ts.frame + 1
        if nxt >= self.n_frames:
            raise StopIteration
        self._load(nxt)
        return self.ts

    __next__ = next

    def __len__(self):
        return self.n_frames

    def __iter__(self):
        for f in range(self.n_frames):
            self._load(f)
            yield self.ts

    def __getitem__(self, item):
        if isinstance(item, slice):
            rng = range(*item.indices(self.n_frames))

            def it():
                for f in rng:
                    self._load(f)
                    yield self.ts
            return it()
        f = int(item)
        if f < 0:
            f += self.n_frames
        if not 0 <= f < self.n_frames:
            raise IndexError(item)
        self._load(f)
        return self.ts


class MemoryTrajectory(_TrajectoryBase):
    """Trajectory held as one float32 [F][N][3] array."""

    def __init__(self, coordinates, dt=1.0, filename=None, dimensions=None):
        c = np.asarray(coordinates)
        if c.ndim != 3 or c.shape[2] != 3:
            raise ValueError("coordinates must have shape [n_frames][n_atoms][3]")
        self._coords = np.ascontiguousarray(c, dtype=np.float32)
        self.n_frames, self.n_atoms = self._coords.shape[:2]
        self.dt = float(dt)
        self.filename = filename
        self._dimensions = None if dimensions is None else np.asarray(dimensions, dtype=np.float32)
        self.ts = Timestep(self.n_atoms)
        if self.n_frames:
            self._load(0)

    def _load(self, frame):
        self.ts._pos = self._coords[frame]  # a view: no copy per frame
        self.ts.frame = frame
        if self._dimensions is not None:
            self.ts.dimensions = self._dimensions

    def coordinate_array(self, start=0, stop=None, step=1):
        return self._coords[start:stop:step]


class DCDTrajectory(_TrajectoryBase):
    """Trajectory read from a DCD file (coordinates converted to [N][3] per frame)."""

    def __init__(self, filename):
        self._dcd = DCDFile(filename)
        self.filename = filename
        self.n_frames = self._dcd.n_frames
        self.n_atoms = self._dcd.n_atoms
        self.dt = self._dcd.dt
        self.ts = Timestep(self.n_atoms)
        if self.n_frames:
            self._load(0)

    def _load(self, frame):
        x, y, z = self._dcd.planes(frame)
        pos = self.ts._pos
        if not pos.flags.writeable or pos.shape != (self.n_atoms, 3):
            pos = self.ts._pos = np.empty((self.n_atoms, 3), dtype=np.float32)
        pos[:, 0] = x
        pos[:, 1] = y
        pos[:, 2] = z
        self.ts.frame = frame
        self.ts.dimensions = self._dcd.dimensions(frame)

    def coordinate_array(self, start=0, stop=None, step=1):
        stop = self.n_frames if stop is None else stop
        a = self._dcd.coordinate_array(start, stop)
        return a[::step] if step != 1 else a

    def plane_array(self, k, start=0, stop=None):
        return self._dcd.plane_array(k, start, stop)


class Atom:
    __slots__ = ("index", "segid", "resid", "resname", "name", "type", "charge", "mass")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


_ATTRS = ("segid", "resid", "resname", "name", "type", "charge", "mass")


class AtomGroup:
    def __init__(self, universe, indices):
        self.universe = universe
        self.indices = np.asarray(indices, dtype=np.int64)

    @property
    def atoms(self):
        return self

    @property
    def n_atoms(self):
        return len(self.indices)

    def __len__(self):
        return len(self.indices)

    @property
    def positions(self):
        pos = self.universe.trajectory.ts.positions
        if len(self.indices) == pos.shape[0] and self.universe._is_identity(self.indices):
            return pos
        return pos[self.indices]

    def select_atoms(self, selection):
        sub = self.universe.select_atoms(selection)
        keep = np.intersect1d(self.indices, sub.indices)
        return AtomGroup(self.universe, keep)

    def __iter__(self):
        t = self.universe._topology
        for i in self.indices:
            yield Atom(index=int(i), **{a: t[a][i].item() if hasattr(t[a][i], "item") else t[a][i] for a in _ATTRS})

    def __getattr__(self, name):
        # names, resids, charges, masses, types, segids, resnames
        plural = {"names": "name", "resids": "resid", "charges": "charge", "masses": "mass", "types": "type",
                  "segids": "segid", "resnames": "resname"}
        if name in plural:
            return self.universe._topology[plural[name]][self.indices]
        raise AttributeError(name)


class Universe:
    """Universe(psf, dcd) | Universe(coordinates=array[, names=..., dt=...]).

    Only what the hop hot path needs: an atom table, a trajectory and ``select_atoms``."""

    def __init__(self, topology=None, trajectory=None, coordinates=None, names=None, resnames=None, resids=None,
                 segids=None, types=None, charges=None, masses=None, dt=1.0, filename=None, **kwargs):
        if coordinates is not None:
            self.trajectory = MemoryTrajectory(coordinates, dt=dt, filename=trajectory if isinstance(trajectory, str) else None)
            n = self.trajectory.n_atoms
            self.filename = filename if filename is not None else (topology if isinstance(topology, str) else None)
            topo = None
        elif isinstance(topology, str) and isinstance(trajectory, str):
            topo = read_psf_atoms(topology)
            self.trajectory = DCDTrajectory(trajectory)
            n = self.trajectory.n_atoms
            if len(topo["index"]) != n:
                raise ValueError("topology has %d atoms, trajectory %d" % (len(topo["index"]), n))
            self.filename = topology
        elif isinstance(trajectory, _TrajectoryBase):
            self.trajectory = trajectory
            n = trajectory.n_atoms
            topo = None
            self.filename = filename
        else:
            raise ValueError("Universe needs (psf, dcd) file names or coordinates=")

        def col(v, default, dtype=None):
            if v is None:
                return np.array([default] * n, dtype=dtype)
            a = np.asarray(v)
            if a.ndim == 0:
                a = np.array([a.item()] * n)
            if len(a) != n:
                raise ValueError("attribute column has the wrong length")
            return a

        if topo is None:
            topo = dict(
                segid=col(segids, "SYSTEM"), resid=col(resids, None) if resids is not None else np.arange(1, n + 1),
                resname=col(resnames, "TIP3"), name=col(names, "OH2"), type=col(types, "OT"),
                charge=col(charges, 0.0, float), mass=col(masses, 15.9994, float))
        self._topology = {k: np.asarray(v) for k, v in topo.items() if k != "index"}
        self.atoms = AtomGroup(self, np.arange(n))

    def _is_identity(self, indices):
        n = len(self.atoms.indices)
        return len(indices) == n and (n == 0 or (indices[0] == 0 and indices[-1] == n - 1))

    # ---- selections -----------------------------------------------------------------------------
    def select_atoms(self, selection):
        tokens = re.findall(r"\(|\)|[^\s()]+", selection)
        if not tokens:
            raise SelectionError("empty selection")
        pos = [0]
        n = len(self.atoms.indices)
        topo = self._topology
        keywords = {"name": "name", "resname": "resname", "segid": "segid", "type": "type", "resid": "resid"}
        stop_words = {"and", "or", "not", ")", "("}

        def peek():
            return tokens[pos[0]] if pos[0] < len(tokens) else None

        def take():
            t = peek()
            pos[0] += 1
            return t

        def parse_or():
            m = parse_and()
            while peek() == "or":
                take()
                m = m | parse_and()
            return m

        def parse_and():
            m = parse_not()
            while peek() == "and":
                take()
                m = m & parse_not()
            return m

        def parse_not():
            if peek() == "not":
                take()
                return ~parse_not()
            return parse_atom()

        def parse_atom():
            t = take()
            if t is None:
                raise SelectionError("unexpected end of selection %r" % selection)
            if t == "(":
                m = parse_or()
                if take() != ")":
                    raise SelectionError("missing ) in %r" % selection)
                return m
            if t == "all":
                return np.ones(n, dtype=bool)
            if t == "protein":
                return np.isin(topo["resname"], _PROTEIN_RESNAMES)
            if t in keywords:
                vals = []
                while peek() is not None and peek() not in stop_words and peek() not in keywords \
                        and peek() not in ("all", "index", "bynum", "protein"):
                    vals.append(take())
                if not vals:
                    raise SelectionError("keyword %r needs values" % t)
                colv = topo[keywords[t]]
                m = np.zeros(n, dtype=bool)
                for v in vals:
                    if t == "resid":
                        lo, _, hi = v.partition(":") if ":" in v else v.partition("-")
                        hi = hi or lo
                        m |= (colv >= int(lo)) & (colv <= int(hi))
                    else:
                        m |= np.array([fnmatch.fnmatchcase(str(c), v) for c in colv]) if any(ch in v for ch in "*?[") \
                            else (colv.astype(str) == v)
                return m
            if t in ("index", "bynum"):
                v = take()
                lo, _, hi = v.partition(":") if ":" in v else v.partition("-")
                hi = hi or lo
                off = 1 if t == "bynum" else 0
                idx = np.arange(n)
                return (idx >= int(lo) - off) & (idx <= int(hi) - off)
            raise SelectionError("cannot parse %r in selection %r" % (t, selection))

        mask = parse_or()
        if peek() is not None:
            raise SelectionError("trailing tokens in selection %r" % selection)
        return AtomGroup(self, np.nonzero(mask)[0])


_PROTEIN_RESNAMES = np.array(["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "HSD", "HSE", "HSP",
                              "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"])


def as_Universe(*args, **kwargs):
    """Return a universe from the input arguments (MDAnalysis.as_Universe)."""
    if len(args) == 0:
        raise TypeError("as_Universe() takes at least one argument")
    if hasattr(args[0], "select_atoms") and hasattr(args[0], "trajectory"):
        return args[0]
    kwargs.pop("permissive", None)
    return Universe(*args, **kwargs)
