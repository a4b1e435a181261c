"""CHARMM/NAMD DCD reader and writer, and the minimal PSF (EXT) reader.

The hopping trajectory is stored "as a standard dcd file, together with a minimal psf"
(hop/trajectory.py:255-297): X record = site label, Y record = orbit site, Z record = 0, preceded
per frame by a fake unit cell [nlabels, nlabels, 1, 90, 90, 90] (trajectory.py:180-184).  The
reference delegates the byte layout to MDAnalysis.Writer, which is not available here, so the
format (little-endian, 32-bit Fortran record markers, CHARMM 24 header with unit cell) is
implemented directly.  Per frame the file holds three contiguous float32 planes of N values --
exactly the [frame][plane][atom] layout the GPU kernel writes.
"""
from __future__ import annotations

import os
import struct

import numpy as np

AKMA_PS = 4.888821e-2  # 1 AKMA time unit in ps (MDAnalysis.units.timeUnit_factor)


class DCDWriter:
    def __init__(self, filename, n_atoms, dt=1.0, nsavc=1, istart=0, remarks="", unitcell=True):
        self.filename = filename
        self.n_atoms = int(n_atoms)
        self.dt = float(dt)
        self.nsavc = int(nsavc)
        self.istart = int(istart)
        self.unitcell = bool(unitcell)
        self.n_frames = 0
        self._fh = open(filename, "wb")
        self._write_header(remarks)

    def _write_header(self, remarks):
        icntrl = np.zeros(20, dtype="<i4")
        icntrl[0] = 0                      # NSET, patched on close
        icntrl[1] = self.istart
        icntrl[2] = self.nsavc
        icntrl[10] = 1 if self.unitcell else 0
        icntrl[19] = 24                    # CHARMM version
        delta = np.float32(self.dt / self.nsavc / AKMA_PS)
        icntrl[9] = np.frombuffer(delta.tobytes(), dtype="<i4")[0]
        f = self._fh
        f.write(struct.pack("<i4s", 84, b"CORD"))
        f.write(icntrl.tobytes())
        f.write(struct.pack("<i", 84))
        titles = [remarks[:80].ljust(80), "Created by hop_b200.dcd.DCDWriter".ljust(80)]
        size = 4 + 80 * len(titles)
        f.write(struct.pack("<ii", size, len(titles)))
        for t in titles:
            f.write(t.encode("ascii", "replace"))
        f.write(struct.pack("<i", size))
        f.write(struct.pack("<iii", 4, self.n_atoms, 4))

    def write_planes(self, x, y, z=None, dimensions=None):
        """Append one frame from separate float32 arrays (z=None writes zeros)."""
        n = self.n_atoms
        f = self._fh
        if self.unitcell:
            d = np.zeros(6) if dimensions is None else np.asarray(dimensions, dtype=np.float64)
            # CHARMM order: A, gamma, B, beta, alpha, C
            cell = np.array([d[0], d[5], d[1], d[4], d[3], d[2]], dtype="<f8")
            f.write(struct.pack("<i", 48))
            f.write(cell.tobytes())
            f.write(struct.pack("<i", 48))
        marker = struct.pack("<i", 4 * n)
        for plane in (x, y, z):
            f.write(marker)
            if plane is None:
                f.write(bytes(4 * n))
            else:
                a = np.ascontiguousarray(plane, dtype="<f4")
                if a.size != n:
                    raise ValueError("plane has %d values, expected %d" % (a.size, n))
                f.write(a.tobytes())
            f.write(marker)
        self.n_frames += 1

    def write_next_timestep(self, ts):
        pos = np.asarray(ts.positions if hasattr(ts, "positions") else ts._pos)
        self.write_planes(pos[:, 0], pos[:, 1], pos[:, 2], getattr(ts, "dimensions", None))

    write = write_next_timestep

    def close(self):
        if self._fh is None:
            return
        self._fh.seek(8)
        self._fh.write(struct.pack("<i", self.n_frames))
        self._fh.close()
        self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class DCDFile:
    """Random-access DCD reader (memory-mapped)."""

    def __init__(self, filename):
        self.filename = filename
        with open(filename, "rb") as f:
            head = f.read(92)
            if len(head) < 92 or struct.unpack("<i", head[:4])[0] != 84 or head[4:8] != b"CORD":
                raise IOError("%s is not a little-endian CHARMM DCD file" % filename)
            icntrl = np.frombuffer(head[8:88], dtype="<i4")
            self.n_frames = int(icntrl[0])
            self.istart = int(icntrl[1])
            self.nsavc = max(int(icntrl[2]), 1)
            self.has_unitcell = bool(icntrl[10])
            delta = np.frombuffer(icntrl[9:10].tobytes(), dtype="<f4")[0]
            self.dt = float(delta) * self.nsavc * AKMA_PS
            size = struct.unpack("<i", f.read(4))[0]
            block = f.read(size)
            ntitle = struct.unpack("<i", block[:4])[0]
            self.remarks = [block[4 + 80 * i: 84 + 80 * i].decode("ascii", "replace").rstrip() for i in range(ntitle)]
            f.read(4)
            _, self.n_atoms, _ = struct.unpack("<iii", f.read(12))
            self._data_start = f.tell()
        n = self.n_atoms
        self._cell_bytes = 56 if self.has_unitcell else 0
        self._frame_bytes = self._cell_bytes + 3 * (4 * n + 8)
        actual = (os.path.getsize(filename) - self._data_start) // self._frame_bytes if self._frame_bytes else 0
        if self.n_frames <= 0 or self.n_frames > actual:
            self.n_frames = int(actual)
        self._mm = np.memmap(filename, dtype=np.uint8, mode="r") if self.n_frames else None

    def _plane(self, frame, k):
        off = self._data_start + frame * self._frame_bytes + self._cell_bytes + k * (4 * self.n_atoms + 8) + 4
        return np.frombuffer(self._mm, dtype="<f4", count=self.n_atoms, offset=off)

    def planes(self, frame):
        if not 0 <= frame < self.n_frames:
            raise IndexError(frame)
        return self._plane(frame, 0), self._plane(frame, 1), self._plane(frame, 2)

    def dimensions(self, frame):
        if not self.has_unitcell:
            return np.zeros(6)
        off = self._data_start + frame * self._frame_bytes + 4
        c = np.frombuffer(self._mm, dtype="<f8", count=6, offset=off)
        return np.array([c[0], c[2], c[5], c[4], c[3], c[1]])

    def plane_array(self, k, start=0, stop=None):
        """[stop-start][N] float32 copy of coordinate plane k (0=x,1=y,2=z) for a frame range."""
        stop = self.n_frames if stop is None else stop
        out = np.empty((max(stop - start, 0), self.n_atoms), dtype=np.float32)
        for i, f in enumerate(range(start, stop)):
            out[i] = self._plane(f, k)
        return out

    def coordinate_array(self, start=0, stop=None):
        stop = self.n_frames if stop is None else stop
        out = np.empty((max(stop - start, 0), self.n_atoms, 3), dtype=np.float32)
        for i, f in enumerate(range(start, stop)):
            for k in range(3):
                out[i, :, k] = self._plane(f, k)
        return out


def write_dcd(filename, coordinates, dt=1.0, remarks="", dimensions=None):
    coordinates = np.asarray(coordinates, dtype=np.float32)
    with DCDWriter(filename, coordinates.shape[1], dt=dt, remarks=remarks) as w:
        for fr in coordinates:
            w.write_planes(fr[:, 0], fr[:, 1], fr[:, 2], dimensions)


def read_psf_atoms(filename):
    """Parse the !NATOM section of a PSF (standard or EXT) into a dict of arrays."""
    with open(filename) as f:
        lines = f.readlines()
    i = 0
    while i < len(lines) and "!NATOM" not in lines[i]:
        i += 1
    if i == len(lines):
        raise IOError("no !NATOM section in %s" % filename)
    n = int(lines[i].split()[0])
    rec = dict(index=[], segid=[], resid=[], resname=[], name=[], type=[], charge=[], mass=[])
    for line in lines[i + 1: i + 1 + n]:
        t = line.split()
        rec["index"].append(int(t[0]) - 1)
        rec["segid"].append(t[1])
        rec["resid"].append(int(t[2]))
        rec["resname"].append(t[3])
        rec["name"].append(t[4])
        rec["type"].append(t[5])
        rec["charge"].append(float(t[6]))
        rec["mass"].append(float(t[7]))
    return {k: np.array(v) for k, v in rec.items()}
