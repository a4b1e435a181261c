"""Minimal OpenDX reader/writer for regular 3-D grids (what hop exports through gridData.OpenDX,
hop/sitemap.py:311-345): data in C order, z fastest, three values per line."""
from __future__ import annotations

import numpy as np


def write_dx(filename, grid, origin, delta, comments=()):
    grid = np.asarray(grid, dtype=np.float64)
    nx, ny, nz = grid.shape
    with open(filename, "w") as f:
        for c in comments:
            f.write("# " + str(c) + "\n")
        f.write("object 1 class gridpositions counts %d %d %d\n" % (nx, ny, nz))
        f.write("origin %.8f %.8f %.8f\n" % tuple(origin))
        f.write("delta %.8f 0 0\n" % delta[0])
        f.write("delta 0 %.8f 0\n" % delta[1])
        f.write("delta 0 0 %.8f\n" % delta[2])
        f.write("object 2 class gridconnections counts %d %d %d\n" % (nx, ny, nz))
        f.write("object 3 class array type double rank 0 items %d data follows\n" % grid.size)
        flat = grid.ravel()
        full = (len(flat) // 3) * 3
        if full:
            np.savetxt(f, flat[:full].reshape(-1, 3), fmt="%.9g")
        if full < len(flat):
            f.write(" ".join("%.9g" % v for v in flat[full:]) + "\n")
        f.write('attribute "dep" string "positions"\n')
        f.write('object "density" class field\n')
        f.write('component "positions" value 1\n')
        f.write('component "connections" value 2\n')
        f.write('component "data" value 3\n')


def read_dx(filename):
    """Returns (grid float64 [nx,ny,nz], edges list of 3 arrays)."""
    shape = None
    origin = None
    deltas = []
    data = []
    n_items = None
    with open(filename) as f:
        for line in f:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            t = s.split()
            if t[0] == "object" and "gridpositions" in t:
                shape = tuple(int(v) for v in t[-3:])
            elif t[0] == "origin":
                origin = [float(v) for v in t[1:4]]
            elif t[0] == "delta":
                deltas.append([float(v) for v in t[1:4]])
            elif t[0] == "object" and "array" in t:
                n_items = int(t[t.index("items") + 1])
            elif n_items is not None and len(data) < n_items and (t[0][0].isdigit() or t[0][0] in "+-."):
                data.extend(float(v) for v in t)
    if shape is None or origin is None or len(deltas) != 3 or n_items is None:
        raise IOError("%s is not a regular-grid OpenDX file" % filename)
    grid = np.array(data[:n_items], dtype=np.float64).reshape(shape)
    d = [deltas[i][i] for i in range(3)]
    edges = [origin[i] - 0.5 * d[i] + d[i] * np.arange(shape[i] + 1) for i in range(3)]
    return grid, edges
