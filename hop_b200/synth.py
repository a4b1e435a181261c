"""Seeded synthetic solvated-box trajectories.

The same generator exists twice: here in numpy (host; parity tests, CPU baseline) and as the
CUDA kernel ``hopb_synth_generate`` (csrc/synth.cu; bench-sized inputs are produced directly in
HBM).  Both use only integer hashing and float32 add/mul/compare in a fixed order (no FMA, no
transcendental functions), so they produce bit-identical coordinates for the same parameters.

Model: a cubic box of side L = (N/rho)^(1/3) with origin 0.  A fraction ``p_site`` of the atoms
(every ``site_stride``-th) are *site-bound*: they sit on one of K site centres (a jittered cubic
lattice inside a central cube) with positional noise sigma_site, hop to another site or go on a
free excursion with probability ``p_hop`` per frame, and rebind with ``p_rebind``.  All other atoms
are *bulk*: uniform start outside the central cube, random walk with step sigma_bulk, reflected at
the box walls and rejected when a step would enter the cube.  No periodic wrapping -- the
reference has none (hop/density.py:21-28).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

RHO_WATER = 0.0334  # atoms / Angstrom^3

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)
_KS = np.uint64(0xD6E8FEB86659FD93)
_KF = np.uint64(0xC2B2AE3D27D4EB4F)


def _mix(z):
    z = (z + _M1)
    z = (z ^ (z >> np.uint64(30))) * _M2
    z = (z ^ (z >> np.uint64(27))) * _M3
    return z ^ (z >> np.uint64(31))


def hash64(seed, atom, frame, stream):
    """Counter-based hash shared with csrc/synth.cu (hopb_hash64)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * _KS + np.asarray(atom, dtype=np.uint64) * _M1
             + np.asarray(frame, dtype=np.uint64) * _KF + np.uint64(stream))
        return _mix(_mix(z))


def _u21(h, k):
    """k-th 21-bit field of h as float32 in [0,1)."""
    v = (h >> np.uint64(21 * k)) & np.uint64(0x1FFFFF)
    return v.astype(np.float32) * np.float32(2.0 ** -21)


def _gauss(h):
    """Unit-variance, zero-mean Irwin-Hall(3) variate, float32, fixed operation order."""
    s = (_u21(h, 0) + _u21(h, 1)) + _u21(h, 2)
    return (s - np.float32(1.5)) * np.float32(2.0)


@dataclass
class SynthParams:
    n_atoms: int
    n_frames: int
    seed: int = 20261019
    rho: float = RHO_WATER
    site_stride: int = 50          # every 50th atom is site-bound (p_site = 0.02)
    site_spacing: float = 3.5      # Angstrom, lattice constant of the site centres
    sigma_site: float = 0.35
    sigma_bulk: float = 1.0
    p_hop: float = 0.01
    p_rebind: float = 0.02
    box: float = field(default=0.0)

    def __post_init__(self):
        if self.box <= 0.0:
            self.box = float(np.float32((self.n_atoms / self.rho) ** (1.0 / 3.0)))

    @property
    def n_bound(self):
        return (self.n_atoms + self.site_stride - 1) // self.site_stride

    @property
    def n_sites(self):
        return max(1, int(math.ceil(1.25 * self.n_bound)))

    def lattice(self):
        """(K,3) float32 site centres and the half-width of the central exclusion cube."""
        K = self.n_sites
        m = int(math.ceil(K ** (1.0 / 3.0) - 1e-9))
        L = np.float32(self.box)
        c = L * np.float32(0.5)
        idx = np.arange(K)
        ijk = np.stack([idx // (m * m), (idx // m) % m, idx % m], axis=1).astype(np.float32)
        base = (ijk - np.float32(0.5 * (m - 1))) * np.float32(self.site_spacing) + c
        jit = np.stack([_u21(hash64(self.seed, idx, 0, 100 + d), 0) for d in range(3)], axis=1)
        centres = (base + (jit - np.float32(0.5)) * np.float32(1.0)).astype(np.float32)
        half = np.float32(0.5 * (m - 1) * self.site_spacing + 1.5)
        return centres, float(half)


def generate(params: SynthParams, frame_start=0, frame_stop=None):
    """Return float32 coordinates [F][N][3] for frames [frame_start, frame_stop).

    The dynamics are sequential in time, so frames before ``frame_start`` are simulated and
    discarded."""
    p = params
    stop = p.n_frames if frame_stop is None else frame_stop
    N = p.n_atoms
    L = np.float32(p.box)
    centres, half = p.lattice()
    K = len(centres)
    c = L * np.float32(0.5)
    lo = np.float32(c - np.float32(half))
    hi = np.float32(c + np.float32(half))
    atoms = np.arange(N)
    bound = (atoms % p.site_stride) == 0
    seed = p.seed
    f32 = np.float32

    # ---- frame 0 ------------------------------------------------------------------------------
    pos = np.stack([_u21(hash64(seed, atoms, 0, 10 + d), 0) * L for d in range(3)], axis=1).astype(f32)
    inside = np.all((pos > lo) & (pos < hi), axis=1)
    shift = np.where(pos[:, 0] < c, -f32(half), f32(half)).astype(f32)
    pos[:, 0] = np.where(inside, pos[:, 0] + shift, pos[:, 0])
    site = np.full(N, -1, dtype=np.int64)
    site[bound] = (hash64(seed, atoms[bound], 0, 20) >> np.uint64(11)) % np.uint64(K)
    if N > 2:  # frame 0 spans the box: it defines the grid (hop/density.py:271-278)
        pos[1] = 0.0
        pos[2] = L * f32(1.0 - 2.0 ** -20)

    def bound_positions(f):
        out = pos.copy()
        b = np.nonzero(site >= 0)[0]
        for d in range(3):
            g = _gauss(hash64(seed, b, f, 30 + d))
            out[b, d] = centres[site[b], d] + g * f32(p.sigma_site)
        return out

    frames = []
    cur = bound_positions(0)
    pos = cur
    if frame_start == 0 and stop > 0:
        frames.append(cur.copy())
    for f in range(1, stop):
        # free movers: bulk atoms and bound atoms on an excursion
        free = site < 0
        step = np.stack([_gauss(hash64(seed, atoms, f, 40 + d)) * f32(p.sigma_bulk) for d in range(3)], axis=1)
        new = (pos + step).astype(f32)
        new = np.where(new < 0, -new, new)
        new = np.where(new >= L, (L + L) - new, new)
        new = np.minimum(new, L * f32(1.0 - 2.0 ** -20)).astype(f32)
        new = np.maximum(new, f32(0.0))
        in_cube = np.all((new > lo) & (new < hi), axis=1)
        reject = in_cube & ~bound            # bulk atoms never enter the central cube
        moved = np.where((free & ~reject)[:, None], new, pos)
        # site dynamics for the bound population
        u = _u21(hash64(seed, atoms, f, 50), 0)
        pick = ((hash64(seed, atoms, f, 51) >> np.uint64(11)) % np.uint64(K)).astype(np.int64)
        on = bound & (site >= 0)
        off = bound & (site < 0)
        jump = on & (u < f32(0.5 * p.p_hop))
        leave = on & ~jump & (u < f32(p.p_hop))
        rebind = off & (u < f32(p.p_rebind))
        site = np.where(jump | rebind, pick, site)
        site = np.where(leave, -1, site)
        pos = moved.astype(f32)
        b = np.nonzero(site >= 0)[0]
        for d in range(3):
            g = _gauss(hash64(seed, b, f, 30 + d))
            pos[b, d] = centres[site[b], d] + g * f32(p.sigma_site)
        if f >= frame_start:
            frames.append(pos.copy())
    if not frames:
        return np.zeros((0, N, 3), f32)
    return np.stack(frames).astype(f32)
