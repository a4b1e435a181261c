// Union-find pieces of the site labelling shared by the stand-alone kernels of label.cu and the
// one-launch site map of sitemap_fused.cu.  Every routine is "one warp per (x, y) row, lanes along z",
// run by `nwarps` warps of which the caller is number `warp0`; rowflag[row] & flagbit tells whether the
// row holds a cell of the mask at all.  No __restrict__ / read-only qualifiers here: inside the one-launch site
// map these arrays are written by earlier phases of the same kernel, so their loads must stay coherent.
#pragma once
#include "hopb_common.cuh"

namespace hopb_ccl {

constexpr unsigned kFull = 0xffffffffu;

// ---- 2. merge -------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_parent(const int* parent, int i) {
  return *reinterpret_cast<const volatile int*>(parent + i);
}

// find with path halving.  Every value ever stored in parent[x] is a member of x's set with a smaller
// index, and stores go through atomicMin, so parent[] only decreases: no cycles, and a union that
// races with the compression still sees a valid (if stale) ancestor.
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = ld_parent(parent, x);
  while (p != x) {
    const int gp = ld_parent(parent, p);
    if (gp != p) atomicMin(parent + x, gp);
    x = p;
    p = gp;
  }
  return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }  // link the larger root under the smaller
    const int old = atomicMin(parent + a, b);
    if (old == a) return;
    a = old;
  }
}

__device__ __forceinline__ void merge_rows(int* parent, const uint8_t* rowflag, unsigned flagbit, int nx, int ny,
                                           int nz, int phase, int64_t warp0, int64_t nwarps, int lane) {
  const int64_t n_rows = (int64_t)nx * ny;
  const int64_t sy = nz, sx = (int64_t)ny * nz;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    if (!(rowflag[row] & flagbit)) continue;
    const int y = (int)(row % ny);
    const int x = (int)(row / ny);
    const bool hx = x + 1 < nx, hy = y + 1 < ny, ly = y > 0;
    // neighbour rows without any mask cell cannot contribute an edge
    const bool ry = hy && (rowflag[row + 1] & flagbit);
    const bool rx = hx && (rowflag[row + ny] & flagbit);
    const bool rxy = hx && hy && (rowflag[row + ny + 1] & flagbit);
    if (!(ry | rx | rxy)) continue;
    const int64_t base = row * nz;
    for (int z0 = 0; z0 < nz; z0 += 32) {
      const int z = z0 + lane;
      if (z >= nz) continue;
      const int64_t c = base + z;
      if (parent[c] < 0) continue;
      const bool hz = z + 1 < nz, lz = z > 0;
      auto M = [&](int64_t idx) { return parent[idx] >= 0; };
      const bool m_up = hz && M(c + 1);
      const bool m_dn = lz && M(c - 1);
      // phase 0: heads of z-runs only -- they tie the rows of a blob into one tree (the skeleton);
      // phase 1 (after a flatten): every other cell; inside a blob both ends then already share a
      // root one step away, so the many unions around isolated holes cost two loads each
      if ((phase == 0) == m_dn) continue;
      const bool m_y = ry && M(c + sy);              // (0,1,0)
      const bool m_x = rx && M(c + sx);              // (1,0,0)
      const bool m_xy = rxy && M(c + sx + sy);       // (1,1,0)
      // An edge is skipped when its end points are already joined by edges that are themselves kept
      // or skipped by a rule lower in the (y, z) order -- so the argument is well founded:
      //   rule Z: (c, c+o), o_z = 0, follows from the same edge one cell down both z-runs;
      //   rule Y: the x-edge follows from the x-edge one row down in y plus the two y-edges;
      //   diagonals follow from any two-step path over in-plane / z edges.
      if (m_y && !(m_dn && M(c + sy - 1))) uf_union(parent, (int)c, (int)(c + sy));
      if (m_x && !(m_dn && M(c + sx - 1)) && !(ly && M(c - sy) && M(c + sx - sy))) uf_union(parent, (int)c, (int)(c + sx));
      if (m_xy && !m_y && !m_x && !(m_dn && M(c + sx + sy - 1))) uf_union(parent, (int)c, (int)(c + sx + sy));
      if (hz) {
        // +z partners: (0,1,1), (1,0,1), (1,1,1)
        if (ry && !m_y && !m_up && M(c + sy + 1)) uf_union(parent, (int)c, (int)(c + sy + 1));
        if (rx && !m_x && !m_up && M(c + sx + 1)) uf_union(parent, (int)c, (int)(c + sx + 1));
        // (1,1,1) is also reached through c+x then (0,1,1), or c+y then (1,0,1)
        if (rxy && !m_xy && !m_up && !m_x && !m_y && M(c + sx + sy + 1)) uf_union(parent, (int)c, (int)(c + sx + sy + 1));
      }
    }
  }
}

// ---- 3. flatten -----------------------------------------------------------------------------------
// Only heads of z-runs are ever parents (initial links, unions and path halving all point at
// heads), so flattening is: heads find their root (few cells, possibly long walks), then every
// other cell takes its head's root in one hop.
__device__ __forceinline__ void flatten_rows(int* parent, const uint8_t* rowflag, unsigned flagbit, int64_t n_rows,
                                             int nz, int heads, int64_t warp0, int64_t nwarps, int lane) {
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    if (!(rowflag[row] & flagbit)) continue;
    const int64_t base = row * nz;
    // two chunks of 32 cells per step, all four loads issued before anything depends on them
    for (int z0 = 0; z0 < nz; z0 += 64) {
      const int za = z0 + lane, zb = za + 32;
      const int pa = za < nz ? ld_parent(parent, (int)(base + za)) : -1;
      const int pb = zb < nz ? ld_parent(parent, (int)(base + zb)) : -1;
      const int qa = (za > 0 && za < nz) ? ld_parent(parent, (int)(base + za - 1)) : -1;   // the cell below in z
      const int qb = zb < nz ? ld_parent(parent, (int)(base + zb - 1)) : -1;
      const bool ha = pa >= 0 && qa < 0, hb = pb >= 0 && qb < 0;       // heads of z-runs (z == 0: qa = -1)
      if (heads) {
        if (ha) { const int r = uf_find(parent, (int)(base + za)); parent[base + za] = r; }
        if (hb) { const int r = uf_find(parent, (int)(base + zb)); parent[base + zb] = r; }
      } else {
        // a non-head's parent is a head, flattened by the pass before: one more hop gives the root
        const bool na = pa >= 0 && !ha, nb = pb >= 0 && !hb;
        const int ra = na ? ld_parent(parent, pa) : 0;
        const int rb = nb ? ld_parent(parent, pb) : 0;
        if (na) parent[base + za] = ra;
        if (nb) parent[base + zb] = rb;
      }
    }
  }
}

// ---- 4. sizes ---------------------------------------------------------------------------------------
// size[root] += cells with that root.  Lanes with equal roots are merged first (neighbouring z cells
// share a root) and lane 0 carries the running (root, count) of the first labelled cell of every chunk
// across the chunks of a row AND across the rows of its warp: the one giant component of a bulk map then costs one atomic per warp instead of
// one per row (40 000 atomics on a single address were most of this kernel's time).
__device__ __forceinline__ void size_rows(const int* parent, const uint8_t* rowflag, unsigned flagbit,
                                          int64_t n_rows, int nz, int* size, int64_t warp0, int64_t nwarps, int lane) {
  int run_root = -1, run_cnt = 0;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    if (!(rowflag[row] & flagbit)) continue;
    const int64_t base = row * nz;
    for (int z0 = 0; z0 < nz; z0 += 32) {
      const int z = z0 + lane;
      const int r = z < nz ? parent[base + z] : -1;
      const unsigned valid = __ballot_sync(kFull, r >= 0);
      if (valid == 0) continue;           // nothing but background here (e.g. the padding of the grid)
      const unsigned peers = __match_any_sync(kFull, r);
      const bool leader = (peers & ((1u << lane) - 1)) == 0;
      const int f = __ffs(valid) - 1;     // the chunk's first labelled cell continues the running root
      const int rf = __shfl_sync(kFull, r, f);
      const int nf = __popc(__shfl_sync(kFull, peers, f));
      if (leader && r >= 0 && lane != f) atomicAdd(size + r, __popc(peers));
      if (lane == 0) {
        if (rf == run_root) run_cnt += nf;
        else {
          if (run_root >= 0) atomicAdd(size + run_root, run_cnt);
          run_root = rf; run_cnt = nf;
        }
      }
    }
  }
  if (lane == 0 && run_root >= 0) atomicAdd(size + run_root, run_cnt);
}


// ---- head-driven merge (sitemap_fused.cu) ----------------------------------------------------------------------
// Every union merge_rows performs has a z-run HEAD at one end: phase 0 starts from heads, and the pruning rules
// of phase 1 (rule Z: an in-plane edge follows from the same edge one cell down both runs; a +z partner is only
// tried when the cell below it is outside the mask) leave exactly the edges whose TARGET is a head.  With the
// heads in a compact list (a solid blob with holes has ~2 per row) one lane per head does the same unions:
// forward from the head (= phase 0) and backward from the head to non-head sources (= phase 1).  Same union
// set, hence the same components; roots are minimum indices either way.
// `M(x, y, z)` tells whether a cell is in the mask and answers false outside the grid without branching, so that
// the nine look-ups of a head are independent loads in flight together (one L2 latency instead of nine);
// `M.head(x, y, z)` is the head of the z-run a mask cell belongs to -- only heads carry union-find state.
template <class Mask, class Union>
__device__ __forceinline__ void merge_head_forward(int* parent, int c, int nx, int ny, int nz, const Mask& M, const Union& uf_union) {
  const int z = c % nz;
  const int row = c / nz;
  const int y = row % ny, x = row / ny;
  const bool m_up = M(x, y, z + 1);
  const bool m_y = M(x, y + 1, z);
  const bool m_x = M(x + 1, y, z);
  const bool m_xy = M(x + 1, y + 1, z);
  const bool b_y = M(x, y - 1, z), b_xy = M(x + 1, y - 1, z);
  const bool u_y = M(x, y + 1, z + 1), u_x = M(x + 1, y, z + 1), u_xy = M(x + 1, y + 1, z + 1);
  if (m_y) uf_union(parent, c, M.head(x, y + 1, z));
  if (m_x && !(b_y && b_xy)) uf_union(parent, c, M.head(x + 1, y, z));
  if (m_xy && !m_y && !m_x) uf_union(parent, c, M.head(x + 1, y + 1, z));
  if (!m_up) {
    if (!m_y && u_y) uf_union(parent, c, M.head(x, y + 1, z + 1));
    if (!m_x && u_x) uf_union(parent, c, M.head(x + 1, y, z + 1));
    if (!m_xy && !m_x && !m_y && u_xy) uf_union(parent, c, M.head(x + 1, y + 1, z + 1));
  }
}

// t is a head; the sources c = t - o that merge_rows' phase 1 would join to it.  In-plane offsets: the source
// has the head's z and must itself not be a head (the cell below it is in the mask).  +z partners: the source
// sits one cell lower; the cell below t is outside the mask because t is a head, which is the !m_y / !m_x /
// !m_xy of the forward rules.
template <class Mask, class Union>
__device__ __forceinline__ void merge_head_backward(int* parent, int t, int nx, int ny, int nz, const Mask& M, const Union& uf_union) {
  const int z = t % nz;
  const int row = t / nz;
  const int y = row % ny, x = row / ny;
  const bool y0 = M(x, y - 1, z), y1 = M(x, y - 1, z - 1), y2 = M(x, y - 1, z - 2);
  const bool x0 = M(x - 1, y, z), x1 = M(x - 1, y, z - 1), x2 = M(x - 1, y, z - 2);
  const bool d0 = M(x - 1, y - 1, z), d1 = M(x - 1, y - 1, z - 1), d2 = M(x - 1, y - 1, z - 2);
  if (y0 && y1) uf_union(parent, M.head(x, y - 1, z), t);                                   // (0, 1, 0)
  if (x0 && x1 && !(d0 && y0)) uf_union(parent, M.head(x - 1, y, z), t);                    // (1, 0, 0): not when the x-edge one row down in y exists
  if (d0 && d1 && !x0 && !y0) uf_union(parent, M.head(x - 1, y - 1, z), t);                // (1, 1, 0)
  if (y1 && y2 && !y0) uf_union(parent, M.head(x, y - 1, z - 1), t);                        // (0, 1, 1)
  if (x1 && x2 && !x0) uf_union(parent, M.head(x - 1, y, z - 1), t);                        // (1, 0, 1)
  if (d1 && d2 && !d0 && !y1 && !x1) uf_union(parent, M.head(x - 1, y - 1, z - 1), t);     // (1, 1, 1)
}


// ---- union-find with randomised linking ------------------------------------------------------------------------
// Linking the larger INDEX under the smaller (uf_union) makes the root the canonical minimum cell for free, but on
// the regular structures of a solvent box -- thousands of row heads joined along y, planes joined along x -- it
// builds chains hundreds of links long that every later find has to walk (70 us of dependent L2 round trips at
// 201^3, whatever the number of rows).  Linking by a hashed priority instead (the root with the larger hash goes
// under the other; the hash is a bijection, so the order is total and there are no cycles) gives trees of
// logarithmic expected height.  The minimum cell of a component is then collected separately (it is always a run
// head).  Links and path halving go through compare-and-swap: a link only ever replaces parent[a] == a.
__device__ __forceinline__ uint32_t uf_prio(int i) {
  uint32_t x = (uint32_t)i;
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ int uf_find_h(int* parent, int x) {
  int p = ld_parent(parent, x);
  while (p != x) {
    const int gp = ld_parent(parent, p);
    if (gp != p) atomicCAS(parent + x, p, gp);   // path halving; losing the race is harmless
    x = p;
    p = gp;
  }
  return x;
}

struct UnionHashed {
  __device__ __forceinline__ void operator()(int* parent, int a, int b) const {
    while (true) {
      a = uf_find_h(parent, a);
      b = uf_find_h(parent, b);
      if (a == b) return;
      if (uf_prio(a) < uf_prio(b)) { const int t = a; a = b; b = t; }   // a: the larger priority value, goes under b
      if (atomicCAS(parent + a, a, b) == a) return;
    }
  }
};

struct UnionMinIndex {
  __device__ __forceinline__ void operator()(int* parent, int a, int b) const { uf_union(parent, a, b); }
};

}  // namespace hopb_ccl
