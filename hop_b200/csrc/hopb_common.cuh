// Internal definitions shared by the translation units of libhopb200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/hopb200.h"
#include "hopb_binning.h"
#include "hopb_machine.h"

#define HOPB_NUM_SCRATCH 20

struct hopb_handle {
  int device;
  int num_sms;
  char err[512];
  void* scratch[HOPB_NUM_SCRATCH];
  size_t scratch_bytes[HOPB_NUM_SCRATCH];
  int64_t hist_l2_budget;  // bytes of the count grid that one histogram pass may reduce into
  // count thresholds cached in scratch slot 9: valid for this (grid, n_frames, factor, threshold)
  uint64_t cmin_grid; int64_t cmin_frames; double cmin_factor, cmin_threshold; void* cmin_ptr;
  // tables of hopb_site_map_fused (scratch slots 11, 12): valid for fused_key = (grid, n_frames, factor, thresholds)
  void* fused_cmin; void* fused_dtable; double fused_key[5];
};

struct hopb_grid {
  hopb_handle* h;
  HopbGridTab tab;
  int n[3];
  float* d_table;      // packed up32 tables (tab.table_len floats)
  HopbPair* d_table2;  // packed pair tables (tab.table2_len entries)
  double* d_width[3];  // diff(edges[d]) in float64, n[d] entries
  double* d_mid[3];    // 0.5*(e[:-1]+e[1:]), n[d] entries
  // bin widths by class: np.diff(np.linspace(...)) takes a handful of distinct values per axis.
  // n_cls[d] = 0: more than HOPB_MAX_WIDTH_CLASSES distinct widths on some axis (feature off)
  int n_cls[3];
  uint8_t* d_cls[3];                               // class of every bin, n[d] entries
  double cls_width[3][16];                         // width of every class
  double* d_cls_width;                             // the same on the device, [3][16]
  uint64_t serial;                                 // unique per created grid (cache keys)
};
#define HOPB_MAX_WIDTH_CLASSES 16

static inline int hopb_fail(hopb_handle* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define HOPB_CUDA(h, call)                                                                       \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return hopb_fail((h), HOPB_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call,             \
                       cudaGetErrorString(_e));                                                  \
  } while (0)

#define HOPB_LAUNCH_CHECK(h) HOPB_CUDA(h, cudaGetLastError())

// every kernel launch of the library goes through HOPB_K so that hopb_launch_count() can report how
// many of our kernels ran (bench.py's "gpu_launches")
extern unsigned long long g_hopb_launches;
static inline int hopb_note_launch() { ++g_hopb_launches; return 0; }
#define HOPB_K(kernel) hopb_note_launch(), kernel

#define HOPB_REQUIRE(h, cond, msg)                                                               \
  do {                                                                                           \
    if (!(cond)) return hopb_fail((h), HOPB_E_INVALID, "%s: %s", __func__, (msg));               \
  } while (0)

// Growable per-handle scratch slot.  Growing synchronises the device (cudaFree), so steady-state
// calls with stable sizes never allocate.
int hopb_scratch(hopb_handle* h, int slot, size_t bytes, void** out);

// ---- device-wide primitives implemented in sort.cu --------------------------------------------
// Exclusive prefix sum of n uint32 values (in -> out, may alias); total (device uint32*) optional.
int hopb_exclusive_scan_u32(hopb_handle* h, cudaStream_t s, const uint32_t* in, uint32_t* out, int64_t n,
                            uint32_t* total_dev);
// Stable LSD radix sort of (key, value) pairs on key bits [begin_bit, end_bit).  Buffers ping-pong;
// *in_alt tells whether the result ended in the *_alt buffers.
int hopb_radix_sort_pairs(hopb_handle* h, cudaStream_t s, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals,
                          uint32_t* vals_alt, int64_t n, int begin_bit, int end_bit, int* in_alt);

static __host__ __device__ inline int hopb_bits_for(uint64_t v) {  // number of bits needed to represent values in [0, v]
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return b < 1 ? 1 : b;
}

static inline unsigned hopb_grid_for(int64_t n, int block, int num_sms, int per_sm) {
  static const int scale = getenv("HOPB_GRID_SCALE") ? atoi(getenv("HOPB_GRID_SCALE")) : 4;  // tuning hook (x4 measured best on B200)
  int64_t need = (n + block - 1) / block;
  int64_t cap = (int64_t)num_sms * per_sm * scale;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

// The density of a cell with `c` counts, exactly as the reference computes it: DensityCollector.finish
// (/ n_frames), Grid.make_density (/ dx, / dy, / dz in this order), Grid.convert_density (* factor).
// IEEE double divisions: the results are bit-identical to numpy's.  Monotone non-decreasing in c.
__device__ __forceinline__ double hopb_density_value(uint32_t c, double n_frames, double dx, double dy, double dz,
                                                     double factor, int apply_factor) {
  double v = (double)c;
  v = v / n_frames;
  v = v / dx;
  v = v / dy;
  v = v / dz;
  if (apply_factor) v = v * factor;
  return v;
}

// ---- bricked site map ---------------------------------------------------------------------------
// The site map is stored in 4x4x4 bricks of 64 cells (one 128-byte line of int16 each), brick-major.
// Inside a brick the cells are ordered by 2x2x2 sub-brick first: the index of cell (i,j,k) is
//   brick*64 + (i1 j1 k1 i0 j0 k0)   with i1 = bit 1 of i, i0 = bit 0 of i, ...
// so index >> 3 numbers the 2x2x2 sub-bricks (the 1-bit "fine" block map) and index >> 6 the bricks
// (the 2-bit "coarse" block map).  0xFFFFFFFF marks an outlier; bit 30 is bookkeeping of the slabbed
// histogram.
__device__ __forceinline__ uint32_t hopb_brick_index(int i, int j, int k, int cny, int cnz) {
  const uint32_t cb = (uint32_t)(((i >> 2) * cny + (j >> 2)) * cnz + (k >> 2));
  return (cb << 6) | (uint32_t)(((i & 2) << 4) | ((j & 2) << 3) | ((k & 2) << 2) | ((i & 1) << 2) | ((j & 1) << 1) | (k & 1));
}
// position inside the brick -> (di, dj, dk) in 0..3
__device__ __forceinline__ void hopb_brick_decode(uint32_t c, int& di, int& dj, int& dk) {
  di = (int)(((c >> 4) & 2u) | ((c >> 2) & 1u));
  dj = (int)(((c >> 3) & 2u) | ((c >> 1) & 1u));
  dk = (int)(((c >> 2) & 2u) | (c & 1u));
}
