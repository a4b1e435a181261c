// Handle, memory helpers and grid tables of libhopb200.so.
#include "hopb_common.cuh"
#include "hopb_gridtab_host.h"

unsigned long long g_hopb_launches = 0;

extern "C" {

int hopb_version(void) { return 100; }

uint64_t hopb_launch_count(void) { return g_hopb_launches; }

int hopb_create(int device, hopb_handle** out) {
  if (!out) return HOPB_E_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return HOPB_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return HOPB_E_CUDA;
  hopb_handle* h = new hopb_handle();
  memset(h, 0, sizeof(*h));
  h->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return HOPB_E_CUDA; }
  h->num_sms = prop.multiProcessorCount;
  h->hist_l2_budget = (int64_t)64 << 20;  // measured on B200: slabs up to ~65 MB reduce at the L2 rate
  *out = h;
  return HOPB_OK;
}

void hopb_destroy(hopb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (int i = 0; i < HOPB_NUM_SCRATCH; ++i)
    if (h->scratch[i]) cudaFree(h->scratch[i]);
  delete h;
}

const char* hopb_last_error(const hopb_handle* h) { return h ? h->err : "null handle"; }

int hopb_set_hist_l2_budget(hopb_handle* h, int64_t bytes) {
  HOPB_REQUIRE(h, h && bytes >= 1024, "bad argument");
  h->hist_l2_budget = bytes;
  return HOPB_OK;
}

int hopb_dev_alloc(hopb_handle* h, uint64_t bytes, void** out) {
  HOPB_REQUIRE(h, out != nullptr, "null out");
  HOPB_CUDA(h, cudaSetDevice(h->device));
  HOPB_CUDA(h, cudaMalloc(out, bytes ? bytes : 1));
  return HOPB_OK;
}

int hopb_dev_free(hopb_handle* h, void* p) {
  HOPB_CUDA(h, cudaFree(p));
  return HOPB_OK;
}

int hopb_memset(hopb_handle* h, void* stream, void* dev, int value, uint64_t bytes) {
  HOPB_CUDA(h, cudaMemsetAsync(dev, value, bytes, (cudaStream_t)stream));
  return HOPB_OK;
}

int hopb_h2d(hopb_handle* h, void* stream, void* dev, const void* host, uint64_t bytes) {
  HOPB_CUDA(h, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return HOPB_OK;
}

int hopb_d2h(hopb_handle* h, void* stream, void* host, const void* dev, uint64_t bytes) {
  HOPB_CUDA(h, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return HOPB_OK;
}

int hopb_stream_sync(hopb_handle* h, void* stream) {
  HOPB_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  return HOPB_OK;
}

int hopb_grid_create(hopb_handle* h, const double* ex, int nx, const double* ey, int ny, const double* ez,
                     int nz, const int* decimals, hopb_grid** out) {
  HOPB_REQUIRE(h, h && out && ex && ey && ez, "null argument");
  *out = nullptr;
  const double* e[3] = {ex, ey, ez};
  const int n[3] = {nx, ny, nz};
  HOPB_REQUIRE(h, (int64_t)nx * ny * nz < ((int64_t)1 << 31), "grid has more than 2^31 cells");
  HopbGridTab tab;
  std::vector<float> table;
  std::vector<HopbPair> table2;
  if (hopb_build_gridtab(e, n, decimals, &tab, &table, &table2) != 0)
    return hopb_fail(h, HOPB_E_INVALID, "hopb_grid_create: edges must be strictly increasing, 1 <= n < %d", HOPB_MAX_EDGES);
  HOPB_CUDA(h, cudaSetDevice(h->device));
  hopb_grid* g = new hopb_grid();
  memset(g, 0, sizeof(*g));
  static uint64_t next_serial = 1;
  g->serial = __atomic_fetch_add(&next_serial, 1, __ATOMIC_RELAXED);
  g->h = h;
  g->tab = tab;
  for (int d = 0; d < 3; ++d) g->n[d] = n[d];
  cudaError_t err = cudaMalloc(&g->d_table, table.size() * sizeof(float));
  if (err == cudaSuccess)
    err = cudaMemcpy(g->d_table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMalloc(&g->d_table2, table2.size() * sizeof(HopbPair));
  if (err == cudaSuccess)
    err = cudaMemcpy(g->d_table2, table2.data(), table2.size() * sizeof(HopbPair), cudaMemcpyHostToDevice);
  for (int d = 0; d < 3 && err == cudaSuccess; ++d) {
    std::vector<double> w(n[d]), m(n[d]);
    for (int k = 0; k < n[d]; ++k) {
      w[k] = e[d][k + 1] - e[d][k];          // numpy.diff(edges)
      m[k] = 0.5 * (e[d][k] + e[d][k + 1]);  // Grid._update midpoints (hop/sitemap.py:163)
    }
    err = cudaMalloc(&g->d_width[d], n[d] * sizeof(double));
    if (err == cudaSuccess) err = cudaMalloc(&g->d_mid[d], n[d] * sizeof(double));
    if (err == cudaSuccess) err = cudaMemcpy(g->d_width[d], w.data(), n[d] * sizeof(double), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(g->d_mid[d], m.data(), n[d] * sizeof(double), cudaMemcpyHostToDevice);
    // width classes (exact comparison of the doubles)
    std::vector<uint8_t> cls(n[d]);
    int nc = 0;
    for (int k = 0; k < n[d] && nc >= 0; ++k) {
      int c = 0;
      while (c < nc && g->cls_width[d][c] != w[k]) ++c;
      if (c == nc) {
        if (nc == HOPB_MAX_WIDTH_CLASSES) { nc = -1; break; }
        g->cls_width[d][nc++] = w[k];
      }
      cls[k] = (uint8_t)c;
    }
    g->n_cls[d] = nc > 0 ? nc : 0;
    if (err == cudaSuccess) err = cudaMalloc(&g->d_cls[d], n[d]);
    if (err == cudaSuccess) err = cudaMemcpy(g->d_cls[d], cls.data(), n[d], cudaMemcpyHostToDevice);
  }
  if (err == cudaSuccess) err = cudaMalloc(&g->d_cls_width, sizeof(g->cls_width));
  if (err == cudaSuccess) err = cudaMemcpy(g->d_cls_width, g->cls_width, sizeof(g->cls_width), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    hopb_grid_destroy(g);
    return hopb_fail(h, HOPB_E_CUDA, "hopb_grid_create: %s", cudaGetErrorString(err));
  }
  *out = g;
  return HOPB_OK;
}

int hopb_grid_set_periodic_box(hopb_grid* g, const double* origin, const double* length) {
  if (!g) return HOPB_E_INVALID;
  if (!origin || !length) {   // off (the default): coordinates are binned as they are
    g->tab.wrap = 0;
    return HOPB_OK;
  }
  for (int d = 0; d < 3; ++d)
    if (!(length[d] > 0.0) || !std::isfinite(length[d]) || !std::isfinite(origin[d]))
      return hopb_fail(g->h, HOPB_E_INVALID, "hopb_grid_set_periodic_box: box lengths must be positive and finite");
  for (int d = 0; d < 3; ++d) {
    g->tab.box_o[d] = (float)origin[d];
    g->tab.box_l[d] = (float)length[d];
    g->tab.box_inv[d] = 1.0f / (float)length[d];
  }
  g->tab.wrap = 1;
  return HOPB_OK;
}

int hopb_grid_width_classes(const hopb_grid* g) {
  if (!g || g->n_cls[0] == 0 || g->n_cls[1] == 0 || g->n_cls[2] == 0) return 0;
  return g->n_cls[0] * g->n_cls[1] * g->n_cls[2];
}

void hopb_grid_destroy(hopb_grid* g) {
  if (!g) return;
  if (g->d_table) cudaFree(g->d_table);
  if (g->d_table2) cudaFree(g->d_table2);
  for (int d = 0; d < 3; ++d) {
    if (g->d_width[d]) cudaFree(g->d_width[d]);
    if (g->d_mid[d]) cudaFree(g->d_mid[d]);
    if (g->d_cls[d]) cudaFree(g->d_cls[d]);
  }
  if (g->d_cls_width) cudaFree(g->d_cls_width);
  delete g;
}

}  // extern "C"

int hopb_scratch(hopb_handle* h, int slot, size_t bytes, void** out) {
  if (slot < 0 || slot >= HOPB_NUM_SCRATCH) return hopb_fail(h, HOPB_E_INVALID, "bad scratch slot");
  if (h->scratch_bytes[slot] < bytes) {
    if (h->scratch[slot]) HOPB_CUDA(h, cudaFree(h->scratch[slot]));
    h->scratch[slot] = nullptr;
    h->scratch_bytes[slot] = 0;
    size_t want = bytes + bytes / 4 + 256;
    HOPB_CUDA(h, cudaMalloc(&h->scratch[slot], want));
    h->scratch_bytes[slot] = want;
  }
  *out = h->scratch[slot];
  return HOPB_OK;
}
