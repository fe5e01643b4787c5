// Graph preparation: COO [2,E] int64 (the reference's edge_index, utils/general_utils.py:280-395)
// -> int32 CSR grouped by one endpoint, stable in COO order; degree scalings of GCN / SAGE.
// Integer work, bit-exact against oracle/graph.py::csr_by_key.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include "common.cuh"

namespace peagnn {

static thread_local char g_err[512] = "";

unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

__global__ void csr_keys_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ val,
                                int64_t E, int32_t N, int drop, int32_t* __restrict__ k32,
                                int32_t* __restrict__ ids) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t k = key[e];
    k32[e] = (drop && k == val[e]) ? N : (int32_t)k;  // dropped edges sort behind every row
    ids[e] = (int32_t)e;
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ val, const int32_t* __restrict__ ids_sorted,
                                int64_t E, int32_t* __restrict__ col, int32_t* __restrict__ eid) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < E; k += stride) {
    const int32_t e = ids_sorted[k];
    eid[k] = e;
    col[k] = (int32_t)val[e];
  }
}

// rowptr[i] = first position whose sorted key is >= i  (i = 0..N).
__global__ void csr_rowptr_kernel(const int32_t* __restrict__ keys_sorted, int64_t E, int32_t N,
                                  int32_t* __restrict__ rowptr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += stride) {
    int64_t lo = 0, hi = E;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys_sorted[mid] < (int32_t)i) lo = mid + 1; else hi = mid;
    }
    rowptr[i] = (int32_t)lo;
  }
}

__global__ void degree_scale_kernel(const int32_t* __restrict__ rowptr, int32_t N, float add,
                                    float power, int clamp1, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float d = (float)(rowptr[i + 1] - rowptr[i]) + add;
  if (clamp1) d = fmaxf(d, 1.f);
  float r;
  if (power == -0.5f) r = d > 0.f ? 1.f / sqrtf(d) : 0.f;   // deg^-1/2, inf -> 0 (GCNConv.norm)
  else if (power == -1.f) r = d > 0.f ? 1.f / d : 0.f;
  else r = powf(d, power);
  out[i] = r;
}

// ---- per-step sub-structure of a CSR: only the edges whose gathered node is marked -------------------------------
__global__ void filter_flags_kernel(const int32_t* __restrict__ col, int64_t E, const uint32_t* __restrict__ active,
                                    int32_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int c = __ldg(col + e);
    flags[e] = (int32_t)((__ldg(active + (c >> 5)) >> (c & 31)) & 1u);
  }
}

__global__ void filter_fill_kernel(const int32_t* __restrict__ col, const int32_t* __restrict__ perm, int64_t E,
                                   const int32_t* __restrict__ flags, const int32_t* __restrict__ pos,
                                   int32_t* __restrict__ col_out, int32_t* __restrict__ perm_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    if (flags[e]) {
      const int32_t k = pos[e];
      col_out[k] = col[e];
      if (perm_out) perm_out[k] = perm ? perm[e] : (int32_t)e;
    }
  }
}

__global__ void filter_rowptr_kernel(const int32_t* __restrict__ rowptr, int32_t nrows, int64_t E,
                                     const int32_t* __restrict__ flags, const int32_t* __restrict__ pos,
                                     int32_t* __restrict__ rowptr_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nrows) return;
  const int64_t e = rowptr[i];
  rowptr_out[i] = e < E ? pos[e] : (E > 0 ? pos[E - 1] + flags[E - 1] : 0);
}

static size_t scan_temp_bytes(int64_t E) {
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr, (int)E);
  return tb;
}

static size_t sort_temp_bytes(int64_t E, int bits) {
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)E, 0, bits);
  return tb;
}

static int key_bits(int32_t N) {
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) <= (int64_t)N) ++bits;  // keys go up to N inclusive
  return bits;
}

}  // namespace peagnn

using namespace peagnn;

extern "C" int peagnn_version(void) { return 100; }
extern "C" const char* peagnn_last_error(void) { return g_err; }
extern "C" unsigned long long peagnn_launch_count(void) { return g_launches; }

extern "C" size_t peagnn_csr_workspace_bytes(int64_t E, int32_t N) {
  if (E <= 0) return 256;
  return 4 * align256((size_t)E * 4) + align256(sort_temp_bytes(E, key_bits(N))) + 256;
}

extern "C" int peagnn_csr_build(const int64_t* key, const int64_t* val, int64_t E, int32_t N,
                                int drop_self_loops, int32_t* rowptr, int32_t* col, int32_t* eid,
                                void* workspace, size_t workspace_bytes, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(N >= 0 && E >= 0 && rowptr, "peagnn_csr_build: bad sizes");
  PEAGNN_REQUIRE(E < ((int64_t)1 << 31) - 1, "peagnn_csr_build: E=%lld exceeds int32 offsets", (long long)E);
  if (E == 0) {
    cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * ((size_t)N + 1), stream);
    return check_launch("peagnn_csr_build(memset)");
  }
  PEAGNN_REQUIRE(key && val && col && eid && workspace, "peagnn_csr_build: null pointer");
  if (workspace_bytes < peagnn_csr_workspace_bytes(E, N)) {
    set_error("peagnn_csr_build: workspace %zu < %zu bytes", workspace_bytes, peagnn_csr_workspace_bytes(E, N));
    return PEAGNN_ERR_WORKSPACE;
  }
  char* ws = static_cast<char*>(workspace);
  ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  const size_t seg = align256((size_t)E * 4);
  int32_t* k_in = reinterpret_cast<int32_t*>(ws);
  int32_t* k_out = reinterpret_cast<int32_t*>(ws + seg);
  int32_t* v_in = reinterpret_cast<int32_t*>(ws + 2 * seg);
  int32_t* v_out = reinterpret_cast<int32_t*>(ws + 3 * seg);
  void* temp = ws + 4 * seg;
  const int bits = key_bits(N);
  size_t temp_bytes = sort_temp_bytes(E, bits);

  const int threads = 256;
  const int blocks = (int)imin64((E + threads - 1) / threads, (int64_t)kNumSMs * 16);
  csr_keys_kernel<<<blocks, threads, 0, stream>>>(key, val, E, N, drop_self_loops, k_in, v_in);
  int rc = check_launch("peagnn_csr_build(keys)");
  if (rc) return rc;
  cudaError_t ce = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k_in, k_out, v_in, v_out, (int)E, 0, bits, stream);
  if (ce != cudaSuccess) {
    set_error("peagnn_csr_build(sort): %s", cudaGetErrorString(ce));
    return PEAGNN_ERR_CUDA;
  }
  csr_fill_kernel<<<blocks, threads, 0, stream>>>(val, v_out, E, col, eid);
  rc = check_launch("peagnn_csr_build(fill)");
  if (rc) return rc;
  const int rblocks = (int)imin64(((int64_t)N + 1 + threads - 1) / threads, (int64_t)kNumSMs * 16);
  csr_rowptr_kernel<<<rblocks, threads, 0, stream>>>(k_out, E, N, rowptr);
  return check_launch("peagnn_csr_build(rowptr)");
}

extern "C" size_t peagnn_csr_filter_workspace_bytes(int64_t E) {
  if (E <= 0) return 256;
  return 2 * align256((size_t)E * 4) + align256(scan_temp_bytes(E)) + 256;
}

extern "C" int peagnn_csr_filter(const peagnn_csr_t* g, const uint32_t* active_cols, const int32_t* perm,
                                 int32_t* rowptr_out, int32_t* col_out, int32_t* perm_out, void* workspace,
                                 size_t workspace_bytes, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(g && g->rowptr && active_cols && rowptr_out && g->row_offset == 0, "peagnn_csr_filter: bad arguments");
  const int64_t E = g->nnz;
  PEAGNN_REQUIRE(E >= 0 && E < ((int64_t)1 << 31) - 1, "peagnn_csr_filter: the view must carry its edge count (nnz)");
  if (E == 0) {
    cudaMemsetAsync(rowptr_out, 0, sizeof(int32_t) * ((size_t)g->nrows + 1), stream);
    return check_launch("peagnn_csr_filter(memset)");
  }
  PEAGNN_REQUIRE(g->col && col_out && workspace, "peagnn_csr_filter: null pointer");
  if (workspace_bytes < peagnn_csr_filter_workspace_bytes(E)) {
    set_error("peagnn_csr_filter: workspace %zu < %zu bytes", workspace_bytes, peagnn_csr_filter_workspace_bytes(E));
    return PEAGNN_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const size_t seg = align256((size_t)E * 4);
  int32_t* flags = reinterpret_cast<int32_t*>(ws);
  int32_t* pos = reinterpret_cast<int32_t*>(ws + seg);
  void* temp = ws + 2 * seg;
  size_t temp_bytes = scan_temp_bytes(E);
  const int threads = 256;
  const int blocks = (int)imin64((E + threads - 1) / threads, (int64_t)kNumSMs * 16);
  filter_flags_kernel<<<blocks, threads, 0, stream>>>(g->col, E, active_cols, flags);
  int rc = check_launch("peagnn_csr_filter(flags)");
  if (rc) return rc;
  cudaError_t ce = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, flags, pos, (int)E, stream);
  if (ce != cudaSuccess) {
    set_error("peagnn_csr_filter(scan): %s", cudaGetErrorString(ce));
    return PEAGNN_ERR_CUDA;
  }
  filter_fill_kernel<<<blocks, threads, 0, stream>>>(g->col, perm, E, flags, pos, col_out, perm_out);
  rc = check_launch("peagnn_csr_filter(fill)");
  if (rc) return rc;
  filter_rowptr_kernel<<<(g->nrows + 1 + threads - 1) / threads, threads, 0, stream>>>(g->rowptr, g->nrows, E, flags, pos, rowptr_out);
  return check_launch("peagnn_csr_filter(rowptr)");
}

extern "C" int peagnn_degree_scale(const int32_t* rowptr, int32_t N, float add, float power,
                                   int clamp_min_one, float* out, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(rowptr && out && N >= 0, "peagnn_degree_scale: bad arguments");
  if (N == 0) return PEAGNN_OK;
  degree_scale_kernel<<<(N + 255) / 256, 256, 0, stream>>>(rowptr, N, add, power, clamp_min_one, out);
  return check_launch("peagnn_degree_scale");
}
