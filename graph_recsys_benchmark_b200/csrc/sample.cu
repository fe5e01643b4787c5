// K8: BPR training rows assembled on the device.
//
// The reference builds the epoch's table on the host (datasets/movielens.py:920-940: every
// interaction repeated num_negative_samples times, one sampled negative per copy, then a shuffle) and
// adds the entity-aware columns per sample in __getitem__ (:1153-1177).  Here row r of that (unshuffled)
// table is a pure function of (seed, epoch, r): a Philox4x32-10 counter-based generator keyed by the seed
// and counted by (r, epoch, draw) supplies the uniform draws, so a batch is produced from its row ids
// alone, nothing of the [E*k, cols] table is ever materialised, and the CPU mirror
// (oracle/device_sampler.py) reproduces it bit for bit.  Same distributions as the reference:
//   random : negative uniform over all items (np.random.randint over the item id range);
//   unseen : negative uniform over the items the user has no TRAIN interaction with
//            (rd.choices(test_pos[u] + neg_map[u])) - the k-th unseen id by binary search over
//            the user's sorted train items;
//   entity columns: e+ uniform over the item's (user's) feature nodes, e- uniform over the id range of
//            e+'s node type, mask 1; (0, 0, 0) when there are no features.
// The host RNG stream of the reference cannot be reproduced on a GPU; the bit-exact host sampler stays
// the default (datasets/, tests/test_host_logic.py) and this path is opt-in (train_args['device_sampling']).
#include "common.cuh"

namespace peagnn {

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// uniform integer in [0, n) from 64 random bits (multiply-high; bias < n / 2^64)
__device__ __forceinline__ int64_t bounded(uint32_t lo, uint32_t hi, int64_t n) {
  return (int64_t)__umul64hi(((uint64_t)hi << 32) | lo, (uint64_t)n);
}

__global__ void __launch_bounds__(256) bpr_rows_kernel(
    const int64_t* __restrict__ row_ids, int64_t B, const int64_t* __restrict__ u2i, int64_t E, int num_neg,
    uint64_t seed, uint64_t epoch, int strategy, int64_t user_lo, int64_t item_lo, int64_t num_items,
    const int64_t* __restrict__ seen_ptr, const int64_t* __restrict__ seen_items, int cols,
    const int64_t* __restrict__ ifeat_ptr, const int64_t* __restrict__ ifeat_nids,
    const int64_t* __restrict__ ufeat_ptr, const int64_t* __restrict__ ufeat_nids,
    const int64_t* __restrict__ type_starts, int num_types, int64_t* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint64_t r = (uint64_t)row_ids[b];
  const int64_t e = (int64_t)(r / (uint64_t)num_neg);
  const int64_t u = u2i[e], pos = u2i[E + e];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t c0 = (uint32_t)r, c1 = (uint32_t)(r >> 32), c2 = (uint32_t)epoch;
  const Philox4 d0 = philox4x32_10(c0, c1, c2, 0u, k0, k1);

  int64_t neg;
  if (strategy == 0) {
    neg = item_lo + bounded(d0.x, d0.y, num_items);
  } else {
    const int64_t s0 = seen_ptr[u - user_lo], deg = seen_ptr[u - user_lo + 1] - s0;
    const int64_t kth = bounded(d0.x, d0.y, num_items - deg);
    // smallest p with (p == deg) or (seen[p] - item_lo - p > kth): p seen items precede the kth unseen one
    int64_t lo = 0, hi = deg;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (seen_items[s0 + mid] - item_lo - mid > kth) hi = mid; else lo = mid + 1;
    }
    neg = item_lo + kth + lo;
  }
  int64_t* o = out + b * cols;
  o[0] = u; o[1] = pos; o[2] = neg;
  if (cols < 9) return;

  const Philox4 d1 = philox4x32_10(c0, c1, c2, 1u, k0, k1);
  const Philox4 d2 = philox4x32_10(c0, c1, c2, 2u, k0, k1);
  auto entity = [&](const int64_t* ptr, const int64_t* nids, int64_t local, uint32_t a_lo, uint32_t a_hi,
                    uint32_t b_lo, uint32_t b_hi, int64_t* dst) {
    const int64_t f0 = ptr[local], cnt = ptr[local + 1] - f0;
    if (cnt == 0) { dst[0] = 0; dst[1] = 0; dst[2] = 0; return; }
    const int64_t pe = nids[f0 + bounded(a_lo, a_hi, cnt)];
    int t = 0;                                        // node type of pe: last start <= pe
    for (int q = 1; q < num_types; ++q) t = type_starts[q] <= pe ? q : t;
    const int64_t tlo = type_starts[t], thi = type_starts[t + 1];
    dst[0] = pe; dst[1] = tlo + bounded(b_lo, b_hi, thi - tlo); dst[2] = 1;
  };
  entity(ifeat_ptr, ifeat_nids, pos - item_lo, d0.z, d0.w, d1.x, d1.y, o + 3);
  entity(ufeat_ptr, ufeat_nids, u - user_lo, d1.z, d1.w, d2.x, d2.y, o + 6);
}

}  // namespace peagnn

using namespace peagnn;

extern "C" int peagnn_bpr_rows(const int64_t* row_ids, int64_t B, const int64_t* u2i, int64_t E, int32_t num_neg,
                               uint64_t seed, uint64_t epoch, int32_t strategy, int64_t user_lo, int64_t item_lo,
                               int64_t num_items, const int64_t* seen_ptr, const int64_t* seen_items, int32_t cols,
                               const int64_t* ifeat_ptr, const int64_t* ifeat_nids, const int64_t* ufeat_ptr,
                               const int64_t* ufeat_nids, const int64_t* type_starts, int32_t num_types,
                               int64_t* out, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(B >= 0 && E > 0 && num_neg > 0 && num_items > 0 && (strategy == 0 || strategy == 1),
                 "peagnn_bpr_rows: bad sizes / strategy %d", strategy);
  PEAGNN_REQUIRE(cols == 3 || cols == 9, "peagnn_bpr_rows: cols = %d, expected 3 or 9", cols);
  PEAGNN_REQUIRE(row_ids && u2i && out && (strategy == 0 || (seen_ptr && seen_items)),
                 "peagnn_bpr_rows: missing table");
  PEAGNN_REQUIRE(cols == 3 || (ifeat_ptr && ifeat_nids && ufeat_ptr && ufeat_nids && type_starts && num_types > 0 && num_types <= 64),
                 "peagnn_bpr_rows: the entity-aware columns need the feature tables and 1..64 node types");
  if (B == 0) return PEAGNN_OK;
  bpr_rows_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(
      row_ids, B, u2i, E, num_neg, seed, epoch, strategy, user_lo, item_lo, num_items, seen_ptr, seen_items, cols,
      ifeat_ptr, ifeat_nids, ufeat_ptr, ufeat_nids, type_starts, num_types, out);
  return check_launch("peagnn_bpr_rows");
}
