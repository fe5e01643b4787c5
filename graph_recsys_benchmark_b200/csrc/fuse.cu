// K5: fusion across metapath channels (reference models/base.py:191-206).
//   'att' : w[n,p] = softmax_p( <Z[n,p,:], att[p,:]> ),  out[n,:] = sum_p w[n,p] Z[n,p,:]
//   'mean': out[n,:] = mean_p Z[n,p,:]
// skip_path zeroes one channel before fusing (metapath ablation, base.py:194-195): the zeroed
// channel still takes part in the softmax with logit 0, exactly as upstream.
// One G-lane group (G = D/4) per node; each lane keeps one float4 of the output row.
#include "common.cuh"

namespace peagnn {

template <int G>
__device__ __forceinline__ float gsum(float v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
  return v;
}

template <int G>
__device__ __forceinline__ unsigned gmask_of() {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << ((lane / G) * G);
  }
}

template <int G>
__global__ void __launch_bounds__(256) fuse_fwd_kernel(const float* __restrict__ Z, int64_t ldz, int64_t n_rows,
                                                       int P, int D, const float* __restrict__ att, int mode,
                                                       int skip, float* __restrict__ out, int64_t ldo) {
  const int gl = threadIdx.x % G;
  const int64_t n = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
  if (n >= n_rows) return;
  const unsigned gmask = gmask_of<G>();
  const bool on = 4 * gl < D;
  const float* zr = Z + n * ldz + 4 * gl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (mode == 1) {
    for (int p = 0; p < P; ++p) {
      if (p == skip || !on) continue;
      const float4 z = ldg4(zr + (int64_t)p * D);
      acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
    }
    const float inv = 1.f / (float)P;
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  } else {
    // pass 1: logits -> max ; pass 2: exp / sum and the weighted rows (rows re-read from L1)
    float m = -INFINITY;
    for (int p = 0; p < P; ++p) {
      float s = 0.f;
      if (p != skip && on) {
        const float4 z = ldg4(zr + (int64_t)p * D);
        const float4 a = ldg4(att + (int64_t)p * D + 4 * gl);
        s = z.x * a.x + z.y * a.y + z.z * a.z + z.w * a.w;
      }
      s = gsum<G>(s, gmask);
      m = fmaxf(m, s);
    }
    float l = 0.f;
    for (int p = 0; p < P; ++p) {
      float s = 0.f;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p != skip && on) {
        z = ldg4(zr + (int64_t)p * D);
        const float4 a = ldg4(att + (int64_t)p * D + 4 * gl);
        s = z.x * a.x + z.y * a.y + z.z * a.z + z.w * a.w;
      }
      s = gsum<G>(s, gmask);
      const float e = expf(s - m);
      l += e;
      acc.x = fmaf(e, z.x, acc.x); acc.y = fmaf(e, z.y, acc.y);
      acc.z = fmaf(e, z.z, acc.z); acc.w = fmaf(e, z.w, acc.w);
    }
    const float inv = 1.f / l;
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  }
  if (on) st4(out + n * ldo + 4 * gl, acc);
}

// Backward.  Per node: w_p, dw_p = <dout, z_p>, t = sum_q w_q dw_q, ds_p = w_p (dw_p - t),
// dz_p = w_p dout + ds_p att_p,  d att_p += ds_p z_p.
// A CTA owns a slab of nodes; each group walks its nodes for a fixed p so that d att_p sits in
// registers, then the CTA folds its groups in order -> partial[cta][P*D] (deterministic).
constexpr int kFuseNodesPerGroup = 8;

template <int G>
__global__ void __launch_bounds__(256) fuse_bwd_kernel(const float* __restrict__ Z, int64_t ldz, int64_t n_rows,
                                                       int P, int D, const float* __restrict__ att, int mode,
                                                       const float* __restrict__ dout, int64_t ldo,
                                                       float* __restrict__ dZ, int64_t lddz,
                                                       float* __restrict__ partial) {
  constexpr int GPB = 256 / G;
  __shared__ __align__(16) float red[256 * 4];
  const int gl = threadIdx.x % G;
  const int grp = threadIdx.x / G;
  const unsigned gmask = gmask_of<G>();
  const bool on = 4 * gl < D;
  const int64_t slab0 = (int64_t)blockIdx.x * GPB * kFuseNodesPerGroup;

  float m[kFuseNodesPerGroup], linv[kFuseNodesPerGroup], t[kFuseNodesPerGroup];
  float4 go[kFuseNodesPerGroup];
#pragma unroll
  for (int q = 0; q < kFuseNodesPerGroup; ++q) {
    const int64_t n = slab0 + (int64_t)q * GPB + grp;
    m[q] = 0.f; linv[q] = 0.f; t[q] = 0.f;
    go[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < n_rows) {
      if (on) go[q] = ldg4(dout + n * ldo + 4 * gl);
      if (mode == 0) {
        const float* zr = Z + n * ldz + 4 * gl;
        float mm = -INFINITY;
        for (int p = 0; p < P; ++p) {
          float s = 0.f;
          if (on) {
            const float4 z = ldg4(zr + (int64_t)p * D);
            const float4 a = ldg4(att + (int64_t)p * D + 4 * gl);
            s = z.x * a.x + z.y * a.y + z.z * a.z + z.w * a.w;
          }
          s = gsum<G>(s, gmask);
          mm = fmaxf(mm, s);
        }
        float l = 0.f, tt = 0.f;
        for (int p = 0; p < P; ++p) {
          float s = 0.f, dw = 0.f;
          if (on) {
            const float4 z = ldg4(zr + (int64_t)p * D);
            const float4 a = ldg4(att + (int64_t)p * D + 4 * gl);
            s = z.x * a.x + z.y * a.y + z.z * a.z + z.w * a.w;
            dw = z.x * go[q].x + z.y * go[q].y + z.z * go[q].z + z.w * go[q].w;
          }
          s = gsum<G>(s, gmask);
          dw = gsum<G>(dw, gmask);
          const float e = expf(s - mm);
          l += e;
          tt = fmaf(e, dw, tt);
        }
        m[q] = mm; linv[q] = 1.f / l; t[q] = tt * linv[q];
      }
    }
  }

  const float invP = 1.f / (float)P;
  for (int p = 0; p < P; ++p) {
    float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode == 0 && on) a = ldg4(att + (int64_t)p * D + 4 * gl);
#pragma unroll
    for (int q = 0; q < kFuseNodesPerGroup; ++q) {
      const int64_t n = slab0 + (int64_t)q * GPB + grp;
      if (n >= n_rows) continue;   // uniform across the group
      float4 dz;
      if (mode == 1) {
        dz = make_float4(go[q].x * invP, go[q].y * invP, go[q].z * invP, go[q].w * invP);
      } else {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float s = 0.f, dw = 0.f;
        if (on) {
          z = ldg4(Z + n * ldz + (int64_t)p * D + 4 * gl);
          s = z.x * a.x + z.y * a.y + z.z * a.z + z.w * a.w;
          dw = z.x * go[q].x + z.y * go[q].y + z.z * go[q].z + z.w * go[q].w;
        }
        s = gsum<G>(s, gmask);
        dw = gsum<G>(dw, gmask);
        const float w = expf(s - m[q]) * linv[q];
        const float ds = w * (dw - t[q]);
        dz.x = fmaf(w, go[q].x, ds * a.x); dz.y = fmaf(w, go[q].y, ds * a.y);
        dz.z = fmaf(w, go[q].z, ds * a.z); dz.w = fmaf(w, go[q].w, ds * a.w);
        da.x = fmaf(ds, z.x, da.x); da.y = fmaf(ds, z.y, da.y);
        da.z = fmaf(ds, z.z, da.z); da.w = fmaf(ds, z.w, da.w);
      }
      if (on) st4(dZ + n * lddz + (int64_t)p * D + 4 * gl, dz);
    }
    if (mode == 0) {
      __syncthreads();
      st4(red + 4 * threadIdx.x, da);   // [grp][gl][4]
      __syncthreads();
      if ((int)threadIdx.x < 4 * G) {    // one thread per output float of this path
        const int lane4 = threadIdx.x;   // = gl*4 + j
        float s = 0.f;
        for (int q = 0; q < GPB; ++q) s += red[q * 4 * G + lane4];
        if (lane4 < D) partial[((size_t)blockIdx.x * P + p) * D + lane4] = s;
      }
    }
  }
}

__global__ void fuse_datt_kernel(const float* __restrict__ partial, int n_parts, int len, float* __restrict__ d_att) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= len) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += partial[(size_t)p * len + idx];
  d_att[idx] = s;
}

static inline int fuse_group(int D) {
  int g = 1;
  while (g < D / 4) g <<= 1;
  return g;
}

}  // namespace peagnn

using namespace peagnn;

static int fuse_check(const float* Z, int64_t ldz, int32_t P, int32_t D, const char* what) {
  PEAGNN_REQUIRE(Z && P > 0 && P <= 32 && D > 0 && D % 4 == 0 && D <= 128 && ldz % 4 == 0 && ldz >= (int64_t)P * D && aligned16(Z),
                 "%s: need 0 < P <= 32, D a multiple of 4 <= 128, aligned rows", what);
  return PEAGNN_OK;
}

extern "C" int peagnn_fuse_forward(const float* Z, int64_t ldz, int64_t n, int32_t P, int32_t D,
                                   const float* att, int mode, int skip_path, float* out, int64_t ldo,
                                   peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = fuse_check(Z, ldz, P, D, "peagnn_fuse_forward");
  if (rc) return rc;
  PEAGNN_REQUIRE(out && aligned16(out) && ldo % 4 == 0 && (mode == 1 || (mode == 0 && att && aligned16(att))),
                 "peagnn_fuse_forward: bad output / att / mode");
  if (n == 0) return PEAGNN_OK;
  const int g = fuse_group(D);
  const unsigned blocks = (unsigned)((n * g + 255) / 256);
#define CASE(G_) case G_: fuse_fwd_kernel<G_><<<blocks, 256, 0, stream>>>(Z, ldz, n, P, D, att, mode, skip_path, out, ldo); break
  switch (g) { CASE(1); CASE(2); CASE(4); CASE(8); CASE(16); default: fuse_fwd_kernel<32><<<blocks, 256, 0, stream>>>(Z, ldz, n, P, D, att, mode, skip_path, out, ldo); }
#undef CASE
  return check_launch("peagnn_fuse_forward");
}

extern "C" size_t peagnn_fuse_workspace_floats(int64_t n, int32_t P, int32_t D) {
  const int g = fuse_group(D);
  const int64_t per_cta = (256 / g) * kFuseNodesPerGroup;
  return (size_t)((n + per_cta - 1) / per_cta) * P * D + 64;
}

extern "C" int peagnn_fuse_backward(const float* Z, int64_t ldz, int64_t n, int32_t P, int32_t D,
                                    const float* att, int mode, const float* dout, int64_t ldo, float* dZ,
                                    int64_t lddz, float* d_att, float* workspace, size_t workspace_floats,
                                    peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = fuse_check(Z, ldz, P, D, "peagnn_fuse_backward");
  if (rc) return rc;
  PEAGNN_REQUIRE(dout && dZ && aligned16(dout) && aligned16(dZ) && ldo % 4 == 0 && lddz % 4 == 0 && lddz >= (int64_t)P * D,
                 "peagnn_fuse_backward: bad pointers");
  PEAGNN_REQUIRE(mode == 1 || (mode == 0 && att && d_att && workspace), "peagnn_fuse_backward: att mode needs att, d_att, workspace");
  if (n == 0) {
    if (mode == 0) cudaMemsetAsync(d_att, 0, sizeof(float) * P * D, stream);
    return check_launch("peagnn_fuse_backward(memset)");
  }
  if (mode == 0 && workspace_floats < peagnn_fuse_workspace_floats(n, P, D)) {
    set_error("peagnn_fuse_backward: workspace %zu < %zu floats", workspace_floats, peagnn_fuse_workspace_floats(n, P, D));
    return PEAGNN_ERR_WORKSPACE;
  }
  const int g = fuse_group(D);
  const int64_t per_cta = (256 / g) * kFuseNodesPerGroup;
  const unsigned blocks = (unsigned)((n + per_cta - 1) / per_cta);
#define CASE(G_) case G_: fuse_bwd_kernel<G_><<<blocks, 256, 0, stream>>>(Z, ldz, n, P, D, att, mode, dout, ldo, dZ, lddz, workspace); break
  switch (g) { CASE(1); CASE(2); CASE(4); CASE(8); CASE(16); default: fuse_bwd_kernel<32><<<blocks, 256, 0, stream>>>(Z, ldz, n, P, D, att, mode, dout, ldo, dZ, lddz, workspace); }
#undef CASE
  rc = check_launch("peagnn_fuse_backward(stage1)");
  if (rc || mode == 1) return rc;
  fuse_datt_kernel<<<(P * D + 255) / 256, 256, 0, stream>>>(workspace, (int)blocks, P * D, d_att);
  return check_launch("peagnn_fuse_backward(stage2)");
}
