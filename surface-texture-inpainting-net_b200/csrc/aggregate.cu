// Segmented aggregation over CSR rows and the fused EdgeConv message stage.
// One warp owns one (CSR row, 128-channel chunk) pair -- 32 lanes x float4 -- so wide rows at the coarse levels (few
// vertices, up to 2048 channels) still spread over the whole machine; a row's entries are accumulated sequentially in
// CSR order (= original edge order), so results are deterministic and follow the summation order of torch_scatter's
// CPU kernels.  No atomics anywhere.  (Scalar fallback for unaligned / odd widths: one warp per row.)
#include "common.cuh"

namespace stinet {

constexpr int kWarpsPerCta = 8;
constexpr int kAggThreads = kWarpsPerCta * 32;

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
// true division (not reciprocal-multiply): the reference divides sums by counts, keep the same rounding
__device__ __forceinline__ float4 f4_div(float4 a, float d) { return make_float4(a.x / d, a.y / d, a.z / d, a.w / d); }
__device__ __forceinline__ float relu(float v) { return v > 0.f ? v : 0.f; }

// ---------------------------------------------------------------------------------------------------------------
// generic aggregate: out[i,:] = reduce_k x[col[k],:]

template <int REDUCE, bool VEC>
__global__ void __launch_bounds__(kAggThreads)
aggregate_fwd_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ col, const int32_t* __restrict__ eid, int64_t n_rows,
                     int32_t n_items, int channels, float* __restrict__ out, int64_t ldo, int32_t* __restrict__ arg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = VEC ? (channels >> 2) : 0;
  const int nchunk = VEC ? ((c4n + 31) >> 5) : 1;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t i = it / nchunk;
    const int chunk = (int)(it - i * nchunk);
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float den = (REDUCE == STINET_REDUCE_MEAN) ? (float)max(end - beg, 1) : 1.f;
    if (VEC) {
      for (int c4 = chunk * 32 + lane; c4 < c4n; c4 += c4n) {   // exactly one float4 column per lane
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = beg;
        for (; k + 4 <= end; k += 4) {
          int j0 = col[k], j1 = col[k + 1], j2 = col[k + 2], j3 = col[k + 3];
          float4 v0 = reinterpret_cast<const float4*>(x + j0 * ldx)[c4];
          float4 v1 = reinterpret_cast<const float4*>(x + j1 * ldx)[c4];
          float4 v2 = reinterpret_cast<const float4*>(x + j2 * ldx)[c4];
          float4 v3 = reinterpret_cast<const float4*>(x + j3 * ldx)[c4];
          acc = f4_add(f4_add(f4_add(f4_add(acc, v0), v1), v2), v3);
        }
        for (; k < end; ++k) acc = f4_add(acc, reinterpret_cast<const float4*>(x + (int64_t)col[k] * ldx)[c4]);
        reinterpret_cast<float4*>(out + i * ldo)[c4] = f4_div(acc, den);
      }
    } else {
      for (int c = lane; c < channels; c += 32) {
        if (REDUCE == STINET_REDUCE_MAX) {
          float best = -3.402823466e+38f;
          int32_t best_e = n_items;
          for (int k = beg; k < end; ++k) {
            float v = x[(int64_t)col[k] * ldx + c];
            if (v > best) {
              best = v;
              best_e = eid[k];
            }
          }
          out[i * ldo + c] = (best_e == n_items) ? 0.f : best;
          arg[i * (int64_t)channels + c] = best_e;
        } else {
          float acc = 0.f;
          for (int k = beg; k < end; ++k) acc += x[(int64_t)col[k] * ldx + c];
          out[i * ldo + c] = acc / den;
        }
      }
    }
  }
}

// dx[j,:] = sum_{out-edges j->i} w * g[i,:]
template <int REDUCE, bool VEC>
__global__ void __launch_bounds__(kAggThreads)
aggregate_bwd_kernel(const float* __restrict__ g, int64_t ldg, const int32_t* __restrict__ rowptr_s,
                     const int32_t* __restrict__ col_s, const int32_t* __restrict__ eid_s,
                     const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ arg, int64_t n_rows,
                     int channels, float* __restrict__ dx, int64_t lddx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = VEC ? (channels >> 2) : 0;
  const int nchunk = VEC ? ((c4n + 31) >> 5) : 1;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t j = it / nchunk;
    const int chunk = (int)(it - j * nchunk);
    const int beg = rowptr_s[j], end = rowptr_s[j + 1];
    if (VEC) {
      for (int c4 = chunk * 32 + lane; c4 < c4n; c4 += c4n) {   // exactly one float4 column per lane
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = beg; k < end; ++k) {
          const int i = col_s[k];
          float den = 1.f;
          if (REDUCE == STINET_REDUCE_MEAN) den = (float)max(rowptr_t[i + 1] - rowptr_t[i], 1);
          float4 v = reinterpret_cast<const float4*>(g + (int64_t)i * ldg)[c4];
          acc = f4_add(acc, f4_div(v, den));
        }
        reinterpret_cast<float4*>(dx + j * lddx)[c4] = acc;
      }
    } else {
      for (int c = lane; c < channels; c += 32) {
        float acc = 0.f;
        for (int k = beg; k < end; ++k) {
          const int i = col_s[k];
          float v = g[(int64_t)i * ldg + c];
          if (REDUCE == STINET_REDUCE_MEAN) v /= (float)max(rowptr_t[i + 1] - rowptr_t[i], 1);
          if (REDUCE == STINET_REDUCE_MAX) v = (arg[(int64_t)i * channels + c] == eid_s[k]) ? v : 0.f;
          acc += v;
        }
        dx[j * lddx + c] = acc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused EdgeConv message stage

template <bool VEC>
__global__ void __launch_bounds__(kAggThreads)
edge_message_fwd_kernel(const float* __restrict__ P, int64_t ldp, const float* __restrict__ Q, int64_t ldq,
                        const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                        int hidden, float* __restrict__ hid, int64_t ldh) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = VEC ? (hidden >> 2) : 0;
  const int nchunk = VEC ? ((c4n + 31) >> 5) : 1;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t i = it / nchunk;
    const int chunk = (int)(it - i * nchunk);
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float den = (float)max(end - beg, 1);
    if (VEC) {
      for (int c4 = chunk * 32 + lane; c4 < c4n; c4 += c4n) {   // exactly one float4 column per lane
        const float4 p = reinterpret_cast<const float4*>(P + i * ldp)[c4];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = beg;
        for (; k + 4 <= end; k += 4) {
          int j0 = col[k], j1 = col[k + 1], j2 = col[k + 2], j3 = col[k + 3];
          float4 q0 = reinterpret_cast<const float4*>(Q + j0 * ldq)[c4];
          float4 q1 = reinterpret_cast<const float4*>(Q + j1 * ldq)[c4];
          float4 q2 = reinterpret_cast<const float4*>(Q + j2 * ldq)[c4];
          float4 q3 = reinterpret_cast<const float4*>(Q + j3 * ldq)[c4];
          acc.x += relu(p.x + q0.x); acc.y += relu(p.y + q0.y); acc.z += relu(p.z + q0.z); acc.w += relu(p.w + q0.w);
          acc.x += relu(p.x + q1.x); acc.y += relu(p.y + q1.y); acc.z += relu(p.z + q1.z); acc.w += relu(p.w + q1.w);
          acc.x += relu(p.x + q2.x); acc.y += relu(p.y + q2.y); acc.z += relu(p.z + q2.z); acc.w += relu(p.w + q2.w);
          acc.x += relu(p.x + q3.x); acc.y += relu(p.y + q3.y); acc.z += relu(p.z + q3.z); acc.w += relu(p.w + q3.w);
        }
        for (; k < end; ++k) {
          float4 q = reinterpret_cast<const float4*>(Q + (int64_t)col[k] * ldq)[c4];
          acc.x += relu(p.x + q.x); acc.y += relu(p.y + q.y); acc.z += relu(p.z + q.z); acc.w += relu(p.w + q.w);
        }
        reinterpret_cast<float4*>(hid + i * ldh)[c4] = f4_div(acc, den);
      }
    } else {
      for (int c = lane; c < hidden; c += 32) {
        const float p = P[i * ldp + c];
        float acc = 0.f;
        for (int k = beg; k < end; ++k) acc += relu(p + Q[(int64_t)col[k] * ldq + c]);
        hid[i * ldh + c] = acc / den;
      }
    }
  }
}

// dP[i,:] = (1/deg_i) * dhid[i,:] * #{j->i : P_i+Q_j > 0}   (per channel)
template <bool VEC>
__global__ void __launch_bounds__(kAggThreads)
edge_message_bwd_target_kernel(const float* __restrict__ P, int64_t ldp, const float* __restrict__ Q, int64_t ldq,
                               const float* __restrict__ dhid, int64_t ldd, const int32_t* __restrict__ rowptr,
                               const int32_t* __restrict__ col, int64_t n_rows, int hidden, float* __restrict__ dP,
                               int64_t lddp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = VEC ? (hidden >> 2) : 0;
  const int nchunk = VEC ? ((c4n + 31) >> 5) : 1;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t i = it / nchunk;
    const int chunk = (int)(it - i * nchunk);
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float den = (float)max(end - beg, 1);
    if (VEC) {
      for (int c4 = chunk * 32 + lane; c4 < c4n; c4 += c4n) {   // exactly one float4 column per lane
        const float4 p = reinterpret_cast<const float4*>(P + i * ldp)[c4];
        const float4 d = f4_div(reinterpret_cast<const float4*>(dhid + i * ldd)[c4], den);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = beg; k < end; ++k) {
          float4 q = reinterpret_cast<const float4*>(Q + (int64_t)col[k] * ldq)[c4];
          acc.x += (p.x + q.x > 0.f) ? d.x : 0.f;
          acc.y += (p.y + q.y > 0.f) ? d.y : 0.f;
          acc.z += (p.z + q.z > 0.f) ? d.z : 0.f;
          acc.w += (p.w + q.w > 0.f) ? d.w : 0.f;
        }
        reinterpret_cast<float4*>(dP + i * lddp)[c4] = acc;
      }
    } else {
      for (int c = lane; c < hidden; c += 32) {
        const float p = P[i * ldp + c];
        const float d = dhid[i * ldd + c] / den;
        float acc = 0.f;
        for (int k = beg; k < end; ++k) acc += (p + Q[(int64_t)col[k] * ldq + c] > 0.f) ? d : 0.f;
        dP[i * lddp + c] = acc;
      }
    }
  }
}

// dQ[j,:] = sum_{j->i} (1/deg_i) * dhid[i,:] * [P_i+Q_j > 0]
template <bool VEC>
__global__ void __launch_bounds__(kAggThreads)
edge_message_bwd_source_kernel(const float* __restrict__ P, int64_t ldp, const float* __restrict__ Q, int64_t ldq,
                               const float* __restrict__ dhid, int64_t ldd, const int32_t* __restrict__ rowptr_t,
                               const int32_t* __restrict__ rowptr_s, const int32_t* __restrict__ col_s,
                               int64_t n_rows, int hidden, float* __restrict__ dQ, int64_t lddq) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = VEC ? (hidden >> 2) : 0;
  const int nchunk = VEC ? ((c4n + 31) >> 5) : 1;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t j = it / nchunk;
    const int chunk = (int)(it - j * nchunk);
    const int beg = rowptr_s[j], end = rowptr_s[j + 1];
    if (VEC) {
      for (int c4 = chunk * 32 + lane; c4 < c4n; c4 += c4n) {   // exactly one float4 column per lane
        const float4 q = reinterpret_cast<const float4*>(Q + j * ldq)[c4];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = beg; k < end; ++k) {
          const int i = col_s[k];
          const float den = (float)max(rowptr_t[i + 1] - rowptr_t[i], 1);
          const float4 p = reinterpret_cast<const float4*>(P + (int64_t)i * ldp)[c4];
          const float4 d = f4_div(reinterpret_cast<const float4*>(dhid + (int64_t)i * ldd)[c4], den);
          acc.x += (p.x + q.x > 0.f) ? d.x : 0.f;
          acc.y += (p.y + q.y > 0.f) ? d.y : 0.f;
          acc.z += (p.z + q.z > 0.f) ? d.z : 0.f;
          acc.w += (p.w + q.w > 0.f) ? d.w : 0.f;
        }
        reinterpret_cast<float4*>(dQ + j * lddq)[c4] = acc;
      }
    } else {
      for (int c = lane; c < hidden; c += 32) {
        const float q = Q[j * ldq + c];
        float acc = 0.f;
        for (int k = beg; k < end; ++k) {
          const int i = col_s[k];
          const float den = (float)max(rowptr_t[i + 1] - rowptr_t[i], 1);
          acc += (P[(int64_t)i * ldp + c] + q > 0.f) ? dhid[(int64_t)i * ldd + c] / den : 0.f;
        }
        dQ[j * lddq + c] = acc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused EdgeConv message stage with the ReLU decisions saved for backward.
// Forward stores, per edge POSITION k of the by-target CSR and float4 column c4, one byte whose low nibble holds the
// four decision bits [P_i + Q_j > 0] of channels 4*c4 .. 4*c4+3 (mask[k * hidden/4 + c4]): every lane writes its own
// byte -- no cross-lane traffic (warp ballots made the forward 1.5x slower: VOTE issues at a fraction of the FP32
// rate) -- a warp writes 32 contiguous bytes per edge and a row's masks are contiguous, so they are written and
// re-read as a stream.  The backward kernels then need neither P nor Q: dP is a row-local count (no gather at all),
// dQ gathers one row (dhid) per out-edge instead of two and finds its mask through tpos_s (position of each
// by-source entry in the by-target order).  Same products, same summation order as the recomputing kernels:
// bit-identical results.

__device__ __forceinline__ unsigned nibble(float sx, float sy, float sz, float sw) {
  return (sx > 0.f ? 1u : 0u) | (sy > 0.f ? 2u : 0u) | (sz > 0.f ? 4u : 0u) | (sw > 0.f ? 8u : 0u);
}

__global__ void __launch_bounds__(kAggThreads)
edge_message_fwd_mask_kernel(const float* __restrict__ P, int64_t ldp, const float* __restrict__ Q, int64_t ldq,
                             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                             int64_t n_rows, int hidden, float* __restrict__ hid, int64_t ldh,
                             uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = hidden >> 2;
  const int nchunk = (c4n + 31) >> 5;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t i = it / nchunk;
    const int chunk = (int)(it - i * nchunk);
    const int c4 = chunk * 32 + lane;
    if (c4 >= c4n) continue;
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float den = (float)max(end - beg, 1);
    const float4 p = reinterpret_cast<const float4*>(P + i * ldp)[c4];
    uint8_t* __restrict__ mrow = mask + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = beg;
    for (; k + 4 <= end; k += 4) {
      int j[4];
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) j[u] = col[k + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = reinterpret_cast<const float4*>(Q + (int64_t)j[u] * ldq)[c4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float sx = p.x + q[u].x, sy = p.y + q[u].y, sz = p.z + q[u].z, sw = p.w + q[u].w;
        acc.x += relu(sx); acc.y += relu(sy); acc.z += relu(sz); acc.w += relu(sw);
        mrow[(int64_t)(k + u) * c4n] = (uint8_t)nibble(sx, sy, sz, sw);
      }
    }
    for (; k < end; ++k) {
      const float4 q = reinterpret_cast<const float4*>(Q + (int64_t)col[k] * ldq)[c4];
      const float sx = p.x + q.x, sy = p.y + q.y, sz = p.z + q.z, sw = p.w + q.w;
      acc.x += relu(sx); acc.y += relu(sy); acc.z += relu(sz); acc.w += relu(sw);
      mrow[(int64_t)k * c4n] = (uint8_t)nibble(sx, sy, sz, sw);
    }
    reinterpret_cast<float4*>(hid + i * ldh)[c4] = f4_div(acc, den);
  }
}

// dP[i,:] = (1/deg_i) * dhid[i,:] * #{in-edges of i whose decision bit is set}: reads the row's own dhid and one mask
// byte per in-edge and column, no neighbour rows
__global__ void __launch_bounds__(kAggThreads)
edge_message_bwd_target_mask_kernel(const float* __restrict__ dhid, int64_t ldd, const int32_t* __restrict__ rowptr,
                                    const uint8_t* __restrict__ mask, int64_t n_rows, int hidden,
                                    float* __restrict__ dP, int64_t lddp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = hidden >> 2;
  const int nchunk = (c4n + 31) >> 5;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t i = it / nchunk;
    const int chunk = (int)(it - i * nchunk);
    const int c4 = chunk * 32 + lane;
    if (c4 >= c4n) continue;
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float den = (float)max(end - beg, 1);
    const float4 d = f4_div(reinterpret_cast<const float4*>(dhid + i * ldd)[c4], den);
    const uint8_t* __restrict__ mrow = mask + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = beg; k < end; ++k) {
      const unsigned m = __ldg(mrow + (int64_t)k * c4n);
      acc.x += (m & 1u) ? d.x : 0.f;
      acc.y += (m & 2u) ? d.y : 0.f;
      acc.z += (m & 4u) ? d.z : 0.f;
      acc.w += (m & 8u) ? d.w : 0.f;
    }
    reinterpret_cast<float4*>(dP + i * lddp)[c4] = acc;
  }
}

// dQ[j,:] = sum over out-edges (j->i) of (1/deg_i) * dhid[i,:] * bit: one gathered row (dhid) + one mask byte per
// out-edge and column
__global__ void __launch_bounds__(kAggThreads)
edge_message_bwd_source_mask_kernel(const float* __restrict__ dhid, int64_t ldd, const int32_t* __restrict__ rowptr_t,
                                    const int32_t* __restrict__ rowptr_s, const int32_t* __restrict__ col_s,
                                    const int32_t* __restrict__ tpos_s, const uint8_t* __restrict__ mask, int64_t n_rows,
                                    int hidden, float* __restrict__ dQ, int64_t lddq) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = hidden >> 2;
  const int nchunk = (c4n + 31) >> 5;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t j = it / nchunk;
    const int chunk = (int)(it - j * nchunk);
    const int c4 = chunk * 32 + lane;
    if (c4 >= c4n) continue;
    const int beg = rowptr_s[j], end = rowptr_s[j + 1];
    const uint8_t* __restrict__ mrow = mask + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = beg; k < end; ++k) {
      const int i = col_s[k];
      const float den = (float)max(rowptr_t[i + 1] - rowptr_t[i], 1);
      const unsigned m = __ldg(mrow + (int64_t)tpos_s[k] * c4n);
      const float4 d = f4_div(reinterpret_cast<const float4*>(dhid + (int64_t)i * ldd)[c4], den);
      acc.x += (m & 1u) ? d.x : 0.f;
      acc.y += (m & 2u) ? d.y : 0.f;
      acc.z += (m & 4u) ? d.z : 0.f;
      acc.w += (m & 8u) ? d.w : 0.f;
    }
    reinterpret_cast<float4*>(dQ + j * lddq)[c4] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same three kernels writing fp16 operand PLANES instead of fp32 (include/stinet_b200.h, "dense layers on operand
// PLANES"): hid and dPQ are read by nothing but tensor-core GEMMs, so they never exist as fp32 matrices.  The plane
// scale needs max|result| BEFORE the first element is written; both have cheap upper bounds from a number the producing
// GEMM's epilogue already knows:
//     0 <= hid  = mean relu(P_i + Q_j)            <= 2 max|PQ|
//     |dP_i|   <= |dhid_i|,   |dQ_j| <= sum_{j->i} |dhid_i| / deg_i  <= max|dhid| * dq_factor,
//     dq_factor = max_j sum_{j->i} 1/deg_i   (a property of the edge set: stinet_csr_dq_factor)
// A bound within 2^16 of the true maximum costs no accuracy (22-bit planes, fp16 exponent range).

template <bool MASK>
__global__ void __launch_bounds__(kAggThreads)
edge_message_fwd_planes_kernel(const float* __restrict__ P, int64_t ldp, const float* __restrict__ Q, int64_t ldq,
                               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                               int hidden, const unsigned* __restrict__ pq_amax, __half* __restrict__ hi,
                               __half* __restrict__ lo, int64_t ldh, int32_t* __restrict__ exp_out,
                               uint8_t* __restrict__ mask) {
  const int sft = plane_shift(__float_as_uint(2.f * __uint_as_float(__ldg(pq_amax))));
  const float scale = plane_scale(sft);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = -sft;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = hidden >> 2;
  const int nchunk = (c4n + 31) >> 5;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t i = it / nchunk;
    const int chunk = (int)(it - i * nchunk);
    const int c4 = chunk * 32 + lane;
    if (c4 >= c4n) continue;
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float den = (float)max(end - beg, 1);
    const float4 p = reinterpret_cast<const float4*>(P + i * ldp)[c4];
    uint8_t* __restrict__ mrow = mask + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = beg;
    for (; k + 4 <= end; k += 4) {       // (eight rows in flight per lane was tried: 1.4x SLOWER -- register pressure)
      int j[4];
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) j[u] = col[k + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = reinterpret_cast<const float4*>(Q + (int64_t)j[u] * ldq)[c4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float sx = p.x + q[u].x, sy = p.y + q[u].y, sz = p.z + q[u].z, sw = p.w + q[u].w;
        acc.x += relu(sx); acc.y += relu(sy); acc.z += relu(sz); acc.w += relu(sw);
        if (MASK) mrow[(int64_t)(k + u) * c4n] = (uint8_t)nibble(sx, sy, sz, sw);
      }
    }
    for (; k < end; ++k) {
      const float4 q = reinterpret_cast<const float4*>(Q + (int64_t)col[k] * ldq)[c4];
      const float sx = p.x + q.x, sy = p.y + q.y, sz = p.z + q.z, sw = p.w + q.w;
      acc.x += relu(sx); acc.y += relu(sy); acc.z += relu(sz); acc.w += relu(sw);
      if (MASK) mrow[(int64_t)k * c4n] = (uint8_t)nibble(sx, sy, sz, sw);
    }
    split_store4(f4_div(acc, den), scale, hi + i * ldh + 4 * c4, lo != nullptr ? lo + i * ldh + 4 * c4 : nullptr);
  }
}

// the plane scale of dPQ = [dP | dQ] from max|dhid| and the edge set's dq_factor; both backward kernels derive the same
__device__ __forceinline__ int dpq_shift(const unsigned* dhid_amax, const float* dq_factor) {
  return plane_shift(__float_as_uint(__uint_as_float(__ldg(dhid_amax)) * fmaxf(1.f, __ldg(dq_factor))));
}

// COLSUM: the CTA also leaves the column sums of the dP rows it produced in colpart[blockIdx.x][hidden] (the bias gradient
// of the hoisted first Linear is sum_i dP_i; a second stage adds the CTA partials in fixed order).  Needs
// (gridDim.x * 8) % nchunk == 0, so that a warp keeps one 128-channel chunk for its whole row loop.
// A warp works on TWO rows per iteration and loads the decision bytes of up to eight in-edges of each before it uses any
// of them: the kernel is a pure stream (a row of dhid, its mask bytes, a row of planes out), and with one dependent load
// chain per warp it sat at 0.4 of the HBM rate; sixteen independent loads in flight per lane hide the latency.
__device__ __forceinline__ float4 masked_count_sum(const uint8_t* __restrict__ mrow, int c4n, int beg, int end, float4 d) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = beg; k0 < end; k0 += 8) {
    unsigned m[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) m[u] = (k0 + u < end) ? (unsigned)__ldg(mrow + (int64_t)(k0 + u) * c4n) : 0u;
#pragma unroll
    for (int u = 0; u < 8; ++u) {                       // same order as the one-by-one loop: bit-identical sums
      if (k0 + u < end) {
        acc.x += (m[u] & 1u) ? d.x : 0.f;
        acc.y += (m[u] & 2u) ? d.y : 0.f;
        acc.z += (m[u] & 4u) ? d.z : 0.f;
        acc.w += (m[u] & 8u) ? d.w : 0.f;
      }
    }
  }
  return acc;
}

template <bool COLSUM>
__global__ void __launch_bounds__(kAggThreads)
edge_message_bwd_target_planes_kernel(float* __restrict__ dhid, int64_t ldd, const int32_t* __restrict__ rowptr,
                                      const uint8_t* __restrict__ mask, int64_t n_rows, int hidden,
                                      const unsigned* __restrict__ dhid_amax, const float* __restrict__ dq_factor,
                                      __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldp,
                                      int32_t* __restrict__ exp_out, float* __restrict__ colpart) {
  // dhid[i,:] is overwritten with dhid[i,:] / deg_i: the source kernel that runs next gathers these rows once per
  // out-edge and would otherwise repeat the division (and two row-pointer loads) per edge
  __shared__ float4 csm[COLSUM ? kWarpsPerCta : 1][32];
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int sft = dpq_shift(dhid_amax, dq_factor);
  const float scale = plane_scale(sft);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = -sft;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = hidden >> 2;
  const int nchunk = (c4n + 31) >> 5;
  // unit = (pair of consecutive rows, chunk); a warp's chunk is fixed (nwarps % nchunk == 0), its row pairs stride by
  // nwarps / nchunk
  const int chunk = (int)(warp0 % nchunk);
  const int c4 = chunk * 32 + lane;
  const int64_t pairs = (n_rows + 1) >> 1;
  if (c4 < c4n) {
    const uint8_t* __restrict__ mrow = mask + c4;
    for (int64_t pr = warp0 / nchunk; pr < pairs; pr += nwarps / nchunk) {
      const int64_t i0 = 2 * pr, i1 = i0 + 1;
      const bool two = i1 < n_rows;
      const int b0 = rowptr[i0], e0 = rowptr[i0 + 1];
      const int e1 = two ? rowptr[i1 + 1] : e0;
      float4 d0 = reinterpret_cast<const float4*>(dhid + i0 * ldd)[c4];
      float4 d1 = two ? reinterpret_cast<const float4*>(dhid + i1 * ldd)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      d0 = f4_div(d0, (float)max(e0 - b0, 1));
      d1 = f4_div(d1, (float)max(e1 - e0, 1));
      reinterpret_cast<float4*>(dhid + i0 * ldd)[c4] = d0;
      if (two) reinterpret_cast<float4*>(dhid + i1 * ldd)[c4] = d1;
      const float4 a0 = masked_count_sum(mrow, c4n, b0, e0, d0);
      const float4 a1 = masked_count_sum(mrow, c4n, e0, e1, d1);
      split_store4(a0, scale, hi + i0 * ldp + 4 * c4, lo != nullptr ? lo + i0 * ldp + 4 * c4 : nullptr);
      if (two) split_store4(a1, scale, hi + i1 * ldp + 4 * c4, lo != nullptr ? lo + i1 * ldp + 4 * c4 : nullptr);
      if (COLSUM) csum = f4_add(f4_add(csum, a0), a1);
    }
  }
  if (COLSUM) {
    const int w = threadIdx.x >> 5;
    csm[w][lane] = csum;
    __syncthreads();
    for (int idx = threadIdx.x; idx < c4n; idx += kAggThreads) {
      const int ch = idx >> 5, l = idx & 31;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ww = 0; ww < kWarpsPerCta; ++ww)      // warps of this CTA that work on chunk `ch`, in warp order
        if ((int)(((int64_t)blockIdx.x * kWarpsPerCta + ww) % nchunk) == ch) t = f4_add(t, csm[ww][l]);
      reinterpret_cast<float4*>(colpart + (int64_t)blockIdx.x * hidden)[idx] = t;
    }
  }
}

// dQ[j,:] = sum over out-edges (j -> i) of ds[i,:] * bit, ds = dhid / deg (left behind by the target kernel): one gathered
// row and one mask byte per out-edge and column; four out-edges' indices, rows and masks in flight per lane
__global__ void __launch_bounds__(kAggThreads)
edge_message_bwd_source_planes_kernel(const float* __restrict__ ds, int64_t ldd, const int32_t* __restrict__ rowptr_s,
                                      const int32_t* __restrict__ col_s, const int32_t* __restrict__ tpos_s,
                                      const uint8_t* __restrict__ mask, int64_t n_rows, int hidden,
                                      const unsigned* __restrict__ dhid_amax, const float* __restrict__ dq_factor,
                                      __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldp) {
  const float scale = plane_scale(dpq_shift(dhid_amax, dq_factor));
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerCta;
  const int c4n = hidden >> 2;
  const int nchunk = (c4n + 31) >> 5;
  for (int64_t it = warp0; it < n_rows * nchunk; it += nwarps) {
    const int64_t j = it / nchunk;
    const int chunk = (int)(it - j * nchunk);
    const int c4 = chunk * 32 + lane;
    if (c4 >= c4n) continue;
    const int beg = rowptr_s[j], end = rowptr_s[j + 1];
    const uint8_t* __restrict__ mrow = mask + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = beg; k0 < end; k0 += 8) {
      int i[8], t[8];
      unsigned m[8];
      float4 d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool on = k0 + u < end;
        i[u] = on ? col_s[k0 + u] : -1;
        t[u] = on ? tpos_s[k0 + u] : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        m[u] = i[u] >= 0 ? (unsigned)__ldg(mrow + (int64_t)t[u] * c4n) : 0u;
        d[u] = i[u] >= 0 ? reinterpret_cast<const float4*>(ds + (int64_t)i[u] * ldd)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (i[u] >= 0) {
          acc.x += (m[u] & 1u) ? d[u].x : 0.f;
          acc.y += (m[u] & 2u) ? d[u].y : 0.f;
          acc.z += (m[u] & 4u) ? d[u].z : 0.f;
          acc.w += (m[u] & 8u) ? d[u].w : 0.f;
        }
      }
    }
    split_store4(acc, scale, hi + j * ldp + 4 * c4, lo != nullptr ? lo + j * ldp + 4 * c4 : nullptr);
  }
}

// dq_factor = max over vertices j of  sum over out-edges (j -> i) of 1 / max(deg_i, 1)   (>= 0; one thread per vertex)
__global__ void __launch_bounds__(256) csr_dq_factor_kernel(const int32_t* __restrict__ rowptr_t,
                                                            const int32_t* __restrict__ rowptr_s,
                                                            const int32_t* __restrict__ col_s, int64_t n,
                                                            unsigned* __restrict__ out) {
  float f = 0.f;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int k = rowptr_s[j]; k < rowptr_s[j + 1]; ++k) {
      const int i = col_s[k];
      t += 1.f / (float)max(rowptr_t[i + 1] - rowptr_t[i], 1);
    }
    f = fmaxf(f, t);
  }
  amax_publish(__float_as_uint(f), out);
}

inline bool vec_ok(int64_t channels, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> lds) {
  if (channels & 3) return false;
  for (auto p : ptrs)
    if (!aligned16(p)) return false;
  for (auto l : lds)
    if (l & 3) return false;
  return true;
}

// warps needed: one per (row, 128-channel chunk) on the vector path, one per row otherwise
inline int row_grid(int64_t n_rows, int64_t channels, bool vec) {
  const int64_t nchunk = vec ? ceil_div(channels >> 2, 32) : 1;
  return wave_grid(n_rows * nchunk, kWarpsPerCta, 8, 16);
}

}  // namespace stinet

using namespace stinet;

extern "C" int stinet_aggregate_fwd(const float* x, int64_t ldx, const int32_t* rowptr, const int32_t* col,
                                    const int32_t* eid, int64_t n_rows, int64_t n_items, int64_t channels,
                                    int reduce, float* out, int64_t ldo, int32_t* arg, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && rowptr && out && (col || n_items == 0), STINET_ERR_ARG, "aggregate_fwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && ldx >= channels && ldo >= channels, STINET_ERR_ARG,
                 "aggregate_fwd: bad shape");
  if (n_rows == 0) return STINET_OK;
  const int ch = (int)channels;
  const bool v = vec_ok(channels, {x, out}, {ldx, ldo});
  const int g = row_grid(n_rows, channels, v && reduce != STINET_REDUCE_MAX);
  switch (reduce) {
    case STINET_REDUCE_ADD:
      if (v) K(aggregate_fwd_kernel<STINET_REDUCE_ADD, true><<<g, kAggThreads, 0, s>>>(x, ldx, rowptr, col, eid, n_rows, (int32_t)n_items, ch, out, ldo, arg));
      else K(aggregate_fwd_kernel<STINET_REDUCE_ADD, false><<<g, kAggThreads, 0, s>>>(x, ldx, rowptr, col, eid, n_rows, (int32_t)n_items, ch, out, ldo, arg));
      break;
    case STINET_REDUCE_MEAN:
      if (v) K(aggregate_fwd_kernel<STINET_REDUCE_MEAN, true><<<g, kAggThreads, 0, s>>>(x, ldx, rowptr, col, eid, n_rows, (int32_t)n_items, ch, out, ldo, arg));
      else K(aggregate_fwd_kernel<STINET_REDUCE_MEAN, false><<<g, kAggThreads, 0, s>>>(x, ldx, rowptr, col, eid, n_rows, (int32_t)n_items, ch, out, ldo, arg));
      break;
    case STINET_REDUCE_MAX:
      STINET_REQUIRE(arg && eid, STINET_ERR_ARG, "aggregate_fwd(max): arg and eid are required");
      K(aggregate_fwd_kernel<STINET_REDUCE_MAX, false><<<g, kAggThreads, 0, s>>>(x, ldx, rowptr, col, eid, n_rows, (int32_t)n_items, ch, out, ldo, arg));
      break;
    default:
      STINET_REQUIRE(false, STINET_ERR_ARG, "aggregate_fwd: unknown reduce %d", reduce);
  }
  return check_launch("aggregate_fwd");
}

extern "C" int stinet_aggregate_bwd(const float* g_, int64_t ldg, const int32_t* rowptr_s, const int32_t* col_s,
                                    const int32_t* eid_s, const int32_t* rowptr_t, const int32_t* arg,
                                    int64_t n_rows, int64_t channels, int reduce, float* dx, int64_t lddx,
                                    stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(g_ && rowptr_s && dx, STINET_ERR_ARG, "aggregate_bwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && ldg >= channels && lddx >= channels, STINET_ERR_ARG,
                 "aggregate_bwd: bad shape");
  if (n_rows == 0) return STINET_OK;
  const int ch = (int)channels;
  const bool v = vec_ok(channels, {g_, dx}, {ldg, lddx});
  const int g = row_grid(n_rows, channels, v && reduce != STINET_REDUCE_MAX);
  switch (reduce) {
    case STINET_REDUCE_ADD:
      if (v) K(aggregate_bwd_kernel<STINET_REDUCE_ADD, true><<<g, kAggThreads, 0, s>>>(g_, ldg, rowptr_s, col_s, eid_s, rowptr_t, arg, n_rows, ch, dx, lddx));
      else K(aggregate_bwd_kernel<STINET_REDUCE_ADD, false><<<g, kAggThreads, 0, s>>>(g_, ldg, rowptr_s, col_s, eid_s, rowptr_t, arg, n_rows, ch, dx, lddx));
      break;
    case STINET_REDUCE_MEAN:
      STINET_REQUIRE(rowptr_t, STINET_ERR_ARG, "aggregate_bwd(mean): rowptr_t required");
      if (v) K(aggregate_bwd_kernel<STINET_REDUCE_MEAN, true><<<g, kAggThreads, 0, s>>>(g_, ldg, rowptr_s, col_s, eid_s, rowptr_t, arg, n_rows, ch, dx, lddx));
      else K(aggregate_bwd_kernel<STINET_REDUCE_MEAN, false><<<g, kAggThreads, 0, s>>>(g_, ldg, rowptr_s, col_s, eid_s, rowptr_t, arg, n_rows, ch, dx, lddx));
      break;
    case STINET_REDUCE_MAX:
      STINET_REQUIRE(arg && eid_s, STINET_ERR_ARG, "aggregate_bwd(max): arg and eid_s required");
      K(aggregate_bwd_kernel<STINET_REDUCE_MAX, false><<<g, kAggThreads, 0, s>>>(g_, ldg, rowptr_s, col_s, eid_s, rowptr_t, arg, n_rows, ch, dx, lddx));
      break;
    default:
      STINET_REQUIRE(false, STINET_ERR_ARG, "aggregate_bwd: unknown reduce %d", reduce);
  }
  return check_launch("aggregate_bwd");
}

extern "C" int stinet_edge_message_fwd(const float* P, int64_t ldp, const float* Q, int64_t ldq,
                                       const int32_t* rowptr_t, const int32_t* col_t, int64_t n_rows,
                                       int64_t hidden, float* hid, int64_t ldh, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(P && Q && rowptr_t && hid, STINET_ERR_ARG, "edge_message_fwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldp >= hidden && ldq >= hidden && ldh >= hidden, STINET_ERR_ARG,
                 "edge_message_fwd: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool v = vec_ok(hidden, {P, Q, hid}, {ldp, ldq, ldh});
  const int g = row_grid(n_rows, hidden, v);
  if (v)
    K(edge_message_fwd_kernel<true><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, rowptr_t, col_t, n_rows, (int)hidden, hid, ldh));
  else
    K(edge_message_fwd_kernel<false><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, rowptr_t, col_t, n_rows, (int)hidden, hid, ldh));
  return check_launch("edge_message_fwd");
}

extern "C" int stinet_edge_message_bwd_target(const float* P, int64_t ldp, const float* Q, int64_t ldq,
                                              const float* dhid, int64_t ldd, const int32_t* rowptr_t,
                                              const int32_t* col_t, int64_t n_rows, int64_t hidden, float* dP,
                                              int64_t lddp, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(P && Q && dhid && rowptr_t && dP, STINET_ERR_ARG, "edge_message_bwd_target: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldp >= hidden && ldq >= hidden && ldd >= hidden && lddp >= hidden,
                 STINET_ERR_ARG, "edge_message_bwd_target: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool v = vec_ok(hidden, {P, Q, dhid, dP}, {ldp, ldq, ldd, lddp});
  const int g = row_grid(n_rows, hidden, v);
  if (v)
    K(edge_message_bwd_target_kernel<true><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, dhid, ldd, rowptr_t, col_t, n_rows, (int)hidden, dP, lddp));
  else
    K(edge_message_bwd_target_kernel<false><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, dhid, ldd, rowptr_t, col_t, n_rows, (int)hidden, dP, lddp));
  return check_launch("edge_message_bwd_target");
}

extern "C" int stinet_edge_message_bwd_source(const float* P, int64_t ldp, const float* Q, int64_t ldq,
                                              const float* dhid, int64_t ldd, const int32_t* rowptr_t,
                                              const int32_t* rowptr_s, const int32_t* col_s, int64_t n_rows,
                                              int64_t hidden, float* dQ, int64_t lddq, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(P && Q && dhid && rowptr_t && rowptr_s && dQ, STINET_ERR_ARG, "edge_message_bwd_source: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldp >= hidden && ldq >= hidden && ldd >= hidden && lddq >= hidden,
                 STINET_ERR_ARG, "edge_message_bwd_source: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool v = vec_ok(hidden, {P, Q, dhid, dQ}, {ldp, ldq, ldd, lddq});
  const int g = row_grid(n_rows, hidden, v);
  if (v)
    K(edge_message_bwd_source_kernel<true><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, dhid, ldd, rowptr_t, rowptr_s, col_s, n_rows, (int)hidden, dQ, lddq));
  else
    K(edge_message_bwd_source_kernel<false><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, dhid, ldd, rowptr_t, rowptr_s, col_s, n_rows, (int)hidden, dQ, lddq));
  return check_launch("edge_message_bwd_source");
}

extern "C" int stinet_edge_message_fwd_mask(const float* P, int64_t ldp, const float* Q, int64_t ldq,
                                            const int32_t* rowptr_t, const int32_t* col_t, int64_t n_rows,
                                            int64_t hidden, float* hid, int64_t ldh, void* mask,
                                            stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(P && Q && rowptr_t && hid && mask, STINET_ERR_ARG, "edge_message_fwd_mask: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldp >= hidden && ldq >= hidden && ldh >= hidden, STINET_ERR_ARG,
                 "edge_message_fwd_mask: bad shape");
  STINET_REQUIRE(vec_ok(hidden, {P, Q, hid}, {ldp, ldq, ldh}), STINET_ERR_UNSUPPORTED,
                 "edge_message_fwd_mask: needs hidden %% 4 == 0 and 16-byte aligned rows");
  if (n_rows == 0) return STINET_OK;
  K(edge_message_fwd_mask_kernel<<<row_grid(n_rows, hidden, true), kAggThreads, 0, s>>>(
      P, ldp, Q, ldq, rowptr_t, col_t, n_rows, (int)hidden, hid, ldh, static_cast<uint8_t*>(mask)));
  return check_launch("edge_message_fwd_mask");
}

extern "C" int stinet_edge_message_bwd_target_mask(const float* dhid, int64_t ldd, const int32_t* rowptr_t,
                                                   const void* mask, int64_t n_rows, int64_t hidden, float* dP,
                                                   int64_t lddp, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(dhid && rowptr_t && mask && dP, STINET_ERR_ARG, "edge_message_bwd_target_mask: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldd >= hidden && lddp >= hidden, STINET_ERR_ARG,
                 "edge_message_bwd_target_mask: bad shape");
  STINET_REQUIRE(vec_ok(hidden, {dhid, dP}, {ldd, lddp}), STINET_ERR_UNSUPPORTED,
                 "edge_message_bwd_target_mask: needs hidden %% 4 == 0 and 16-byte aligned rows");
  if (n_rows == 0) return STINET_OK;
  K(edge_message_bwd_target_mask_kernel<<<row_grid(n_rows, hidden, true), kAggThreads, 0, s>>>(
      dhid, ldd, rowptr_t, static_cast<const uint8_t*>(mask), n_rows, (int)hidden, dP, lddp));
  return check_launch("edge_message_bwd_target_mask");
}

extern "C" int stinet_edge_message_bwd_source_mask(const float* dhid, int64_t ldd, const int32_t* rowptr_t,
                                                   const int32_t* rowptr_s, const int32_t* col_s,
                                                   const int32_t* tpos_s, const void* mask, int64_t n_rows,
                                                   int64_t hidden, float* dQ, int64_t lddq, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(dhid && rowptr_t && rowptr_s && tpos_s && mask && dQ, STINET_ERR_ARG,
                 "edge_message_bwd_source_mask: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldd >= hidden && lddq >= hidden, STINET_ERR_ARG,
                 "edge_message_bwd_source_mask: bad shape");
  STINET_REQUIRE(vec_ok(hidden, {dhid, dQ}, {ldd, lddq}), STINET_ERR_UNSUPPORTED,
                 "edge_message_bwd_source_mask: needs hidden %% 4 == 0 and 16-byte aligned rows");
  if (n_rows == 0) return STINET_OK;
  K(edge_message_bwd_source_mask_kernel<<<row_grid(n_rows, hidden, true), kAggThreads, 0, s>>>(
      dhid, ldd, rowptr_t, rowptr_s, col_s, tpos_s, static_cast<const uint8_t*>(mask), n_rows, (int)hidden, dQ, lddq));
  return check_launch("edge_message_bwd_source_mask");
}

extern "C" int stinet_csr_dq_factor(const int32_t* rowptr_t, const int32_t* rowptr_s, const int32_t* col_s, int64_t n,
                                    float* out, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(rowptr_t && rowptr_s && out, STINET_ERR_ARG, "csr_dq_factor: null pointer");
  STINET_REQUIRE(n >= 0, STINET_ERR_ARG, "csr_dq_factor: bad shape");
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), s);
  STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "csr_dq_factor: cudaMemsetAsync: %s", cudaGetErrorString(e));
  if (n == 0) return STINET_OK;
  K(csr_dq_factor_kernel<<<wave_grid(n, 256, 8), 256, 0, s>>>(rowptr_t, rowptr_s, col_s, n, reinterpret_cast<unsigned*>(out)));
  return check_launch("csr_dq_factor");
}

// grid of the target kernel: one warp per (row pair, 128-channel chunk), at most 4 waves, a multiple of the chunk count
// (so the warp count is one too and a warp keeps its chunk for its whole row loop)
static int edge_bwd_target_grid(int64_t n_rows, int64_t hidden) {
  const int nchunk = (int)ceil_div(hidden >> 2, 32);
  int g = wave_grid(((n_rows > 0 ? n_rows : 1) + 1) / 2 * nchunk, kWarpsPerCta, 8, 4);
  return (int)(ceil_div(g, nchunk) * nchunk);      // 8 * g warps, a multiple of the chunk count
}
static bool edge_bwd_fused_colsum(int64_t hidden) { return hidden <= 256; }
static size_t edge_bwd_colsum_bytes(int64_t n_rows, int64_t hidden) {
  if (!edge_bwd_fused_colsum(hidden)) return sizeof(float) * colsum_part_floats(n_rows, hidden);
  const int g = edge_bwd_target_grid(n_rows, hidden);
  return sizeof(float) * ((size_t)g * (size_t)hidden + colsum_part_floats(g, hidden));
}
extern "C" size_t stinet_edge_message_bwd_workspace_bytes(int64_t n_rows, int64_t hidden) {
  if (n_rows < 0 || hidden <= 0) return 0;
  return edge_bwd_colsum_bytes(n_rows, hidden);
}

static bool planes_ok(int64_t hidden, const void* hi, const void* lo, int64_t ld) {
  return !(hidden & 3) && !(ld & 7) && ld >= hidden && aligned16(hi) && (lo == nullptr || aligned16(lo));
}

extern "C" int stinet_edge_message_fwd_planes(const float* P, int64_t ldp, const float* Q, int64_t ldq,
                                              const int32_t* rowptr_t, const int32_t* col_t, int64_t n_rows,
                                              int64_t hidden, const float* pq_amax, void* hid_hi, void* hid_lo,
                                              int64_t ldh, int32_t* hid_exp, void* mask, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(P && Q && rowptr_t && pq_amax && hid_hi && hid_exp, STINET_ERR_ARG, "edge_message_fwd_planes: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldp >= hidden && ldq >= hidden, STINET_ERR_ARG, "edge_message_fwd_planes: bad shape");
  STINET_REQUIRE(vec_ok(hidden, {P, Q}, {ldp, ldq}) && planes_ok(hidden, hid_hi, hid_lo, ldh), STINET_ERR_UNSUPPORTED,
                 "edge_message_fwd_planes: needs hidden %% 4 == 0, 16-byte aligned rows, plane pitch %% 8 == 0");
  const int g = n_rows > 0 ? row_grid(n_rows, hidden, true) : 1;      // n_rows == 0 still publishes the exponent
  const unsigned* am = reinterpret_cast<const unsigned*>(pq_amax);
  if (mask != nullptr)
    K(edge_message_fwd_planes_kernel<true><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, rowptr_t, col_t, n_rows, (int)hidden, am,
        static_cast<__half*>(hid_hi), static_cast<__half*>(hid_lo), ldh, hid_exp, static_cast<uint8_t*>(mask)));
  else
    K(edge_message_fwd_planes_kernel<false><<<g, kAggThreads, 0, s>>>(P, ldp, Q, ldq, rowptr_t, col_t, n_rows, (int)hidden, am,
        static_cast<__half*>(hid_hi), static_cast<__half*>(hid_lo), ldh, hid_exp, nullptr));
  return check_launch("edge_message_fwd_planes");
}

extern "C" int stinet_edge_message_bwd_planes(float* dhid, int64_t ldd, const float* dhid_amax,
                                              const float* dq_factor, const int32_t* rowptr_t,
                                              const int32_t* rowptr_s, const int32_t* col_s, const int32_t* tpos_s,
                                              const void* mask, int64_t n_rows, int64_t hidden, void* dpq_hi,
                                              void* dpq_lo, int64_t ldp, int32_t* dpq_exp, float* dp_colsum,
                                              void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(dhid && dhid_amax && dq_factor && rowptr_t && rowptr_s && tpos_s && mask && dpq_hi && dpq_exp, STINET_ERR_ARG,
                 "edge_message_bwd_planes: null pointer");
  STINET_REQUIRE(n_rows >= 0 && hidden > 0 && ldd >= hidden && ldp >= 2 * hidden, STINET_ERR_ARG, "edge_message_bwd_planes: bad shape");
  STINET_REQUIRE(vec_ok(hidden, {dhid}, {ldd}) && planes_ok(2 * hidden, dpq_hi, dpq_lo, ldp), STINET_ERR_UNSUPPORTED,
                 "edge_message_bwd_planes: needs hidden %% 4 == 0, 16-byte aligned rows, plane pitch %% 8 == 0");
  const int g = n_rows > 0 ? row_grid(n_rows, hidden, true) : 1;
  const unsigned* am = reinterpret_cast<const unsigned*>(dhid_amax);
  __half* hi = static_cast<__half*>(dpq_hi);
  __half* lo = static_cast<__half*>(dpq_lo);
  const int gt = edge_bwd_target_grid(n_rows, hidden);
  if (dp_colsum != nullptr && !edge_bwd_fused_colsum(hidden)) {
    // wide rows (coarse levels: few rows, thousands of channels): a partial row per CTA would outweigh the data itself;
    // the column sums are taken from the finished dP planes instead
    const size_t need = edge_bwd_colsum_bytes(n_rows, hidden);
    STINET_REQUIRE(workspace && workspace_bytes >= need, STINET_ERR_WORKSPACE, "edge_message_bwd_planes: workspace %zu < %zu",
                   workspace_bytes, need);
    K(edge_message_bwd_target_planes_kernel<false><<<gt, kAggThreads, 0, s>>>(dhid, ldd, rowptr_t, static_cast<const uint8_t*>(mask),
                                                                             n_rows, (int)hidden, am, dq_factor, hi, lo, ldp, dpq_exp, nullptr));
    run_colsum_planes(hi, lo, ldp, dpq_exp, n_rows, hidden, dp_colsum, static_cast<float*>(workspace), s);
  } else if (dp_colsum != nullptr) {
    // dbias of the hoisted first Linear = column sums of dP: per-CTA partials from the target kernel, then the
    // deterministic two-stage column sum over the partial rows
    const size_t need = edge_bwd_colsum_bytes(n_rows, hidden);
    STINET_REQUIRE(workspace && workspace_bytes >= need, STINET_ERR_WORKSPACE, "edge_message_bwd_planes: workspace %zu < %zu",
                   workspace_bytes, need);
    float* part = static_cast<float*>(workspace);
    K(edge_message_bwd_target_planes_kernel<true><<<gt, kAggThreads, 0, s>>>(dhid, ldd, rowptr_t, static_cast<const uint8_t*>(mask),
                                                                            n_rows, (int)hidden, am, dq_factor, hi, lo, ldp, dpq_exp, part));
    run_colsum(part, hidden, nullptr, gt, hidden, dp_colsum, part + (size_t)gt * hidden, s);
  } else {
    K(edge_message_bwd_target_planes_kernel<false><<<gt, kAggThreads, 0, s>>>(dhid, ldd, rowptr_t, static_cast<const uint8_t*>(mask),
                                                                             n_rows, (int)hidden, am, dq_factor, hi, lo, ldp, dpq_exp, nullptr));
  }
  if (n_rows > 0)
    K(edge_message_bwd_source_planes_kernel<<<g, kAggThreads, 0, s>>>(
        dhid, ldd, rowptr_s, col_s, tpos_s, static_cast<const uint8_t*>(mask), n_rows, (int)hidden, am, dq_factor,
        hi + hidden, lo != nullptr ? lo + hidden : nullptr, ldp));
  return check_launch("edge_message_bwd_planes");
}
