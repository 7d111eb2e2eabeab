// Trace-map pooling (segmented max / mean over cluster ids) and unpooling (row gather, segmented-add backward).
// Cluster CSR rows list their fine members in ascending id, so a strict '>' scan reproduces torch_scatter's CPU
// scatter_max tie-break (lowest fine id wins) and sums follow index order.  No atomics; every output element has
// exactly one writer.
#include "common.cuh"

namespace stinet {

constexpr int kPoolWarps = 8;
constexpr int kPoolThreads = kPoolWarps * 32;
constexpr float kLowest = -3.402823466e+38f;

// Work decomposition shared by the row kernels below.  A TEAM of T lanes (T = 4, 8, 16 or 32: the smallest power of
// two that covers the row's float4 columns, at most a warp) owns one (row, 128-channel chunk) item, so a 64-channel
// row keeps both half-warps busy and a 1024-channel row spreads over eight warps.  Every lane owns exactly one float4
// column of its item; no cross-lane communication, so teams of one warp may diverge freely.
template <int T>
struct Teams {
  int tl;           // lane inside the team
  int64_t first;    // first item of this team
  int64_t stride;   // number of teams in the grid
  __device__ __forceinline__ Teams() {
    const int lane = threadIdx.x & 31;
    tl = lane % T;
    first = ((int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5)) * (32 / T) + lane / T;
    stride = (int64_t)gridDim.x * kPoolWarps * (32 / T);
  }
};

#define F4(ptr) reinterpret_cast<const float4*>(ptr)

__device__ __forceinline__ void max_step(float4& best, int4& bi, const float4& v, int i) {
  if (v.x > best.x) { best.x = v.x; bi.x = i; }
  if (v.y > best.y) { best.y = v.y; bi.y = i; }
  if (v.z > best.z) { best.z = v.z; bi.z = i; }
  if (v.w > best.w) { best.w = v.w; bi.w = i; }
}

// vector path: members are walked four at a time (four independent row loads in flight), compared in member order
template <int T>
__global__ void __launch_bounds__(kPoolThreads)
pool_max_fwd_vec_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                        const int32_t* __restrict__ member, int32_t n_fine, int64_t n_coarse, int channels,
                        float* __restrict__ out, int64_t ldo, int32_t* __restrict__ arg) {
  const Teams<T> tm;
  const int c4n = channels >> 2;
  const int nchunk = (c4n + 31) >> 5;
  for (int64_t it = tm.first; it < n_coarse * nchunk; it += tm.stride) {
    const int64_t c = it / nchunk;
    const int c4 = (int)(it - c * nchunk) * 32 + tm.tl;
    if (c4 >= c4n) continue;
    const int beg = rowptr[c], end = rowptr[c + 1];
    float4 best = make_float4(kLowest, kLowest, kLowest, kLowest);
    int4 bi = make_int4(n_fine, n_fine, n_fine, n_fine);
    int k = beg;
    for (; k + 4 <= end; k += 4) {
      const int i0 = member[k], i1 = member[k + 1], i2 = member[k + 2], i3 = member[k + 3];
      const float4 v0 = F4(x + (int64_t)i0 * ldx)[c4];
      const float4 v1 = F4(x + (int64_t)i1 * ldx)[c4];
      const float4 v2 = F4(x + (int64_t)i2 * ldx)[c4];
      const float4 v3 = F4(x + (int64_t)i3 * ldx)[c4];
      max_step(best, bi, v0, i0);
      max_step(best, bi, v1, i1);
      max_step(best, bi, v2, i2);
      max_step(best, bi, v3, i3);
    }
    if (k < end) {                       // 1..3 members left: issue the loads together as well
      const int i0 = member[k];
      const int i1 = k + 1 < end ? member[k + 1] : i0;
      const int i2 = k + 2 < end ? member[k + 2] : i0;
      const float4 v0 = F4(x + (int64_t)i0 * ldx)[c4];
      const float4 v1 = F4(x + (int64_t)i1 * ldx)[c4];
      const float4 v2 = F4(x + (int64_t)i2 * ldx)[c4];
      max_step(best, bi, v0, i0);
      if (k + 1 < end) max_step(best, bi, v1, i1);
      if (k + 2 < end) max_step(best, bi, v2, i2);
    }
    if (bi.x == n_fine) best.x = 0.f;
    if (bi.y == n_fine) best.y = 0.f;
    if (bi.z == n_fine) best.z = 0.f;
    if (bi.w == n_fine) best.w = 0.f;
    reinterpret_cast<float4*>(out + c * ldo)[c4] = best;
    reinterpret_cast<int4*>(arg + c * (int64_t)channels)[c4] = bi;
  }
}

__global__ void __launch_bounds__(kPoolThreads)
pool_max_fwd_scalar_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                           const int32_t* __restrict__ member, int32_t n_fine, int64_t n_coarse, int channels,
                           float* __restrict__ out, int64_t ldo, int32_t* __restrict__ arg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t c = warp0; c < n_coarse; c += nwarps) {
    const int beg = rowptr[c], end = rowptr[c + 1];
    for (int ch = lane; ch < channels; ch += 32) {
      float best = kLowest;
      int bi = n_fine;
      for (int k = beg; k < end; ++k) {
        const int i = member[k];
        const float v = x[(int64_t)i * ldx + ch];
        if (v > best) { best = v; bi = i; }
      }
      out[c * ldo + ch] = (bi == n_fine) ? 0.f : best;
      arg[c * (int64_t)channels + ch] = bi;
    }
  }
}

// dx[i,ch] = (arg[trace[i],ch] == i) ? g[trace[i],ch] : 0   -- gather form: coalesced, writes every element once.
// A team handles two fine rows per step so four independent loads are in flight behind the two trace reads.
template <int T>
__global__ void __launch_bounds__(kPoolThreads)
pool_max_bwd_vec_kernel(const float* __restrict__ g, int64_t ldg, const int32_t* __restrict__ arg,
                        const int32_t* __restrict__ trace, int64_t n_fine, int channels, float* __restrict__ dx,
                        int64_t lddx) {
  const Teams<T> tm;
  const int c4n = channels >> 2;
  const int nchunk = (c4n + 31) >> 5;
  const int64_t pairs = (n_fine + 1) >> 1;
  for (int64_t it = tm.first; it < pairs * nchunk; it += tm.stride) {
    const int64_t pr = it / nchunk;
    const int c4 = (int)(it - pr * nchunk) * 32 + tm.tl;
    if (c4 >= c4n) continue;
    const int64_t i0 = 2 * pr, i1 = i0 + 1;
    const bool two = i1 < n_fine;
    const int64_t c0 = trace[i0], c1 = two ? trace[i1] : c0;
    const int4 a0 = reinterpret_cast<const int4*>(arg + c0 * channels)[c4];
    const float4 v0 = F4(g + c0 * ldg)[c4];
    const int4 a1 = reinterpret_cast<const int4*>(arg + c1 * channels)[c4];
    const float4 v1 = F4(g + c1 * ldg)[c4];
    const int ii0 = (int)i0, ii1 = (int)i1;
    float4 o;
    o.x = (a0.x == ii0) ? v0.x : 0.f;
    o.y = (a0.y == ii0) ? v0.y : 0.f;
    o.z = (a0.z == ii0) ? v0.z : 0.f;
    o.w = (a0.w == ii0) ? v0.w : 0.f;
    reinterpret_cast<float4*>(dx + i0 * lddx)[c4] = o;
    if (two) {
      o.x = (a1.x == ii1) ? v1.x : 0.f;
      o.y = (a1.y == ii1) ? v1.y : 0.f;
      o.z = (a1.z == ii1) ? v1.z : 0.f;
      o.w = (a1.w == ii1) ? v1.w : 0.f;
      reinterpret_cast<float4*>(dx + i1 * lddx)[c4] = o;
    }
  }
}

__global__ void __launch_bounds__(kPoolThreads)
pool_max_bwd_scalar_kernel(const float* __restrict__ g, int64_t ldg, const int32_t* __restrict__ arg,
                           const int32_t* __restrict__ trace, int64_t n_fine, int channels, float* __restrict__ dx,
                           int64_t lddx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t i = warp0; i < n_fine; i += nwarps) {
    const int64_t c = trace[i];
    const int ii = (int)i;
    for (int ch = lane; ch < channels; ch += 32)
      dx[i * lddx + ch] = (arg[c * channels + ch] == ii) ? g[c * ldg + ch] : 0.f;
  }
}

// segmented sum over cluster members in ascending member order; MEAN divides by max(count,1).
// Used for pool-mean fwd and unpool bwd.
template <bool MEAN, int T>
__global__ void __launch_bounds__(kPoolThreads)
cluster_sum_vec_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                       const int32_t* __restrict__ member, int64_t n_coarse, int channels, float* __restrict__ out,
                       int64_t ldo) {
  const Teams<T> tm;
  const int c4n = channels >> 2;
  const int nchunk = (c4n + 31) >> 5;
  for (int64_t it = tm.first; it < n_coarse * nchunk; it += tm.stride) {
    const int64_t c = it / nchunk;
    const int c4 = (int)(it - c * nchunk) * 32 + tm.tl;
    if (c4 >= c4n) continue;
    const int beg = rowptr[c], end = rowptr[c + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = beg;
    for (; k + 4 <= end; k += 4) {
      const int i0 = member[k], i1 = member[k + 1], i2 = member[k + 2], i3 = member[k + 3];
      const float4 v0 = F4(x + (int64_t)i0 * ldx)[c4];
      const float4 v1 = F4(x + (int64_t)i1 * ldx)[c4];
      const float4 v2 = F4(x + (int64_t)i2 * ldx)[c4];
      const float4 v3 = F4(x + (int64_t)i3 * ldx)[c4];
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
      acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
      acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
    }
    if (k < end) {
      const int i0 = member[k];
      const int i1 = k + 1 < end ? member[k + 1] : i0;
      const int i2 = k + 2 < end ? member[k + 2] : i0;
      const float4 v0 = F4(x + (int64_t)i0 * ldx)[c4];
      const float4 v1 = F4(x + (int64_t)i1 * ldx)[c4];
      const float4 v2 = F4(x + (int64_t)i2 * ldx)[c4];
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      if (k + 1 < end) { acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w; }
      if (k + 2 < end) { acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w; }
    }
    if (MEAN) {
      const float den = (float)max(end - beg, 1);
      acc.x /= den; acc.y /= den; acc.z /= den; acc.w /= den;
    }
    reinterpret_cast<float4*>(out + c * ldo)[c4] = acc;
  }
}

template <bool MEAN>
__global__ void __launch_bounds__(kPoolThreads)
cluster_sum_scalar_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                          const int32_t* __restrict__ member, int64_t n_coarse, int channels, float* __restrict__ out,
                          int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t c = warp0; c < n_coarse; c += nwarps) {
    const int beg = rowptr[c], end = rowptr[c + 1];
    const float den = MEAN ? (float)max(end - beg, 1) : 1.f;
    for (int ch = lane; ch < channels; ch += 32) {
      float acc = 0.f;
      for (int k = beg; k < end; ++k) acc += x[(int64_t)member[k] * ldx + ch];
      out[c * ldo + ch] = MEAN ? acc / den : acc;
    }
  }
}

// out[i,:] = src[trace[i],:] (optionally divided by the cluster size: pool-mean backward); two fine rows per step
template <bool MEAN, int T>
__global__ void __launch_bounds__(kPoolThreads)
row_gather_vec_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ trace,
                      const int32_t* __restrict__ rowptr, int64_t n_fine, int channels, float* __restrict__ out,
                      int64_t ldo) {
  const Teams<T> tm;
  const int c4n = channels >> 2;
  const int nchunk = (c4n + 31) >> 5;
  const int64_t pairs = (n_fine + 1) >> 1;
  for (int64_t it = tm.first; it < pairs * nchunk; it += tm.stride) {
    const int64_t pr = it / nchunk;
    const int c4 = (int)(it - pr * nchunk) * 32 + tm.tl;
    if (c4 >= c4n) continue;
    const int64_t i0 = 2 * pr, i1 = i0 + 1;
    const bool two = i1 < n_fine;
    const int64_t c0 = trace[i0], c1 = two ? trace[i1] : c0;
    float4 v0 = F4(src + c0 * lds)[c4];
    float4 v1 = F4(src + c1 * lds)[c4];
    if (MEAN) {
      const float d0 = (float)max(rowptr[c0 + 1] - rowptr[c0], 1), d1 = (float)max(rowptr[c1 + 1] - rowptr[c1], 1);
      v0.x /= d0; v0.y /= d0; v0.z /= d0; v0.w /= d0;
      v1.x /= d1; v1.y /= d1; v1.z /= d1; v1.w /= d1;
    }
    reinterpret_cast<float4*>(out + i0 * ldo)[c4] = v0;
    if (two) reinterpret_cast<float4*>(out + i1 * ldo)[c4] = v1;
  }
}

template <bool MEAN>
__global__ void __launch_bounds__(kPoolThreads)
row_gather_scalar_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ trace,
                         const int32_t* __restrict__ rowptr, int64_t n_fine, int channels, float* __restrict__ out,
                         int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t i = warp0; i < n_fine; i += nwarps) {
    const int64_t c = trace[i];
    const float den = MEAN ? (float)max(rowptr[c + 1] - rowptr[c], 1) : 1.f;
    for (int ch = lane; ch < channels; ch += 32) {
      float v = src[c * lds + ch];
      out[i * ldo + ch] = MEAN ? v / den : v;
    }
  }
}

__global__ void pool_max_i32_kernel(const int32_t* __restrict__ v, const int32_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ member, int64_t n_coarse, int32_t* __restrict__ out) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_coarse;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int beg = rowptr[c], end = rowptr[c + 1];
    int best = INT32_MIN;
    for (int k = beg; k < end; ++k) best = max(best, v[member[k]]);
    out[c] = (end > beg) ? best : 0;
  }
}

inline bool vec4(int64_t channels, const void* a, int64_t lda, const void* b, int64_t ldb) {
  return !(channels & 3) && !(lda & 3) && !(ldb & 3) && aligned16(a) && aligned16(b);
}
inline int pgrid(int64_t rows) { return wave_grid(rows, kPoolWarps, 8, 16); }
// team width for a row of `channels` floats and the grid that gives every (row, chunk) item of `rows` rows one team
inline int team_width(int64_t channels) {
  const int64_t c4n = channels >> 2;
  return c4n <= 4 ? 4 : c4n <= 8 ? 8 : c4n <= 16 ? 16 : 32;
}
inline int tgrid(int64_t rows, int64_t channels) {
  const int t = team_width(channels);
  const int64_t items = rows * ceil_div(channels >> 2, 32);
  return pgrid(ceil_div(items, 32 / t));
}
// launch KERNEL<..., T> with T = team_width(channels)
#define TEAM_DISPATCH(channels, LAUNCH)           \
  switch (team_width(channels)) {                 \
    case 4: { constexpr int T_ = 4; LAUNCH; } break;   \
    case 8: { constexpr int T_ = 8; LAUNCH; } break;   \
    case 16: { constexpr int T_ = 16; LAUNCH; } break; \
    default: { constexpr int T_ = 32; LAUNCH; } break; \
  }

}  // namespace stinet

using namespace stinet;

extern "C" int stinet_pool_max_fwd(const float* x, int64_t ldx, const int32_t* rowptr_c, const int32_t* member,
                                   int64_t n_fine, int64_t n_coarse, int64_t channels, float* out, int64_t ldo,
                                   int32_t* arg, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && rowptr_c && out && arg, STINET_ERR_ARG, "pool_max_fwd: null pointer");
  STINET_REQUIRE(n_fine >= 0 && n_coarse >= 0 && channels > 0 && ldx >= channels && ldo >= channels, STINET_ERR_ARG,
                 "pool_max_fwd: bad shape");
  if (n_coarse == 0) return STINET_OK;
  if (vec4(channels, x, ldx, out, ldo) && aligned16(arg)) {
    const int grid = tgrid(n_coarse, channels);
    TEAM_DISPATCH(channels, K(pool_max_fwd_vec_kernel<T_><<<grid, kPoolThreads, 0, s>>>(
                                x, ldx, rowptr_c, member, (int32_t)n_fine, n_coarse, (int)channels, out, ldo, arg)));
  } else {
    K(pool_max_fwd_scalar_kernel<<<pgrid(n_coarse), kPoolThreads, 0, s>>>(x, ldx, rowptr_c, member, (int32_t)n_fine, n_coarse, (int)channels, out, ldo, arg));
  }
  return check_launch("pool_max_fwd");
}

extern "C" int stinet_pool_max_bwd(const float* g, int64_t ldg, const int32_t* arg, const int32_t* trace32,
                                   int64_t n_fine, int64_t channels, float* dx, int64_t lddx,
                                   stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(g && arg && trace32 && dx, STINET_ERR_ARG, "pool_max_bwd: null pointer");
  STINET_REQUIRE(n_fine >= 0 && channels > 0 && ldg >= channels && lddx >= channels, STINET_ERR_ARG,
                 "pool_max_bwd: bad shape");
  if (n_fine == 0) return STINET_OK;
  if (vec4(channels, g, ldg, dx, lddx) && aligned16(arg)) {
    const int grid = tgrid((n_fine + 1) / 2, channels);
    TEAM_DISPATCH(channels, K(pool_max_bwd_vec_kernel<T_><<<grid, kPoolThreads, 0, s>>>(g, ldg, arg, trace32, n_fine,
                                                                                       (int)channels, dx, lddx)));
  } else {
    K(pool_max_bwd_scalar_kernel<<<pgrid(n_fine), kPoolThreads, 0, s>>>(g, ldg, arg, trace32, n_fine, (int)channels, dx, lddx));
  }
  return check_launch("pool_max_bwd");
}

extern "C" int stinet_pool_mean_fwd(const float* x, int64_t ldx, const int32_t* rowptr_c, const int32_t* member,
                                    int64_t n_coarse, int64_t channels, float* out, int64_t ldo,
                                    stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && rowptr_c && out, STINET_ERR_ARG, "pool_mean_fwd: null pointer");
  STINET_REQUIRE(n_coarse >= 0 && channels > 0 && ldx >= channels && ldo >= channels, STINET_ERR_ARG,
                 "pool_mean_fwd: bad shape");
  if (n_coarse == 0) return STINET_OK;
  if (vec4(channels, x, ldx, out, ldo)) {
    const int grid = tgrid(n_coarse, channels);
    TEAM_DISPATCH(channels, K(cluster_sum_vec_kernel<true, T_><<<grid, kPoolThreads, 0, s>>>(
                                x, ldx, rowptr_c, member, n_coarse, (int)channels, out, ldo)));
  } else {
    K(cluster_sum_scalar_kernel<true><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(x, ldx, rowptr_c, member, n_coarse, (int)channels, out, ldo));
  }
  return check_launch("pool_mean_fwd");
}

extern "C" int stinet_pool_mean_bwd(const float* g, int64_t ldg, const int32_t* rowptr_c, const int32_t* trace32,
                                    int64_t n_fine, int64_t channels, float* dx, int64_t lddx,
                                    stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(g && rowptr_c && trace32 && dx, STINET_ERR_ARG, "pool_mean_bwd: null pointer");
  STINET_REQUIRE(n_fine >= 0 && channels > 0 && ldg >= channels && lddx >= channels, STINET_ERR_ARG,
                 "pool_mean_bwd: bad shape");
  if (n_fine == 0) return STINET_OK;
  if (vec4(channels, g, ldg, dx, lddx)) {
    const int grid = tgrid((n_fine + 1) / 2, channels);
    TEAM_DISPATCH(channels, K(row_gather_vec_kernel<true, T_><<<grid, kPoolThreads, 0, s>>>(
                                g, ldg, trace32, rowptr_c, n_fine, (int)channels, dx, lddx)));
  } else {
    K(row_gather_scalar_kernel<true><<<pgrid(n_fine), kPoolThreads, 0, s>>>(g, ldg, trace32, rowptr_c, n_fine, (int)channels, dx, lddx));
  }
  return check_launch("pool_mean_bwd");
}

extern "C" int stinet_pool_max_i32(const int32_t* v, const int32_t* rowptr_c, const int32_t* member,
                                   int64_t n_coarse, int32_t* out, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(v && rowptr_c && out, STINET_ERR_ARG, "pool_max_i32: null pointer");
  if (n_coarse <= 0) return STINET_OK;
  K(pool_max_i32_kernel<<<wave_grid(n_coarse, 256, 8), 256, 0, s>>>(v, rowptr_c, member, n_coarse, out));
  return check_launch("pool_max_i32");
}

extern "C" int stinet_unpool_fwd(const float* xc, int64_t ldc, const int32_t* trace32, int64_t n_fine,
                                 int64_t channels, float* out, int64_t ldo, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(xc && trace32 && out, STINET_ERR_ARG, "unpool_fwd: null pointer");
  STINET_REQUIRE(n_fine >= 0 && channels > 0 && ldc >= channels && ldo >= channels, STINET_ERR_ARG,
                 "unpool_fwd: bad shape");
  if (n_fine == 0) return STINET_OK;
  if (vec4(channels, xc, ldc, out, ldo)) {
    const int grid = tgrid((n_fine + 1) / 2, channels);
    TEAM_DISPATCH(channels, K(row_gather_vec_kernel<false, T_><<<grid, kPoolThreads, 0, s>>>(
                                xc, ldc, trace32, nullptr, n_fine, (int)channels, out, ldo)));
  } else {
    K(row_gather_scalar_kernel<false><<<pgrid(n_fine), kPoolThreads, 0, s>>>(xc, ldc, trace32, nullptr, n_fine, (int)channels, out, ldo));
  }
  return check_launch("unpool_fwd");
}

extern "C" int stinet_unpool_bwd(const float* g, int64_t ldg, const int32_t* rowptr_c, const int32_t* member,
                                 int64_t n_coarse, int64_t channels, float* dxc, int64_t ldd,
                                 stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(g && rowptr_c && dxc, STINET_ERR_ARG, "unpool_bwd: null pointer");
  STINET_REQUIRE(n_coarse >= 0 && channels > 0 && ldg >= channels && ldd >= channels, STINET_ERR_ARG,
                 "unpool_bwd: bad shape");
  if (n_coarse == 0) return STINET_OK;
  if (vec4(channels, g, ldg, dxc, ldd)) {
    const int grid = tgrid(n_coarse, channels);
    TEAM_DISPATCH(channels, K(cluster_sum_vec_kernel<false, T_><<<grid, kPoolThreads, 0, s>>>(
                                g, ldg, rowptr_c, member, n_coarse, (int)channels, dxc, ldd)));
  } else {
    K(cluster_sum_scalar_kernel<false><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(g, ldg, rowptr_c, member, n_coarse, (int)channels, dxc, ldd));
  }
  return check_launch("unpool_bwd");
}
