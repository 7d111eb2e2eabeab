// Trace-map pooling (segmented max / mean over cluster ids) and unpooling (row gather, segmented-add backward).
// Cluster CSR rows list their fine members in ascending id, so a strict '>' scan reproduces torch_scatter's CPU
// scatter_max tie-break (lowest fine id wins) and sums follow index order.  No atomics; every output element has
// exactly one writer.
#include "common.cuh"

namespace stinet {

constexpr int kPoolWarps = 8;
constexpr int kPoolThreads = kPoolWarps * 32;
constexpr float kLowest = -3.402823466e+38f;

template <bool VEC>
__global__ void __launch_bounds__(kPoolThreads)
pool_max_fwd_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ member, int32_t n_fine, int64_t n_coarse, int channels,
                    float* __restrict__ out, int64_t ldo, int32_t* __restrict__ arg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t c = warp0; c < n_coarse; c += nwarps) {
    const int beg = rowptr[c], end = rowptr[c + 1];
    if (VEC) {
      const int c4n = channels >> 2;
      for (int c4 = lane; c4 < c4n; c4 += 32) {
        float4 best = make_float4(kLowest, kLowest, kLowest, kLowest);
        int4 bi = make_int4(n_fine, n_fine, n_fine, n_fine);
        for (int k = beg; k < end; ++k) {
          const int i = member[k];
          const float4 v = reinterpret_cast<const float4*>(x + (int64_t)i * ldx)[c4];
          if (v.x > best.x) { best.x = v.x; bi.x = i; }
          if (v.y > best.y) { best.y = v.y; bi.y = i; }
          if (v.z > best.z) { best.z = v.z; bi.z = i; }
          if (v.w > best.w) { best.w = v.w; bi.w = i; }
        }
        if (bi.x == n_fine) best.x = 0.f;
        if (bi.y == n_fine) best.y = 0.f;
        if (bi.z == n_fine) best.z = 0.f;
        if (bi.w == n_fine) best.w = 0.f;
        reinterpret_cast<float4*>(out + c * ldo)[c4] = best;
        reinterpret_cast<int4*>(arg + c * (int64_t)channels)[c4] = bi;
      }
    } else {
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
}

// dx[i,ch] = (arg[trace[i],ch] == i) ? g[trace[i],ch] : 0   -- gather form: coalesced, writes every element once
template <bool VEC>
__global__ void __launch_bounds__(kPoolThreads)
pool_max_bwd_kernel(const float* __restrict__ g, int64_t ldg, const int32_t* __restrict__ arg,
                    const int32_t* __restrict__ trace, int64_t n_fine, int channels, float* __restrict__ dx,
                    int64_t lddx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t i = warp0; i < n_fine; i += nwarps) {
    const int64_t c = trace[i];
    const int ii = (int)i;
    if (VEC) {
      const int c4n = channels >> 2;
      for (int c4 = lane; c4 < c4n; c4 += 32) {
        const int4 a = reinterpret_cast<const int4*>(arg + c * channels)[c4];
        const float4 v = reinterpret_cast<const float4*>(g + c * ldg)[c4];
        float4 o;
        o.x = (a.x == ii) ? v.x : 0.f;
        o.y = (a.y == ii) ? v.y : 0.f;
        o.z = (a.z == ii) ? v.z : 0.f;
        o.w = (a.w == ii) ? v.w : 0.f;
        reinterpret_cast<float4*>(dx + i * lddx)[c4] = o;
      }
    } else {
      for (int ch = lane; ch < channels; ch += 32)
        dx[i * lddx + ch] = (arg[c * channels + ch] == ii) ? g[c * ldg + ch] : 0.f;
    }
  }
}

// segmented sum over cluster members; MEAN divides by max(count,1).  Used for pool-mean fwd and unpool bwd.
template <bool MEAN, bool VEC>
__global__ void __launch_bounds__(kPoolThreads)
cluster_sum_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                   const int32_t* __restrict__ member, int64_t n_coarse, int channels, float* __restrict__ out,
                   int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t c = warp0; c < n_coarse; c += nwarps) {
    const int beg = rowptr[c], end = rowptr[c + 1];
    const float den = MEAN ? (float)max(end - beg, 1) : 1.f;
    if (VEC) {
      const int c4n = channels >> 2;
      for (int c4 = lane; c4 < c4n; c4 += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = beg; k < end; ++k) {
          const float4 v = reinterpret_cast<const float4*>(x + (int64_t)member[k] * ldx)[c4];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (MEAN) { acc.x /= den; acc.y /= den; acc.z /= den; acc.w /= den; }
        reinterpret_cast<float4*>(out + c * ldo)[c4] = acc;
      }
    } else {
      for (int ch = lane; ch < channels; ch += 32) {
        float acc = 0.f;
        for (int k = beg; k < end; ++k) acc += x[(int64_t)member[k] * ldx + ch];
        out[c * ldo + ch] = MEAN ? acc / den : acc;
      }
    }
  }
}

// out[i,:] = src[trace[i],:] (optionally divided by the cluster size: pool-mean backward)
template <bool MEAN, bool VEC>
__global__ void __launch_bounds__(kPoolThreads)
row_gather_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ trace,
                  const int32_t* __restrict__ rowptr, int64_t n_fine, int channels, float* __restrict__ out,
                  int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPoolWarps;
  for (int64_t i = warp0; i < n_fine; i += nwarps) {
    const int64_t c = trace[i];
    const float den = MEAN ? (float)max(rowptr[c + 1] - rowptr[c], 1) : 1.f;
    if (VEC) {
      const int c4n = channels >> 2;
      for (int c4 = lane; c4 < c4n; c4 += 32) {
        float4 v = reinterpret_cast<const float4*>(src + c * lds)[c4];
        if (MEAN) { v.x /= den; v.y /= den; v.z /= den; v.w /= den; }
        reinterpret_cast<float4*>(out + i * ldo)[c4] = v;
      }
    } else {
      for (int ch = lane; ch < channels; ch += 32) {
        float v = src[c * lds + ch];
        out[i * ldo + ch] = MEAN ? v / den : v;
      }
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
  if (vec4(channels, x, ldx, out, ldo) && aligned16(arg))
    K(pool_max_fwd_kernel<true><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(x, ldx, rowptr_c, member, (int32_t)n_fine, n_coarse, (int)channels, out, ldo, arg));
  else
    K(pool_max_fwd_kernel<false><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(x, ldx, rowptr_c, member, (int32_t)n_fine, n_coarse, (int)channels, out, ldo, arg));
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
  if (vec4(channels, g, ldg, dx, lddx) && aligned16(arg))
    K(pool_max_bwd_kernel<true><<<pgrid(n_fine), kPoolThreads, 0, s>>>(g, ldg, arg, trace32, n_fine, (int)channels, dx, lddx));
  else
    K(pool_max_bwd_kernel<false><<<pgrid(n_fine), kPoolThreads, 0, s>>>(g, ldg, arg, trace32, n_fine, (int)channels, dx, lddx));
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
  if (vec4(channels, x, ldx, out, ldo))
    K(cluster_sum_kernel<true, true><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(x, ldx, rowptr_c, member, n_coarse, (int)channels, out, ldo));
  else
    K(cluster_sum_kernel<true, false><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(x, ldx, rowptr_c, member, n_coarse, (int)channels, out, ldo));
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
  if (vec4(channels, g, ldg, dx, lddx))
    K(row_gather_kernel<true, true><<<pgrid(n_fine), kPoolThreads, 0, s>>>(g, ldg, trace32, rowptr_c, n_fine, (int)channels, dx, lddx));
  else
    K(row_gather_kernel<true, false><<<pgrid(n_fine), kPoolThreads, 0, s>>>(g, ldg, trace32, rowptr_c, n_fine, (int)channels, dx, lddx));
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
  if (vec4(channels, xc, ldc, out, ldo))
    K(row_gather_kernel<false, true><<<pgrid(n_fine), kPoolThreads, 0, s>>>(xc, ldc, trace32, nullptr, n_fine, (int)channels, out, ldo));
  else
    K(row_gather_kernel<false, false><<<pgrid(n_fine), kPoolThreads, 0, s>>>(xc, ldc, trace32, nullptr, n_fine, (int)channels, out, ldo));
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
  if (vec4(channels, g, ldg, dxc, ldd))
    K(cluster_sum_kernel<false, true><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(g, ldg, rowptr_c, member, n_coarse, (int)channels, dxc, ldd));
  else
    K(cluster_sum_kernel<false, false><<<pgrid(n_coarse), kPoolThreads, 0, s>>>(g, ldg, rowptr_c, member, n_coarse, (int)channels, dxc, ldd));
  return check_launch("unpool_bwd");
}
