// Per-graph instance norm (FastInstanceNorm semantics) fused with the ELU + residual tail of GraphResnetBlock.
//   stats : two deterministic segmented column reductions (sum, then centred sum of squares -- the reference is
//           two-pass as well), each = per-chunk partials + a fixed-order finalize;
//   apply : out = residual + act((x - mean[g]) * rstd[g])            (one read of x, one write)
//   bwd   : one reduction pass (sum dz, sum dz*yhat) + one apply pass.
#include "common.cuh"

namespace stinet {

constexpr int kNormThreads = 256;
constexpr int kChunkRows = 256;   // rows per partial
constexpr int kTileGroups = 32;   // channel groups (float4 or scalar) per CTA tile

enum { MODE_SUM = 0, MODE_CSQ = 1, MODE_BWD = 2 };

struct NormArgs {
  const float* x; int64_t ldx;
  const float* dout; int64_t ldg;
  const int32_t* slice_ptr; const int32_t* gid;
  const float* mean; const float* rstd;
  float* part0; float* part1;
  int channels; int max_chunks; int act;
};

template <int MODE, bool VEC>
__global__ void __launch_bounds__(kNormThreads) seg_colreduce_kernel(NormArgs a) {
  constexpr int W = VEC ? 4 : 1;
  const int groups = a.channels / W;
  const int txw = min(groups, kTileGroups);         // threads along channels
  const int tyn = kNormThreads / txw;                // rows in flight
  const int tx = threadIdx.x % txw, ty = threadIdx.x / txw;
  const int s = blockIdx.y;
  const int grp = blockIdx.z * kTileGroups + tx;
  const int r0 = a.slice_ptr[s] + blockIdx.x * kChunkRows;
  const int r1 = min(r0 + kChunkRows, a.slice_ptr[s + 1]);
  float acc0[W], acc1[W];
#pragma unroll
  for (int w = 0; w < W; ++w) acc0[w] = acc1[w] = 0.f;
  if (grp < groups && ty < tyn) {
    for (int r = r0 + ty; r < r1; r += tyn) {
      float v[W], d[W], m[W], rs[W];
      const int g = a.gid ? a.gid[r] : s;
      if (VEC) {
        float4 t = reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[grp];
        v[0] = t.x; v[1 % W] = t.y; v[2 % W] = t.z; v[3 % W] = t.w;
      } else {
        v[0] = a.x[(int64_t)r * a.ldx + grp];
      }
      if (MODE != MODE_SUM) {
#pragma unroll
        for (int w = 0; w < W; ++w) m[w] = a.mean ? a.mean[(int64_t)g * a.channels + grp * W + w] : 0.f;
      }
      if (MODE == MODE_BWD) {
        if (VEC) {
          float4 t = reinterpret_cast<const float4*>(a.dout + (int64_t)r * a.ldg)[grp];
          d[0] = t.x; d[1 % W] = t.y; d[2 % W] = t.z; d[3 % W] = t.w;
        } else {
          d[0] = a.dout[(int64_t)r * a.ldg + grp];
        }
#pragma unroll
        for (int w = 0; w < W; ++w) rs[w] = a.rstd ? a.rstd[(int64_t)g * a.channels + grp * W + w] : 1.f;
      }
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if (MODE == MODE_SUM) {
          acc0[w] += v[w];
        } else if (MODE == MODE_CSQ) {
          const float c = v[w] - m[w];
          acc0[w] += c * c;
        } else {
          const float yh = (v[w] - m[w]) * rs[w];
          const float dz = (a.act == STINET_ACT_ELU) ? d[w] * elu1_grad(yh) : d[w];
          acc0[w] += dz;
          acc1[w] += dz * yh;
        }
      }
    }
  }
  // fixed-order reduction over ty
  __shared__ float sm0[kNormThreads * 4];
  __shared__ float sm1[kNormThreads * 4];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    sm0[threadIdx.x * W + w] = acc0[w];
    if (MODE == MODE_BWD) sm1[threadIdx.x * W + w] = acc1[w];
  }
  __syncthreads();
  if (ty == 0 && grp < groups) {
    const int64_t o = ((int64_t)s * a.max_chunks + blockIdx.x) * a.channels + grp * W;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      float t0 = 0.f, t1 = 0.f;
      for (int y = 0; y < tyn; ++y) {
        t0 += sm0[(y * txw + tx) * W + w];
        if (MODE == MODE_BWD) t1 += sm1[(y * txw + tx) * W + w];
      }
      a.part0[o + w] = t0;
      if (MODE == MODE_BWD) a.part1[o + w] = t1;
    }
  }
}

// FIN: 0 -> out = sum/cnt ; 1 -> out = 1/sqrt(sum/cnt + eps).  One CTA per (32 channels, segment): 32 channel lanes x
// 32 chunk lanes read the partials coalesced, then a fixed-order sum over the chunk lanes (deterministic).
template <int FIN>
__global__ void __launch_bounds__(1024)
seg_finalize_kernel(const float* __restrict__ part, const int32_t* __restrict__ slice_ptr,
                    const float* __restrict__ cnt, int n_seg, int channels, int max_chunks, float eps,
                    float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int s = blockIdx.y;
  const int c = blockIdx.x * 32 + tx;
  const int len = slice_ptr[s + 1] - slice_ptr[s];
  const int chunks = (len + kChunkRows - 1) / kChunkRows;
  float t = 0.f;
  if (c < channels)
    for (int k = ty; k < chunks; k += 32) t += part[((int64_t)s * max_chunks + k) * channels + c];
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && c < channels) {
    float r = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) r += sm[y][tx];
    const float m = r / cnt[s];
    out[(int64_t)s * channels + c] = FIN == 0 ? m : 1.f / sqrtf(m + eps);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kNormThreads)
segnorm_apply_kernel(const float* __restrict__ x, int64_t ldx, int64_t n_rows, int channels,
                     const int32_t* __restrict__ gid, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ res, int64_t ldr, int act, float* __restrict__ out, int64_t ldo) {
  constexpr int W = VEC ? 4 : 1;
  const int groups = channels / W;
  const int64_t total = n_rows * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / groups;
    const int grp = (int)(idx - r * groups);
    const int g = gid ? gid[r] : 0;
    float v[W], o[W];
    if (VEC) {
      float4 t = reinterpret_cast<const float4*>(x + r * ldx)[grp];
      v[0] = t.x; v[1 % W] = t.y; v[2 % W] = t.z; v[3 % W] = t.w;
    } else {
      v[0] = x[r * ldx + grp];
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int c = grp * W + w;
      float yh = v[w];
      if (mean) yh = (yh - mean[(int64_t)g * channels + c]) * rstd[(int64_t)g * channels + c];
      o[w] = (act == STINET_ACT_ELU) ? elu1(yh) : yh;
    }
    if (res) {
      if (VEC) {
        float4 t = reinterpret_cast<const float4*>(res + r * ldr)[grp];
        o[0] += t.x; o[1 % W] += t.y; o[2 % W] += t.z; o[3 % W] += t.w;
      } else {
        o[0] += res[r * ldr + grp];
      }
    }
    if (VEC) reinterpret_cast<float4*>(out + r * ldo)[grp] = make_float4(o[0], o[1 % W], o[2 % W], o[3 % W]);
    else out[r * ldo + grp] = o[0];
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kNormThreads)
segnorm_bwd_apply_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dout, int64_t ldg,
                         int64_t n_rows, int channels, const int32_t* __restrict__ gid,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ s1, const float* __restrict__ s2, int act, float* __restrict__ dx,
                         int64_t lddx) {
  constexpr int W = VEC ? 4 : 1;
  const int groups = channels / W;
  const int64_t total = n_rows * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / groups;
    const int grp = (int)(idx - r * groups);
    const int g = gid ? gid[r] : 0;
    float v[W], d[W], o[W];
    if (VEC) {
      float4 t = reinterpret_cast<const float4*>(x + r * ldx)[grp];
      v[0] = t.x; v[1 % W] = t.y; v[2 % W] = t.z; v[3 % W] = t.w;
      float4 u = reinterpret_cast<const float4*>(dout + r * ldg)[grp];
      d[0] = u.x; d[1 % W] = u.y; d[2 % W] = u.z; d[3 % W] = u.w;
    } else {
      v[0] = x[r * ldx + grp];
      d[0] = dout[r * ldg + grp];
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int64_t sc = (int64_t)g * channels + grp * W + w;
      if (mean) {
        const float rs = rstd[sc];
        const float yh = (v[w] - mean[sc]) * rs;
        const float dz = (act == STINET_ACT_ELU) ? d[w] * elu1_grad(yh) : d[w];
        o[w] = rs * (dz - s1[sc] - yh * s2[sc]);
      } else {
        o[w] = (act == STINET_ACT_ELU) ? d[w] * elu1_grad(v[w]) : d[w];
      }
    }
    if (VEC) reinterpret_cast<float4*>(dx + r * lddx)[grp] = make_float4(o[0], o[1 % W], o[2 % W], o[3 % W]);
    else dx[r * lddx + grp] = o[0];
  }
}

struct NormWs {
  float *part0, *part1, *s1, *s2;
  size_t bytes;
  int max_chunks;
};
static NormWs carve_norm(void* base, int64_t max_seg_rows, int64_t channels, int64_t n_seg) {
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  NormWs w;
  w.max_chunks = (int)ceil_div(max_seg_rows > 0 ? max_seg_rows : 1, kChunkRows);
  size_t part = up(sizeof(float) * (size_t)n_seg * w.max_chunks * channels);
  size_t st = up(sizeof(float) * (size_t)n_seg * channels);
  char* p = static_cast<char*>(base);
  w.part0 = reinterpret_cast<float*>(p);
  w.part1 = reinterpret_cast<float*>(p + part);
  w.s1 = reinterpret_cast<float*>(p + 2 * part);
  w.s2 = reinterpret_cast<float*>(p + 2 * part + st);
  w.bytes = 2 * part + 2 * st;
  return w;
}

inline bool nvec(int64_t channels, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> lds) {
  if (channels & 3) return false;
  for (auto p : ptrs)
    if (p && !aligned16(p)) return false;
  for (auto l : lds)
    if (l & 3) return false;
  return true;
}

template <int MODE>
static void launch_colreduce(bool vec, const NormArgs& a, int64_t n_seg, cudaStream_t s) {
  const int W = vec ? 4 : 1;
  const int groups = a.channels / W;
  dim3 grid(a.max_chunks, (unsigned)n_seg, (unsigned)ceil_div(groups, kTileGroups));
  if (vec) K(seg_colreduce_kernel<MODE, true><<<grid, kNormThreads, 0, s>>>(a));
  else K(seg_colreduce_kernel<MODE, false><<<grid, kNormThreads, 0, s>>>(a));
}

}  // namespace stinet

using namespace stinet;

extern "C" size_t stinet_segnorm_workspace_bytes(int64_t max_seg_rows, int64_t channels, int64_t n_seg) {
  if (max_seg_rows < 0 || channels <= 0 || n_seg <= 0) return 0;
  return carve_norm(nullptr, max_seg_rows, channels, n_seg).bytes;
}

extern "C" int stinet_segnorm_stats(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, int64_t n_seg,
                                    int64_t max_seg_rows, const int32_t* slice_ptr, const float* cnt,
                                    const int32_t* gid, float eps, float* mean, float* rstd, void* workspace,
                                    size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && slice_ptr && cnt && mean && rstd, STINET_ERR_ARG, "segnorm_stats: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && n_seg > 0 && ldx >= channels && n_seg <= 65535, STINET_ERR_ARG,
                 "segnorm_stats: bad shape");
  NormWs w = carve_norm(workspace, max_seg_rows, channels, n_seg);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "segnorm_stats: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  const bool vec = nvec(channels, {x}, {ldx});
  NormArgs a{x, ldx, nullptr, 0, slice_ptr, gid, nullptr, nullptr, w.part0, w.part1, (int)channels, w.max_chunks, 0};
  const dim3 fin_grid((unsigned)ceil_div(channels, 32), (unsigned)n_seg);
  launch_colreduce<MODE_SUM>(vec, a, n_seg, s);
  K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.part0, slice_ptr, cnt, (int)n_seg, (int)channels, w.max_chunks, eps, mean));
  a.mean = mean;
  launch_colreduce<MODE_CSQ>(vec, a, n_seg, s);
  K(seg_finalize_kernel<1><<<fin_grid, 1024, 0, s>>>(w.part0, slice_ptr, cnt, (int)n_seg, (int)channels, w.max_chunks, eps, rstd));
  return check_launch("segnorm_stats");
}

extern "C" int stinet_segnorm_apply(const float* x, int64_t ldx, int64_t n_rows, int64_t channels,
                                    const int32_t* gid, const float* mean, const float* rstd, const float* residual,
                                    int64_t ldr, int act, float* out, int64_t ldo, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && out && ((mean == nullptr) == (rstd == nullptr)), STINET_ERR_ARG, "segnorm_apply: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && ldx >= channels && ldo >= channels && (!residual || ldr >= channels),
                 STINET_ERR_ARG, "segnorm_apply: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool vec = nvec(channels, {x, out, residual}, {ldx, ldo, residual ? ldr : 0});
  const int grid = wave_grid(n_rows * (channels / (vec ? 4 : 1)), kNormThreads * 4, 8, 8);
  if (vec) K(segnorm_apply_kernel<true><<<grid, kNormThreads, 0, s>>>(x, ldx, n_rows, (int)channels, gid, mean, rstd, residual, ldr, act, out, ldo));
  else K(segnorm_apply_kernel<false><<<grid, kNormThreads, 0, s>>>(x, ldx, n_rows, (int)channels, gid, mean, rstd, residual, ldr, act, out, ldo));
  return check_launch("segnorm_apply");
}

extern "C" int stinet_segnorm_bwd(const float* x, int64_t ldx, const float* dout, int64_t ldg, int64_t n_rows,
                                  int64_t channels, int64_t n_seg, int64_t max_seg_rows, const int32_t* slice_ptr,
                                  const float* cnt, const int32_t* gid, const float* mean, const float* rstd, int act,
                                  float* dx, int64_t lddx, void* workspace, size_t workspace_bytes,
                                  stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && dout && dx && ((mean == nullptr) == (rstd == nullptr)), STINET_ERR_ARG, "segnorm_bwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && ldx >= channels && ldg >= channels && lddx >= channels,
                 STINET_ERR_ARG, "segnorm_bwd: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool vec = nvec(channels, {x, dout, dx}, {ldx, ldg, lddx});
  const float *s1 = nullptr, *s2 = nullptr;
  if (mean) {
    STINET_REQUIRE(slice_ptr && cnt && n_seg > 0 && n_seg <= 65535, STINET_ERR_ARG, "segnorm_bwd: segments required");
    NormWs w = carve_norm(workspace, max_seg_rows, channels, n_seg);
    STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "segnorm_bwd: workspace %zu < %zu",
                   workspace_bytes, w.bytes);
    NormArgs a{x, ldx, dout, ldg, slice_ptr, gid, mean, rstd, w.part0, w.part1, (int)channels, w.max_chunks, act};
    launch_colreduce<MODE_BWD>(vec, a, n_seg, s);
    const dim3 fin_grid((unsigned)ceil_div(channels, 32), (unsigned)n_seg);
    K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.part0, slice_ptr, cnt, (int)n_seg, (int)channels, w.max_chunks, 0.f, w.s1));
    K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.part1, slice_ptr, cnt, (int)n_seg, (int)channels, w.max_chunks, 0.f, w.s2));
    s1 = w.s1;
    s2 = w.s2;
  }
  const int grid = wave_grid(n_rows * (channels / (vec ? 4 : 1)), kNormThreads * 4, 8, 8);
  if (vec) K(segnorm_bwd_apply_kernel<true><<<grid, kNormThreads, 0, s>>>(x, ldx, dout, ldg, n_rows, (int)channels, gid, mean, rstd, s1, s2, act, dx, lddx));
  else K(segnorm_bwd_apply_kernel<false><<<grid, kNormThreads, 0, s>>>(x, ldx, dout, ldg, n_rows, (int)channels, gid, mean, rstd, s1, s2, act, dx, lddx));
  return check_launch("segnorm_bwd");
}
