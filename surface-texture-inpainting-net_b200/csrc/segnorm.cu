// Per-graph instance norm (FastInstanceNorm semantics) fused with the ELU + residual tail of GraphResnetBlock.
//   stats : two deterministic segmented column reductions (sum, then centred sum of squares -- the reference is
//           two-pass as well), each = per-chunk partials + a fixed-order finalize;
//   apply : out = residual + act((x - mean[g]) * rstd[g])            (one read of x, one write)
//   bwd   : one reduction pass (sum dz, sum dz*yhat) + one apply pass.
// Slices that are true segments (gid == NULL) get two faster forms:
//   * short slices: ONE kernel per direction.  A thread-block cluster owns a (slice, 32-channel slab) pair, its CTAs
//     split the rows, per-CTA partial sums are exchanged through distributed shared memory in rank order
//     (deterministic), and -- when the CTA's rows fit -- the rows stay in registers between the passes, so x is read
//     from HBM once;
//   * long slices: the same partial/finalize reductions, then slice-indexed apply kernels in which a thread keeps one
//     float4 column (mean / rstd loaded once) and streams rows four at a time.
#include <cooperative_groups.h>
#include "common.cuh"
namespace cg = cooperative_groups;

namespace stinet {

constexpr int kNormThreads = 256;
constexpr int kChunkRows = 256;   // rows per partial
constexpr int kTileGroups = 32;   // channel groups (float4 or scalar) per CTA tile

enum { MODE_SUM = 0, MODE_CSQ = 1, MODE_BWD = 2 };

// Optional second form of a forward result: fp16 operand planes for the dense layer that reads it next (include/
// stinet_b200.h, "dense layers on operand PLANES"), written by the same pass that writes the fp32 rows.  The scale must be
// known before the first row is written, so it comes from an upper bound:  |residual + act(yhat)| <= max|residual| + add,
// add = sqrt(longest slice) (|yhat| <= sqrt(n) for a biased-variance norm over n rows; ELU keeps |.| below max(1, yhat)).
struct PlaneOut {
  __half* hi; __half* lo; int64_t ldp;     // hi == nullptr: no planes
  const float* res_amax;                    // nullable: max|residual| (exact or a bound); no residual: nullptr
  float add;
  int32_t* exp_out;
};
__device__ __forceinline__ float plane_out_scale(const PlaneOut& po, bool writer) {
  if (po.hi == nullptr) return 0.f;
  const float bound = (po.res_amax ? __ldg(po.res_amax) : 0.f) + po.add;
  const int sft = plane_shift(__float_as_uint(bound));
  if (writer) *po.exp_out = -sft;
  return plane_scale(sft);
}

struct NormArgs {
  const float* x; int64_t ldx;
  const float* dout; int64_t ldg;
  const int32_t* slice_ptr; const int32_t* gid;
  const float* mean; const float* rstd;
  float* part0; float* part1;
  int channels; int max_chunks; int act;
  unsigned* aux_amax;                  // MODE_BWD, nullable: max |dout| as a bit pattern (atomicMax into a zeroed slot)
};

template <int MODE, bool VEC>
__global__ void __launch_bounds__(kNormThreads) seg_colreduce_kernel(NormArgs a) {
  constexpr int W = VEC ? 4 : 1;
  constexpr int U = 4;                               // rows in flight per thread
  const int groups = a.channels / W;
  const int txw = min(groups, kTileGroups);         // threads along channels
  const int tyn = kNormThreads / txw;                // row lanes
  const int tx = threadIdx.x % txw, ty = threadIdx.x / txw;
  const int s = blockIdx.y;
  const int grp = blockIdx.z * kTileGroups + tx;
  const int r0 = a.slice_ptr[s] + blockIdx.x * kChunkRows;
  const int r1 = min(r0 + kChunkRows, a.slice_ptr[s + 1]);
  float acc0[W], acc1[W];
  unsigned dmax = 0u;
#pragma unroll
  for (int w = 0; w < W; ++w) acc0[w] = acc1[w] = 0.f;
  if (grp < groups && ty < tyn) {
    // statistics of the slice itself (gid == NULL): loaded once per thread instead of once per row
    float ms[W], rss[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      ms[w] = (MODE != MODE_SUM && !a.gid && a.mean) ? a.mean[(int64_t)s * a.channels + grp * W + w] : 0.f;
      rss[w] = (MODE == MODE_BWD && !a.gid && a.rstd) ? a.rstd[(int64_t)s * a.channels + grp * W + w] : 1.f;
    }
    for (int rb = r0 + ty; rb < r1; rb += U * tyn) {
      float v[U][W], d[U][W];
      int g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = rb + u * tyn;
        const bool ok = r < r1;
        g[u] = (MODE != MODE_SUM && a.gid && ok) ? a.gid[r] : s;
        if (VEC) {
          float4 t = ok ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[grp] : make_float4(0.f, 0.f, 0.f, 0.f);
          v[u][0] = t.x; v[u][1 % W] = t.y; v[u][2 % W] = t.z; v[u][3 % W] = t.w;
        } else {
          v[u][0] = ok ? a.x[(int64_t)r * a.ldx + grp] : 0.f;
        }
        if (MODE == MODE_BWD) {
          if (VEC) {
            float4 t = ok ? reinterpret_cast<const float4*>(a.dout + (int64_t)r * a.ldg)[grp] : make_float4(0.f, 0.f, 0.f, 0.f);
            d[u][0] = t.x; d[u][1 % W] = t.y; d[u][2 % W] = t.z; d[u][3 % W] = t.w;
          } else {
            d[u][0] = ok ? a.dout[(int64_t)r * a.ldg + grp] : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (rb + u * tyn < r1) {                     // rows are accumulated in ascending order, as before
#pragma unroll
          for (int w = 0; w < W; ++w) {
            float m = ms[w], rs = rss[w];
            if (MODE != MODE_SUM && a.gid && a.mean) m = a.mean[(int64_t)g[u] * a.channels + grp * W + w];
            if (MODE == MODE_BWD && a.gid && a.rstd) rs = a.rstd[(int64_t)g[u] * a.channels + grp * W + w];
            if (MODE == MODE_SUM) {
              acc0[w] += v[u][w];
            } else if (MODE == MODE_CSQ) {
              const float c = v[u][w] - m;
              acc0[w] += c * c;
            } else {
              const float yh = (v[u][w] - m) * rs;
              const float dz = (a.act == STINET_ACT_ELU) ? d[u][w] * elu1_grad(yh) : d[u][w];
              dmax = max(dmax, __float_as_uint(d[u][w]) & 0x7FFFFFFFu);
              acc0[w] += dz;
              acc1[w] += dz * yh;
            }
          }
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
  if (MODE == MODE_BWD && a.aux_amax != nullptr) amax_publish(dmax, a.aux_amax);
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

// ---------------------------------------------------------------------------------------------------------------
// slices == segments: slice-indexed apply kernels (long slices) and the single-kernel cluster forms (short slices)

constexpr int kApplyU = 4;  // rows in flight per thread

// grid (row chunks, slices, 32-group slabs); a thread owns one float4 column and streams rows kApplyU at a time.
//   BWD = false: out = res + act((x - mean[s]) * rstd[s])                                   (aux = residual, nullable)
//   BWD = true : dx  = rstd[s] * (dz - s1[s] - yhat * s2[s]),  dz = dout * act'(yhat)      (aux = dout)
template <bool BWD>
__global__ void __launch_bounds__(kNormThreads)
segnorm_slice_apply_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ aux, int64_t lda,
                           const int32_t* __restrict__ slice_ptr, int channels, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ s1, const float* __restrict__ s2,
                           int act, float* __restrict__ out, int64_t ldo, unsigned* __restrict__ amax_out, PlaneOut po) {
  const float pscale = plane_out_scale(po, blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0);
  const int groups = channels >> 2;
  const int txw = min(groups, kTileGroups);
  const int tyn = kNormThreads / txw;
  const int tx = threadIdx.x % txw, ty = threadIdx.x / txw;
  const int s = blockIdx.y;
  const int grp = blockIdx.z * kTileGroups + tx;
  unsigned amax_bits = 0u;
  if (grp < groups && ty < tyn) {
  const int rows_per_cta = tyn * kApplyU * 2;
  const int r0 = slice_ptr[s] + blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, slice_ptr[s + 1]);
  const int64_t sc = (int64_t)s * channels + grp * 4;
  const float4 m = *reinterpret_cast<const float4*>(mean + sc);
  const float4 rs = *reinterpret_cast<const float4*>(rstd + sc);
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  if (BWD) {
    a1 = *reinterpret_cast<const float4*>(s1 + sc);
    a2 = *reinterpret_cast<const float4*>(s2 + sc);
  }
  const float mv[4] = {m.x, m.y, m.z, m.w}, rv[4] = {rs.x, rs.y, rs.z, rs.w};
  const float p1[4] = {a1.x, a1.y, a1.z, a1.w}, p2[4] = {a2.x, a2.y, a2.z, a2.w};
  for (int rb = r0 + ty; rb < r1; rb += kApplyU * tyn) {
    float4 v[kApplyU], w[kApplyU];
#pragma unroll
    for (int u = 0; u < kApplyU; ++u) {
      const int r = rb + u * tyn;
      v[u] = w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < r1) {
        v[u] = ld_stream(reinterpret_cast<const float4*>(x + (int64_t)r * ldx) + grp);
        if (aux) w[u] = ld_stream(reinterpret_cast<const float4*>(aux + (int64_t)r * lda) + grp);
      }
    }
#pragma unroll
    for (int u = 0; u < kApplyU; ++u) {
      const int r = rb + u * tyn;
      if (r < r1) {
        const float xv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const float wv[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float yh = (xv[q] - mv[q]) * rv[q];
          if (BWD) {
            const float dz = (act == STINET_ACT_ELU) ? wv[q] * elu1_grad(yh) : wv[q];
            o[q] = rv[q] * (dz - p1[q] - yh * p2[q]);
          } else {
            o[q] = (act == STINET_ACT_ELU) ? elu1(yh) : yh;
            if (aux) o[q] += wv[q];
          }
        }
        const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
        amax_bits = amax4(amax_bits, ov);
        st_stream(reinterpret_cast<float4*>(out + (int64_t)r * ldo) + grp, ov);
        if (!BWD && po.hi != nullptr)
          split_store4(ov, pscale, po.hi + (int64_t)r * po.ldp + 4 * grp, po.lo ? po.lo + (int64_t)r * po.ldp + 4 * grp : nullptr);
      }
    }
  }
  }
  if (amax_out != nullptr) amax_publish(amax_bits, amax_out);   // max |out| for the plane scale of the next dense layer
}

constexpr int kFusedG = 8;        // float4 columns per CTA: a 32-channel slab = one 128-byte line per row
constexpr int kFusedLanes = 32;   // row lanes per CTA (256 threads)
constexpr int kFusedRR = 8;       // rows a thread can keep in registers -> 256 rows per CTA
constexpr int kFusedMaxCluster = 8;
static_assert(kFusedG * kFusedLanes == kNormThreads, "fused norm kernels: 8 columns x 32 row lanes");

struct FusedArgs {
  const float* x; int64_t ldx;
  const float* aux; int64_t lda;       // fwd: residual (nullable); bwd: dout
  const int32_t* slice_ptr; const float* cnt;
  float* mean; float* rstd;            // fwd: outputs; bwd: inputs
  float* out; int64_t ldo;
  int channels; int act; float eps;
  unsigned* amax_out;                  // nullable: max |out| as a bit pattern (atomicMax into a zeroed slot)
  PlaneOut po;                         // fwd: optional fp16 planes of out
  unsigned* aux_amax;                  // bwd, nullable: max |dout| as a bit pattern
};

__device__ __forceinline__ float4 f4add(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Sum of `v` over the row lanes of this CTA and then over the CTAs of the cluster, both in a fixed order.
// `sm` is a 256-entry scratch, `part` a per-pass 8-entry exchange buffer (never reused, so one cluster.sync suffices).
__device__ __forceinline__ float4 cluster_column_sum(float4 v, float4* sm, float4* part, cg::cluster_group& cluster) {
  const int tx = threadIdx.x & (kFusedG - 1), ty = threadIdx.x >> 3;
  __syncthreads();                      // previous users of sm are done
  sm[threadIdx.x] = v;
  __syncthreads();
#pragma unroll
  for (int off = kFusedLanes / 2; off >= 1; off >>= 1) {
    if (ty < off) sm[threadIdx.x] = f4add(sm[threadIdx.x], sm[threadIdx.x + off * kFusedG]);
    __syncthreads();
  }
  if (ty == 0) part[tx] = sm[tx];
  cluster.sync();
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  const unsigned cs = cluster.num_blocks();
  for (unsigned k = 0; k < cs; ++k) t = f4add(t, cluster.map_shared_rank(part, k)[tx]);
  return t;
}

// RR > 0: the CTA's rows (<= 32*RR) are loaded once and kept in registers; RR == 0: every pass re-reads them (L2).
template <int RR>
__global__ void __launch_bounds__(kNormThreads) segnorm_fused_fwd_kernel(FusedArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float4 sm[kNormThreads];
  __shared__ float4 part_a[kFusedG], part_b[kFusedG];
  const int cs = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tx = threadIdx.x & (kFusedG - 1), ty = threadIdx.x >> 3;
  const int s = blockIdx.y;
  const int c4 = (blockIdx.x / cs) * kFusedG + tx;
  const bool col = c4 < (a.channels >> 2);
  const int b = a.slice_ptr[s], e = a.slice_ptr[s + 1];
  const int per = (e - b + cs - 1) / cs;
  const int r0 = b + rank * per, r1 = min(r0 + per, e);
  const float cnt = a.cnt[s];
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int NR = RR > 0 ? RR : 1;
  float4 v[NR];
  float4 acc = zero;
  if (RR > 0) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = r0 + ty + i * kFusedLanes;
      v[i] = (col && r < r1) ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[c4] : zero;
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) acc = f4add(acc, v[i]);
  } else if (col) {
    for (int rb = r0 + ty; rb < r1; rb += 4 * kFusedLanes) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * kFusedLanes;
        t[u] = r < r1 ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[c4] : zero;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = f4add(acc, t[u]);
    }
  }
  float4 tot = cluster_column_sum(acc, sm, part_a, cluster);
  const float4 m = make_float4(tot.x / cnt, tot.y / cnt, tot.z / cnt, tot.w / cnt);
  acc = zero;
  if (RR > 0) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = r0 + ty + i * kFusedLanes;
      if (col && r < r1) {
        const float dx = v[i].x - m.x, dy = v[i].y - m.y, dz = v[i].z - m.z, dw = v[i].w - m.w;
        acc.x += dx * dx; acc.y += dy * dy; acc.z += dz * dz; acc.w += dw * dw;
      }
    }
  } else if (col) {
    for (int rb = r0 + ty; rb < r1; rb += 4 * kFusedLanes) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * kFusedLanes;
        t[u] = r < r1 ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[c4] : m;   // (m - m)^2 adds exactly 0
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float dx = t[u].x - m.x, dy = t[u].y - m.y, dz = t[u].z - m.z, dw = t[u].w - m.w;
        acc.x += dx * dx; acc.y += dy * dy; acc.z += dz * dz; acc.w += dw * dw;
      }
    }
  }
  tot = cluster_column_sum(acc, sm, part_b, cluster);
  const float4 rs = make_float4(1.f / sqrtf(tot.x / cnt + a.eps), 1.f / sqrtf(tot.y / cnt + a.eps),
                                1.f / sqrtf(tot.z / cnt + a.eps), 1.f / sqrtf(tot.w / cnt + a.eps));
  if (rank == 0 && ty == 0 && col) {
    reinterpret_cast<float4*>(a.mean + (int64_t)s * a.channels)[c4] = m;
    reinterpret_cast<float4*>(a.rstd + (int64_t)s * a.channels)[c4] = rs;
  }
  unsigned amax_bits = 0u;
  const float pscale = plane_out_scale(a.po, blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
  auto finish = [&](const float4& xv, int r) {
    float4 o = make_float4((xv.x - m.x) * rs.x, (xv.y - m.y) * rs.y, (xv.z - m.z) * rs.z, (xv.w - m.w) * rs.w);
    if (a.act == STINET_ACT_ELU) { o.x = elu1(o.x); o.y = elu1(o.y); o.z = elu1(o.z); o.w = elu1(o.w); }
    if (a.aux) o = f4add(o, reinterpret_cast<const float4*>(a.aux + (int64_t)r * a.lda)[c4]);
    amax_bits = amax4(amax_bits, o);
    reinterpret_cast<float4*>(a.out + (int64_t)r * a.ldo)[c4] = o;
    if (a.po.hi != nullptr)
      split_store4(o, pscale, a.po.hi + (int64_t)r * a.po.ldp + 4 * c4, a.po.lo ? a.po.lo + (int64_t)r * a.po.ldp + 4 * c4 : nullptr);
  };
  if (RR > 0) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = r0 + ty + i * kFusedLanes;
      if (col && r < r1) finish(v[i], r);
    }
  } else if (col) {
    for (int rb = r0 + ty; rb < r1; rb += 4 * kFusedLanes) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * kFusedLanes;
        t[u] = r < r1 ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[c4] : zero;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * kFusedLanes;
        if (r < r1) finish(t[u], r);
      }
    }
  }
  if (a.amax_out != nullptr) amax_publish(amax_bits, a.amax_out);
  cluster.sync();   // no CTA may exit while a peer can still read its exchange buffers
}

template <int RR>
__global__ void __launch_bounds__(kNormThreads) segnorm_fused_bwd_kernel(FusedArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float4 sm[kNormThreads];
  __shared__ float4 part_a[kFusedG], part_b[kFusedG];
  const int cs = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tx = threadIdx.x & (kFusedG - 1), ty = threadIdx.x >> 3;
  const int s = blockIdx.y;
  const int c4 = (blockIdx.x / cs) * kFusedG + tx;
  const bool col = c4 < (a.channels >> 2);
  const int b = a.slice_ptr[s], e = a.slice_ptr[s + 1];
  const int per = (e - b + cs - 1) / cs;
  const int r0 = b + rank * per, r1 = min(r0 + per, e);
  const float cnt = a.cnt[s];
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 m = col ? reinterpret_cast<const float4*>(a.mean + (int64_t)s * a.channels)[c4] : zero;
  const float4 rs = col ? reinterpret_cast<const float4*>(a.rstd + (int64_t)s * a.channels)[c4] : zero;
  const bool elu = a.act == STINET_ACT_ELU;
  // (yhat, dz) of one element quad
  unsigned dmax = 0u;
  auto prep = [&](const float4& xv, const float4& dv, float4& yh, float4& dz) {
    yh = make_float4((xv.x - m.x) * rs.x, (xv.y - m.y) * rs.y, (xv.z - m.z) * rs.z, (xv.w - m.w) * rs.w);
    dz = dv;
    if (elu) { dz.x *= elu1_grad(yh.x); dz.y *= elu1_grad(yh.y); dz.z *= elu1_grad(yh.z); dz.w *= elu1_grad(yh.w); }
  };
  constexpr int NR = RR > 0 ? RR : 1;
  float4 yhc[NR], dzc[NR];
  float4 acc0 = zero, acc1 = zero;
  if (RR > 0) {
    float4 xv[NR], dv[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = r0 + ty + i * kFusedLanes;
      const bool ok = col && r < r1;
      xv[i] = ok ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[c4] : m;
      dv[i] = ok ? reinterpret_cast<const float4*>(a.aux + (int64_t)r * a.lda)[c4] : zero;
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      prep(xv[i], dv[i], yhc[i], dzc[i]);     // padded rows: yhat = 0, dz = 0 -> contribute exactly 0
      dmax = amax4(dmax, dv[i]);
      acc0 = f4add(acc0, dzc[i]);
      acc1.x += dzc[i].x * yhc[i].x; acc1.y += dzc[i].y * yhc[i].y; acc1.z += dzc[i].z * yhc[i].z; acc1.w += dzc[i].w * yhc[i].w;
    }
  } else if (col) {
    for (int rb = r0 + ty; rb < r1; rb += 4 * kFusedLanes) {
      float4 xv[4], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * kFusedLanes;
        xv[u] = r < r1 ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[c4] : m;
        dv[u] = r < r1 ? reinterpret_cast<const float4*>(a.aux + (int64_t)r * a.lda)[c4] : zero;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 yh, dz;
        prep(xv[u], dv[u], yh, dz);
        dmax = amax4(dmax, dv[u]);
        acc0 = f4add(acc0, dz);
        acc1.x += dz.x * yh.x; acc1.y += dz.y * yh.y; acc1.z += dz.z * yh.z; acc1.w += dz.w * yh.w;
      }
    }
  }
  const float4 t0 = cluster_column_sum(acc0, sm, part_a, cluster);
  const float4 t1 = cluster_column_sum(acc1, sm, part_b, cluster);
  const float4 s1 = make_float4(t0.x / cnt, t0.y / cnt, t0.z / cnt, t0.w / cnt);
  const float4 s2 = make_float4(t1.x / cnt, t1.y / cnt, t1.z / cnt, t1.w / cnt);
  unsigned amax_bits = 0u;
  auto finish = [&](const float4& yh, const float4& dz, int r) {
    const float4 o = make_float4(rs.x * (dz.x - s1.x - yh.x * s2.x), rs.y * (dz.y - s1.y - yh.y * s2.y),
                                 rs.z * (dz.z - s1.z - yh.z * s2.z), rs.w * (dz.w - s1.w - yh.w * s2.w));
    amax_bits = amax4(amax_bits, o);
    reinterpret_cast<float4*>(a.out + (int64_t)r * a.ldo)[c4] = o;
  };
  if (RR > 0) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = r0 + ty + i * kFusedLanes;
      if (col && r < r1) finish(yhc[i], dzc[i], r);
    }
  } else if (col) {
    for (int rb = r0 + ty; rb < r1; rb += 4 * kFusedLanes) {
      float4 xv[4], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * kFusedLanes;
        xv[u] = r < r1 ? reinterpret_cast<const float4*>(a.x + (int64_t)r * a.ldx)[c4] : m;
        dv[u] = r < r1 ? reinterpret_cast<const float4*>(a.aux + (int64_t)r * a.lda)[c4] : zero;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * kFusedLanes;
        if (r < r1) {
          float4 yh, dz;
          prep(xv[u], dv[u], yh, dz);
          finish(yh, dz, r);
        }
      }
    }
  }
  if (a.amax_out != nullptr) amax_publish(amax_bits, a.amax_out);
  if (a.aux_amax != nullptr) amax_publish(dmax, a.aux_amax);
  cluster.sync();
}

// Launch geometry of the single-kernel forms: the smallest cluster (1, 2, 4, 8 CTAs) whose CTAs can keep their rows
// in registers, the largest one otherwise.  Slices longer than kFusedMaxRows use the multi-kernel path.
constexpr int64_t kFusedMaxRows = 16384;
struct FusedPlan { int cluster; bool cached; };
static FusedPlan fused_plan(int64_t max_seg_rows) {
  int cs = 1;
  while (cs < kFusedMaxCluster && ceil_div(max_seg_rows, cs) > kFusedRR * kFusedLanes) cs *= 2;
  return FusedPlan{cs, ceil_div(max_seg_rows, cs) <= kFusedRR * kFusedLanes};
}

template <typename Kern>
static int launch_fused(Kern kern, const FusedArgs& a, int64_t n_seg, int cluster, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  const unsigned slabs = (unsigned)ceil_div(a.channels >> 2, kFusedG);
  cfg.gridDim = dim3(slabs * cluster, (unsigned)n_seg, 1);
  cfg.blockDim = dim3(kNormThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  count_launch();
  if (e != cudaSuccess) {
    set_error("segnorm fused launch: %s", cudaGetErrorString(e));
    return STINET_ERR_CUDA;
  }
  return STINET_OK;
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

// ---------------------------------------------------------------------------------------------------------------
// affine segmented norms: BatchNorm over the node rows (BatchNorm2Param, nn.BatchNorm1d inside the
// with_norm message MLP) and SingleBatchGraphNorm share one form
//     y = gamma * (x - alpha * m[g]) * r[g] + beta,   r = 1 / sqrt(v + eps)
// with m = slice mean, and v = slice mean of (x - m)^2 (kind 0, batch norm) or of x^2 (kind 1: the reference's GraphNorm
// takes the second moment of the UN-shifted x, singlebatchgroupnorm.py:66-68).  Statistics come from the deterministic
// column reductions above; the kernels below are the elementwise apply, its backward and the parameter gradients.

template <bool VEC>
__global__ void __launch_bounds__(kNormThreads)
affnorm_apply_kernel(const float* __restrict__ x, int64_t ldx, int64_t n_rows, int channels, const int32_t* __restrict__ gid,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ alpha,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out, int64_t ldo) {
  constexpr int W = VEC ? 4 : 1;
  const int groups = channels / W;
  const int64_t total = n_rows * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / groups;
    const int grp = (int)(idx - r * groups);
    const int g = gid ? gid[r] : 0;
    float v[W], o[W];
    if (VEC) {
      const float4 t = reinterpret_cast<const float4*>(x + r * ldx)[grp];
      v[0] = t.x; v[1 % W] = t.y; v[2 % W] = t.z; v[3 % W] = t.w;
    } else {
      v[0] = x[r * ldx + grp];
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int c = grp * W + w;
      const int64_t sc = (int64_t)g * channels + c;
      const float u = v[w] - (alpha ? alpha[c] : 1.f) * mean[sc];
      o[w] = (gamma ? gamma[c] : 1.f) * u * rstd[sc] + (beta ? beta[c] : 0.f);
    }
    if (VEC) reinterpret_cast<float4*>(out + r * ldo)[grp] = make_float4(o[0], o[1 % W], o[2 % W], o[3 % W]);
    else out[r * ldo + grp] = o[0];
  }
}

// am[s,c] = alpha[c] * mean[s,c] (the shift the reductions of the backward centre x with); ones[s] = 1
__global__ void affnorm_prep_kernel(const float* __restrict__ mean, const float* __restrict__ alpha, int n_seg, int channels,
                                    float* __restrict__ am, float* __restrict__ ones) {
  const int64_t total = (int64_t)n_seg * channels;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x)
    am[idx] = (alpha ? alpha[idx % channels] : 1.f) * mean[idx];
  if (blockIdx.x == 0)
    for (int sidx = threadIdx.x; sidx < n_seg; sidx += blockDim.x) ones[sidx] = 1.f;
}

// dx_i = r gamma dy_i - alpha r gamma T1 / n - r^3 gamma T2 w_i / n,  w_i = x_i - m (kind 0) or x_i (kind 1),
// T1 = sum_slice dy, T2 = sum_slice dy (x - alpha m), n = the forward's divisor of the slice
template <bool VEC>
__global__ void __launch_bounds__(kNormThreads)
affnorm_bwd_apply_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t ldg, int64_t n_rows,
                         int channels, const int32_t* __restrict__ gid, const float* __restrict__ cnt, int kind,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ alpha,
                         const float* __restrict__ gamma, const float* __restrict__ t1, const float* __restrict__ t2,
                         float* __restrict__ dx, int64_t lddx) {
  constexpr int W = VEC ? 4 : 1;
  const int groups = channels / W;
  const int64_t total = n_rows * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / groups;
    const int grp = (int)(idx - r * groups);
    const int g = gid ? gid[r] : 0;
    const float n = cnt[g];
    float v[W], d[W], o[W];
    if (VEC) {
      const float4 t = reinterpret_cast<const float4*>(x + r * ldx)[grp];
      v[0] = t.x; v[1 % W] = t.y; v[2 % W] = t.z; v[3 % W] = t.w;
      const float4 q = reinterpret_cast<const float4*>(dy + r * ldg)[grp];
      d[0] = q.x; d[1 % W] = q.y; d[2 % W] = q.z; d[3 % W] = q.w;
    } else {
      v[0] = x[r * ldx + grp];
      d[0] = dy[r * ldg + grp];
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int c = grp * W + w;
      const int64_t sc = (int64_t)g * channels + c;
      const float rs = rstd[sc], gm = gamma ? gamma[c] : 1.f, al = alpha ? alpha[c] : 1.f;
      const float wi = kind == 0 ? v[w] - mean[sc] : v[w];
      o[w] = rs * gm * (d[w] - al * t1[sc] / n - rs * rs * t2[sc] * wi / n);
    }
    if (VEC) reinterpret_cast<float4*>(dx + r * lddx)[grp] = make_float4(o[0], o[1 % W], o[2 % W], o[3 % W]);
    else dx[r * lddx + grp] = o[0];
  }
}

// dgamma[c] = sum_s r T2, dbeta[c] = sum_s T1, dalpha[c] = -gamma sum_s m r T1   (segments in ascending order)
__global__ void affnorm_param_grads_kernel(const float* __restrict__ mean, const float* __restrict__ rstd,
                                           const float* __restrict__ gamma, const float* __restrict__ t1,
                                           const float* __restrict__ t2, int n_seg, int channels, float* __restrict__ dgamma,
                                           float* __restrict__ dbeta, float* __restrict__ dalpha) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  float g = 0.f, b = 0.f, a = 0.f;
  for (int sidx = 0; sidx < n_seg; ++sidx) {
    const int64_t sc = (int64_t)sidx * channels + c;
    g += rstd[sc] * t2[sc];
    b += t1[sc];
    a += mean[sc] * rstd[sc] * t1[sc];
  }
  if (dgamma) dgamma[c] = g;
  if (dbeta) dbeta[c] = b;
  if (dalpha) dalpha[c] = -(gamma ? gamma[c] : 1.f) * a;
}

// running statistics of nn.BatchNorm1d after one training step: mean as is, variance unbiased (n / (n - 1))
__global__ void bn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, float n, float eps,
                                         float momentum, int channels, float* __restrict__ running_mean,
                                         float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  const float rs = rstd[c];
  const float var = 1.f / (rs * rs) - eps;
  const float unbiased = n > 1.f ? var * (n / (n - 1.f)) : var;
  running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
  running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
}

struct AffWs { NormWs n; float* am; float* ones; size_t bytes; };
static AffWs carve_aff(void* base, int64_t max_seg_rows, int64_t channels, int64_t n_seg) {
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  AffWs w;
  w.n = carve_norm(base, max_seg_rows, channels, n_seg);
  char* p = static_cast<char*>(base) + up(w.n.bytes);
  w.am = reinterpret_cast<float*>(p);
  w.ones = reinterpret_cast<float*>(p + up(sizeof(float) * (size_t)n_seg * channels));
  w.bytes = up(w.n.bytes) + up(sizeof(float) * (size_t)n_seg * channels) + up(sizeof(float) * (size_t)n_seg);
  return w;
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

extern "C" int stinet_segnorm_fwd(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, int64_t n_seg,
                                  int64_t max_seg_rows, const int32_t* slice_ptr, const float* cnt, float eps,
                                  const float* residual, int64_t ldr, int act, float* out, int64_t ldo, float* mean,
                                  float* rstd, float* amax_out, void* out_hi, void* out_lo, int64_t ldp,
                                  const float* res_amax, int32_t* out_exp, void* workspace, size_t workspace_bytes,
                                  stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  unsigned* am = reinterpret_cast<unsigned*>(amax_out);
  const float nrm = sqrtf((float)(max_seg_rows > 1 ? max_seg_rows : 1));
  PlaneOut po{static_cast<__half*>(out_hi), static_cast<__half*>(out_lo), ldp, res_amax,
              act == STINET_ACT_ELU ? fmaxf(1.f, nrm) : nrm, out_exp};
  if (out_hi != nullptr) {
    STINET_REQUIRE(out_exp && ldp >= channels && ldp % 8 == 0 && aligned16(out_hi) && (!out_lo || aligned16(out_lo)) &&
                       (residual == nullptr || res_amax != nullptr),
                   STINET_ERR_ARG, "segnorm_fwd: plane output needs exp, a pitch %% 8 == 0 and max|residual|");
  }
  if (am != nullptr) {
    cudaError_t e = cudaMemsetAsync(am, 0, sizeof(unsigned), s);
    STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "segnorm_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
  }
  STINET_REQUIRE(x && slice_ptr && cnt && mean && rstd && out, STINET_ERR_ARG, "segnorm_fwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && n_seg > 0 && n_seg <= 65535 && ldx >= channels && ldo >= channels &&
                     (!residual || ldr >= channels),
                 STINET_ERR_ARG, "segnorm_fwd: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool vec = nvec(channels, {x, out, residual, mean, rstd}, {ldx, ldo, residual ? ldr : 0});
  if (vec && max_seg_rows <= kFusedMaxRows) {
    const FusedPlan pl = fused_plan(max_seg_rows);
    FusedArgs a{x, ldx, residual, ldr, slice_ptr, cnt, mean, rstd, out, ldo, (int)channels, act, eps, am, po};
    return pl.cached ? launch_fused(segnorm_fused_fwd_kernel<kFusedRR>, a, n_seg, pl.cluster, s)
                     : launch_fused(segnorm_fused_fwd_kernel<0>, a, n_seg, pl.cluster, s);
  }
  int rc = stinet_segnorm_stats(x, ldx, n_rows, channels, n_seg, max_seg_rows, slice_ptr, cnt, nullptr, eps, mean, rstd,
                                workspace, workspace_bytes, stream_);
  if (rc != STINET_OK) return rc;
  if (vec) {
    const int groups = (int)(channels >> 2);
    const int tyn = kNormThreads / (groups < kTileGroups ? groups : kTileGroups);
    dim3 grid((unsigned)ceil_div(max_seg_rows, tyn * kApplyU * 2), (unsigned)n_seg, (unsigned)ceil_div(groups, kTileGroups));
    K(segnorm_slice_apply_kernel<false><<<grid, kNormThreads, 0, s>>>(x, ldx, residual, ldr, slice_ptr, (int)channels, mean,
                                                                      rstd, nullptr, nullptr, act, out, ldo, am, po));
    return check_launch("segnorm_fwd");
  }
  STINET_REQUIRE(out_hi == nullptr, STINET_ERR_UNSUPPORTED, "segnorm_fwd: plane output needs the vector path (channels %% 4 == 0, aligned rows)");
  // odd widths: the per-row lookup kernel needs a graph id; a single slice uses id 0
  STINET_REQUIRE(n_seg == 1, STINET_ERR_UNSUPPORTED, "segnorm_fwd: channel count %lld (not a multiple of 4) with %lld slices",
                 (long long)channels, (long long)n_seg);
  rc = stinet_segnorm_apply(x, ldx, n_rows, channels, nullptr, mean, rstd, residual, ldr, act, out, ldo, stream_);
  if (rc == STINET_OK && amax_out != nullptr) rc = stinet_f16_amax(out, ldo, n_rows, channels, amax_out, stream_);
  return rc;
}

extern "C" int stinet_segnorm_bwd(const float* x, int64_t ldx, const float* dout, int64_t ldg, int64_t n_rows,
                                  int64_t channels, int64_t n_seg, int64_t max_seg_rows, const int32_t* slice_ptr,
                                  const float* cnt, const int32_t* gid, const float* mean, const float* rstd, int act,
                                  float* dx, int64_t lddx, float* amax_out, float* dout_amax, void* workspace,
                                  size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  unsigned* am = reinterpret_cast<unsigned*>(amax_out);
  unsigned* dam = reinterpret_cast<unsigned*>(dout_amax);
  for (unsigned* slot : {am, dam})
    if (slot != nullptr) {
      cudaError_t e = cudaMemsetAsync(slot, 0, sizeof(unsigned), s);
      STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "segnorm_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
  STINET_REQUIRE(x && dout && dx && ((mean == nullptr) == (rstd == nullptr)), STINET_ERR_ARG, "segnorm_bwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && ldx >= channels && ldg >= channels && lddx >= channels,
                 STINET_ERR_ARG, "segnorm_bwd: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool vec = nvec(channels, {x, dout, dx, mean, rstd}, {ldx, ldg, lddx});
  const float *s1 = nullptr, *s2 = nullptr;
  const bool by_slice = mean && !gid && vec;          // slices are the segments
  if (mean && !gid && !vec)
    STINET_REQUIRE(n_seg == 1, STINET_ERR_UNSUPPORTED, "segnorm_bwd: odd channel count with several slices needs gid");
  if (by_slice && max_seg_rows <= kFusedMaxRows) {
    STINET_REQUIRE(slice_ptr && cnt && n_seg > 0 && n_seg <= 65535, STINET_ERR_ARG, "segnorm_bwd: segments required");
    const FusedPlan pl = fused_plan(max_seg_rows);
    FusedArgs a{x, ldx, dout, ldg, slice_ptr, cnt, const_cast<float*>(mean), const_cast<float*>(rstd), dx, lddx,
                (int)channels, act, 0.f, am, PlaneOut{nullptr, nullptr, 0, nullptr, 0.f, nullptr}, dam};
    return pl.cached ? launch_fused(segnorm_fused_bwd_kernel<kFusedRR>, a, n_seg, pl.cluster, s)
                     : launch_fused(segnorm_fused_bwd_kernel<0>, a, n_seg, pl.cluster, s);
  }
  if (mean) {
    STINET_REQUIRE(slice_ptr && cnt && n_seg > 0 && n_seg <= 65535, STINET_ERR_ARG, "segnorm_bwd: segments required");
    NormWs w = carve_norm(workspace, max_seg_rows, channels, n_seg);
    STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "segnorm_bwd: workspace %zu < %zu",
                   workspace_bytes, w.bytes);
    NormArgs a{x, ldx, dout, ldg, slice_ptr, gid, mean, rstd, w.part0, w.part1, (int)channels, w.max_chunks, act, dam};
    launch_colreduce<MODE_BWD>(vec, a, n_seg, s);
    dam = nullptr;                                      // done by the reduction pass
    const dim3 fin_grid((unsigned)ceil_div(channels, 32), (unsigned)n_seg);
    K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.part0, slice_ptr, cnt, (int)n_seg, (int)channels, w.max_chunks, 0.f, w.s1));
    K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.part1, slice_ptr, cnt, (int)n_seg, (int)channels, w.max_chunks, 0.f, w.s2));
    s1 = w.s1;
    s2 = w.s2;
  }
  if (by_slice) {
    const int groups = (int)(channels >> 2);
    const int tyn = kNormThreads / (groups < kTileGroups ? groups : kTileGroups);
    dim3 grid2((unsigned)ceil_div(max_seg_rows, tyn * kApplyU * 2), (unsigned)n_seg, (unsigned)ceil_div(groups, kTileGroups));
    K(segnorm_slice_apply_kernel<true><<<grid2, kNormThreads, 0, s>>>(x, ldx, dout, ldg, slice_ptr, (int)channels, mean, rstd,
                                                                      s1, s2, act, dx, lddx, am,
                                                                      PlaneOut{nullptr, nullptr, 0, nullptr, 0.f, nullptr}));
    return check_launch("segnorm_bwd");
  }
  const int grid = wave_grid(n_rows * (channels / (vec ? 4 : 1)), kNormThreads * 4, 8, 8);
  if (vec) K(segnorm_bwd_apply_kernel<true><<<grid, kNormThreads, 0, s>>>(x, ldx, dout, ldg, n_rows, (int)channels, gid, mean, rstd, s1, s2, act, dx, lddx));
  else K(segnorm_bwd_apply_kernel<false><<<grid, kNormThreads, 0, s>>>(x, ldx, dout, ldg, n_rows, (int)channels, gid, mean, rstd, s1, s2, act, dx, lddx));
  int rc = check_launch("segnorm_bwd");
  if (rc == STINET_OK && amax_out != nullptr) rc = stinet_f16_amax(dx, lddx, n_rows, channels, amax_out, stream_);
  if (rc == STINET_OK && dam != nullptr) rc = stinet_f16_amax(dout, ldg, n_rows, channels, dout_amax, stream_);  // no reduction pass ran
  return rc;
}

// ---- affine segmented norms ------------------------------------------------------------------------------------

extern "C" size_t stinet_affnorm_workspace_bytes(int64_t max_seg_rows, int64_t channels, int64_t n_seg) {
  if (max_seg_rows < 0 || channels <= 0 || n_seg <= 0) return 0;
  return carve_aff(nullptr, max_seg_rows, channels, n_seg).bytes;
}

extern "C" int stinet_affnorm_apply(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, const int32_t* gid,
                                    const float* mean, const float* rstd, const float* alpha, const float* gamma,
                                    const float* beta, float* out, int64_t ldo, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && mean && rstd && out, STINET_ERR_ARG, "affnorm_apply: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && ldx >= channels && ldo >= channels, STINET_ERR_ARG, "affnorm_apply: bad shape");
  if (n_rows == 0) return STINET_OK;
  const bool vec = nvec(channels, {x, out}, {ldx, ldo});
  const int grid = wave_grid(n_rows * (channels / (vec ? 4 : 1)), kNormThreads * 4, 8, 8);
  if (vec) K(affnorm_apply_kernel<true><<<grid, kNormThreads, 0, s>>>(x, ldx, n_rows, (int)channels, gid, mean, rstd, alpha, gamma, beta, out, ldo));
  else K(affnorm_apply_kernel<false><<<grid, kNormThreads, 0, s>>>(x, ldx, n_rows, (int)channels, gid, mean, rstd, alpha, gamma, beta, out, ldo));
  return check_launch("affnorm_apply");
}

extern "C" int stinet_affnorm_fwd(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, int64_t n_seg,
                                  int64_t max_seg_rows, const int32_t* slice_ptr, const float* cnt, const int32_t* gid,
                                  int kind, float eps, const float* alpha, const float* gamma, const float* beta,
                                  float* out, int64_t ldo, float* mean, float* rstd, void* workspace,
                                  size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && slice_ptr && cnt && out && mean && rstd, STINET_ERR_ARG, "affnorm_fwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && n_seg > 0 && n_seg <= 65535 && ldx >= channels && ldo >= channels &&
                     (kind == 0 || kind == 1) && (n_seg == 1 || gid),
                 STINET_ERR_ARG, "affnorm_fwd: bad shape / kind (several segments need gid)");
  AffWs w = carve_aff(workspace, max_seg_rows, channels, n_seg);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "affnorm_fwd: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  const bool vec = nvec(channels, {x}, {ldx});
  const dim3 fin_grid((unsigned)ceil_div(channels, 32), (unsigned)n_seg);
  NormArgs a{x, ldx, nullptr, 0, slice_ptr, nullptr, nullptr, nullptr, w.n.part0, w.n.part1, (int)channels, w.n.max_chunks, 0};
  launch_colreduce<MODE_SUM>(vec, a, n_seg, s);
  K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.n.part0, slice_ptr, cnt, (int)n_seg, (int)channels, w.n.max_chunks, eps, mean));
  if (kind == 0) {
    a.mean = mean;                                   // centred second moment
  } else {
    cudaError_t e = cudaMemsetAsync(w.am, 0, sizeof(float) * (size_t)n_seg * channels, s);   // second moment of x itself
    STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "affnorm_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
    a.mean = w.am;
  }
  launch_colreduce<MODE_CSQ>(vec, a, n_seg, s);
  K(seg_finalize_kernel<1><<<fin_grid, 1024, 0, s>>>(w.n.part0, slice_ptr, cnt, (int)n_seg, (int)channels, w.n.max_chunks, eps, rstd));
  int rc = check_launch("affnorm_fwd");
  if (rc != STINET_OK) return rc;
  return stinet_affnorm_apply(x, ldx, n_rows, channels, n_seg == 1 ? nullptr : gid, mean, rstd, alpha, gamma, beta, out, ldo, stream_);
}

extern "C" int stinet_affnorm_bwd(const float* x, int64_t ldx, const float* dy, int64_t ldg, int64_t n_rows,
                                  int64_t channels, int64_t n_seg, int64_t max_seg_rows, const int32_t* slice_ptr,
                                  const float* cnt, const int32_t* gid, int kind, const float* mean, const float* rstd,
                                  const float* alpha, const float* gamma, float* dx, int64_t lddx, float* dgamma,
                                  float* dbeta, float* dalpha, void* workspace, size_t workspace_bytes,
                                  stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && dy && slice_ptr && cnt && mean && rstd, STINET_ERR_ARG, "affnorm_bwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && channels > 0 && n_seg > 0 && n_seg <= 65535 && ldx >= channels && ldg >= channels &&
                     (!dx || lddx >= channels) && (kind == 0 || kind == 1) && (n_seg == 1 || gid),
                 STINET_ERR_ARG, "affnorm_bwd: bad shape / kind (several segments need gid)");
  AffWs w = carve_aff(workspace, max_seg_rows, channels, n_seg);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "affnorm_bwd: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  // T1 = sum_slice dy, T2 = sum_slice dy (x - alpha m): the backward column reduction with shift alpha*m and unit scale
  K(affnorm_prep_kernel<<<wave_grid(n_seg * channels, 256, 8), 256, 0, s>>>(mean, alpha, (int)n_seg, (int)channels, w.am, w.ones));
  const bool vec = nvec(channels, {x, dy}, {ldx, ldg});
  NormArgs a{x, ldx, dy, ldg, slice_ptr, nullptr, w.am, nullptr, w.n.part0, w.n.part1, (int)channels, w.n.max_chunks, STINET_ACT_NONE};
  launch_colreduce<MODE_BWD>(vec, a, n_seg, s);
  const dim3 fin_grid((unsigned)ceil_div(channels, 32), (unsigned)n_seg);
  K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.n.part0, slice_ptr, w.ones, (int)n_seg, (int)channels, w.n.max_chunks, 0.f, w.n.s1));
  K(seg_finalize_kernel<0><<<fin_grid, 1024, 0, s>>>(w.n.part1, slice_ptr, w.ones, (int)n_seg, (int)channels, w.n.max_chunks, 0.f, w.n.s2));
  if (dx != nullptr && n_rows > 0) {
    const bool v2 = nvec(channels, {x, dy, dx}, {ldx, ldg, lddx});
    const int grid = wave_grid(n_rows * (channels / (v2 ? 4 : 1)), kNormThreads * 4, 8, 8);
    const int32_t* g = n_seg == 1 ? nullptr : gid;
    if (v2) K(affnorm_bwd_apply_kernel<true><<<grid, kNormThreads, 0, s>>>(x, ldx, dy, ldg, n_rows, (int)channels, g, cnt, kind, mean, rstd, alpha, gamma, w.n.s1, w.n.s2, dx, lddx));
    else K(affnorm_bwd_apply_kernel<false><<<grid, kNormThreads, 0, s>>>(x, ldx, dy, ldg, n_rows, (int)channels, g, cnt, kind, mean, rstd, alpha, gamma, w.n.s1, w.n.s2, dx, lddx));
  }
  if (dgamma || dbeta || dalpha)
    K(affnorm_param_grads_kernel<<<(unsigned)ceil_div(channels, 128), 128, 0, s>>>(mean, rstd, gamma, w.n.s1, w.n.s2, (int)n_seg,
                                                                                  (int)channels, dgamma, dbeta, dalpha));
  return check_launch("affnorm_bwd");
}

extern "C" int stinet_bn_running_update(const float* mean, const float* rstd, int64_t n_rows, float eps, float momentum,
                                        int64_t channels, float* running_mean, float* running_var, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(mean && rstd && running_mean && running_var, STINET_ERR_ARG, "bn_running_update: null pointer");
  STINET_REQUIRE(channels > 0 && n_rows >= 0, STINET_ERR_ARG, "bn_running_update: bad shape");
  K(bn_running_update_kernel<<<(unsigned)ceil_div(channels, 128), 128, 0, s>>>(mean, rstd, (float)n_rows, eps, momentum, (int)channels,
                                                                             running_mean, running_var));
  return check_launch("bn_running_update");
}
