// Dense layers (nn.Linear fwd / dgrad / wgrad): entry points, planning and the exact-fp32 FFMA fallback.
//   C[i,j] = sum_t A'(i,t) * B'(t,j), with each operand stored either t-contiguous or i/j-contiguous, which covers
//   fwd   (A[M,K] W[N,K]^T),  dgrad (dC[M,N] W[N,K])  and  wgrad (dC[M,N]^T A[M,K], split over the row dimension
//   with a fixed-order second-stage reduction -- deterministic, no atomics).
// Main path: the tcgen05/TMA kernel of gemm_tc.cu (3xTF32 for fp32 parity, or bf16 operands cast into the workspace).
// Fallback for operands TMA cannot address (row pitch not a multiple of 16 B: the 10-channel input, the 3-channel
// head): 128x64x16 FFMA tiles, 256 threads, 8x4 outputs per thread, register-staged double buffering.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_tc.cuh"

namespace stinet {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int kGemmThreads = 256;

struct GemmArgs {
  const float* A; int64_t lda;     // A'(i,t): A_TC ? A[i*lda+t] : A[t*lda+i]
  const float* B; int64_t ldb;     // B'(t,j): B_TC ? B[j*ldb+t] : B[t*ldb+j]
  float* C; int64_t ldc;           // C[i*ldc+j]  (or split partial s: C + s*I*J, ldc = J)
  const float* bias; const int32_t* rowmask;
  int I, J, T;
  int t_per_split;
};

template <bool A_TC, bool B_TC>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int t_begin = blockIdx.z * g.t_per_split;
  const int t_end = min(g.T, t_begin + g.t_per_split);
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  // register staging: A tile = 2048 elems -> 8 per thread, B tile = 1024 -> 4 per thread
  float ra[8], rb[4];
  const bool a_vec = ((g.lda & 3) == 0) && aligned16(g.A);
  const bool b_vec = ((g.ldb & 3) == 0) && aligned16(g.B);

  auto load_tiles = [&](int t0) {
    // ---- A
    if (A_TC) {
      // (i,t): t fastest.  thread -> i = tid/4 + 64*r, t = (tid%4)*4 .. +3
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = i0 + (tid >> 2) + 64 * r;
        const int t = t0 + (tid & 3) * 4;
        const float* p = g.A + (int64_t)i * g.lda + t;
        if (a_vec && i < g.I && t + 3 < t_end) {
          float4 v = *reinterpret_cast<const float4*>(p);
          ra[r * 4 + 0] = v.x; ra[r * 4 + 1] = v.y; ra[r * 4 + 2] = v.z; ra[r * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) ra[r * 4 + q] = (i < g.I && t + q < t_end) ? p[q] : 0.f;
        }
      }
    } else {
      // (i,t): i fastest.  thread -> t = tid/32 + 8*r, i = (tid%32)*4 .. +3
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int t = t0 + (tid >> 5) + 8 * r;
        const int i = i0 + (tid & 31) * 4;
        const float* p = g.A + (int64_t)t * g.lda + i;
        if (a_vec && t < t_end && i + 3 < g.I) {
          float4 v = *reinterpret_cast<const float4*>(p);
          ra[r * 4 + 0] = v.x; ra[r * 4 + 1] = v.y; ra[r * 4 + 2] = v.z; ra[r * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) ra[r * 4 + q] = (t < t_end && i + q < g.I) ? p[q] : 0.f;
        }
      }
    }
    // ---- B
    if (B_TC) {
      // (t,j): t fastest.  thread -> j = tid/4, t = (tid%4)*4 .. +3
      const int j = j0 + (tid >> 2);
      const int t = t0 + (tid & 3) * 4;
      const float* p = g.B + (int64_t)j * g.ldb + t;
      if (b_vec && j < g.J && t + 3 < t_end) {
        float4 v = *reinterpret_cast<const float4*>(p);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) rb[q] = (j < g.J && t + q < t_end) ? p[q] : 0.f;
      }
    } else {
      // (t,j): j fastest.  thread -> t = tid/16, j = (tid%16)*4 .. +3
      const int t = t0 + (tid >> 4);
      const int j = j0 + (tid & 15) * 4;
      const float* p = g.B + (int64_t)t * g.ldb + j;
      if (b_vec && t < t_end && j + 3 < g.J) {
        float4 v = *reinterpret_cast<const float4*>(p);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) rb[q] = (t < t_end && j + q < g.J) ? p[q] : 0.f;
      }
    }
  };

  auto store_tiles = [&](int buf) {
    if (A_TC) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) As[buf][(tid & 3) * 4 + q][(tid >> 2) + 64 * r] = ra[r * 4 + q];
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r)
        *reinterpret_cast<float4*>(&As[buf][(tid >> 5) + 8 * r][(tid & 31) * 4]) =
            make_float4(ra[r * 4 + 0], ra[r * 4 + 1], ra[r * 4 + 2], ra[r * 4 + 3]);
    }
    if (B_TC) {
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[buf][(tid & 3) * 4 + q][tid >> 2] = rb[q];
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };

  int buf = 0;
  if (t_begin < t_end) {
    load_tiles(t_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int t0 = t_begin; t0 < t_end; t0 += BK) {
    const bool more = t0 + BK < t_end;
    if (more) load_tiles(t0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
    if (more) store_tiles(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  float* C = g.C + (int64_t)blockIdx.z * g.I * (int64_t)g.ldc;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int i = i0 + ty * 8 + a;
    if (i >= g.I) continue;
    const bool add_bias = g.bias && (!g.rowmask || g.rowmask[i] > 0);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j < g.J) C[(int64_t)i * g.ldc + j] = acc[a][c] + (add_bias ? g.bias[j] : 0.f);
    }
  }
}

// fixed-order sum of the split partials (+ bias on rows with rowmask > 0): deterministic second stage of a split GEMM
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t IJ, int J,
                                     float* __restrict__ out, int64_t ldo, const float* __restrict__ bias,
                                     const int32_t* __restrict__ rowmask, unsigned* __restrict__ amax_out = nullptr) {
  unsigned m = 0u;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < IJ;
       idx += (int64_t)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < splits; ++s) t += part[(int64_t)s * IJ + idx];
    const int64_t i = idx / J;
    const int j = (int)(idx - i * J);
    if (bias != nullptr && (rowmask == nullptr || rowmask[i] > 0)) t += bias[j];
    out[i * ldo + j] = t;
    m = max(m, __float_as_uint(t) & 0x7FFFFFFFu);
  }
  if (amax_out != nullptr) {   // max |out| as a bit pattern (one atomic per warp that can raise it)
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31) == 0 && m > *reinterpret_cast<volatile unsigned*>(amax_out)) atomicMax(amax_out, m);
  }
}

// the same for many splits of a small output (wgrad of the fine levels: up to 296 partials of a 64 x 128 matrix): the
// 8 warps of a CTA each sum every 8th split of 128 consecutive elements (128-bit loads), then warp 0 adds the eight
// partial sums in fixed order.  IJ % 4 == 0, J % 4 == 0, out rows 16-byte aligned.
__global__ void __launch_bounds__(256) splitk_reduce_wide_kernel(const float* __restrict__ part, int splits, int64_t IJ, int J,
                                                                 float* __restrict__ out, int64_t ldo) {
  __shared__ float4 sm[8][32];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t idx = ((int64_t)blockIdx.x * 32 + lane) * 4;
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx < IJ) {
    for (int s0 = wrp; s0 < splits; s0 += 32) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sp = s0 + 8 * u;
        v[u] = sp < splits ? ld_stream(reinterpret_cast<const float4*>(part + (int64_t)sp * IJ + idx))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { t.x += v[u].x; t.y += v[u].y; t.z += v[u].z; t.w += v[u].w; }
    }
  }
  sm[wrp][lane] = t;
  __syncthreads();
  if (wrp == 0 && idx < IJ) {
    float4 r = sm[0][lane];
#pragma unroll
    for (int y = 1; y < 8; ++y) { const float4 q = sm[y][lane]; r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w; }
    const int64_t i = idx / J;
    const int j = (int)(idx - i * J);
    *reinterpret_cast<float4*>(out + i * ldo + j) = r;
  }
}

// dbias: masked column sums of dC in two deterministic stages.
// stage 1: partial[chunk][n] = sum over the chunk's rows of dC[m,n] * mask(m).  Vector form: CTA = CL column lanes
//          (4 columns each: 128-bit loads, full 512 B / 256 B row segments) x 256/CL row lanes, four rows in flight
//          per thread; the rows of a chunk are chosen by the host so that even a 1296-row matrix fills the SMs.
template <int CL>
__global__ void __launch_bounds__(256) colsum_partial_vec_kernel(const float* __restrict__ x, int64_t ldx,
                                                                 const int32_t* __restrict__ rowmask, int64_t M, int N,
                                                                 int rows_per_chunk, float* __restrict__ part) {
  constexpr int RL = 256 / CL;
  __shared__ float4 sm[RL][CL];
  const int tx = threadIdx.x % CL, ty = threadIdx.x / CL;
  const int n0 = (blockIdx.x * CL + tx) * 4;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t m1 = min(M, m0 + rows_per_chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n0 < N) {
    for (int64_t m = m0 + ty; m < m1; m += 4 * RL) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t mm = m + u * RL;
        const bool on = mm < m1 && (rowmask == nullptr || rowmask[mm] > 0);
        v[u] = on ? ld_stream(reinterpret_cast<const float4*>(x + mm * ldx + n0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n0 < N) {
    float4 t = sm[0][tx];
#pragma unroll
    for (int y = 1; y < RL; ++y) { const float4 q = sm[y][tx]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
    *reinterpret_cast<float4*>(part + (int64_t)blockIdx.y * N + n0) = t;      // N % 4 == 0 on this path
  }
}
// scalar form for widths / pitches that are not multiples of four floats
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, int64_t ldx,
                                                             const int32_t* __restrict__ rowmask, int64_t M, int N,
                                                             int rows_per_chunk, float* __restrict__ part) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 32 + tx;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t m1 = min(M, m0 + rows_per_chunk);
  float acc = 0.f;
  if (n0 < N) {
    for (int64_t m = m0 + ty; m < m1; m += 8) {
      if (rowmask && rowmask[m] <= 0) continue;
      acc += x[m * ldx + n0];
    }
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n0 < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sm[y][tx];
    part[(int64_t)blockIdx.y * N + n0] = t;
  }
}
size_t colsum_part_floats(int64_t M, int64_t N);
// rows per stage-1 chunk: 512 for tall matrices, fewer (down to 32) when that is needed to put ~8 CTAs on every SM
static int colsum_rows(int64_t M, int64_t N) {
  const int64_t col_ctas = ceil_div(N, 128);
  int r = 512;
  while (r > 32 && col_ctas * ceil_div(M > 0 ? M : 1, r) < 8 * kSMs) r >>= 1;
  return r;
}
// stage 2: out[n] = sum_chunks partial[chunk][n]; 32 column lanes x 32 chunk lanes, fixed-order tree over the lanes
__global__ void __launch_bounds__(1024) colsum_final_kernel(const float* __restrict__ part, int chunks, int N,
                                                            float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float t = 0.f;
  if (n < N)
    for (int c = ty; c < chunks; c += 32) t += part[(int64_t)c * N + n];
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && n < N) {
    float r = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) r += sm[y][tx];
    out[n] = r;
  }
}


size_t colsum_part_floats(int64_t M, int64_t N) {
  const int64_t rows = M > 0 ? M : 1;
  return (size_t)ceil_div(rows, colsum_rows(rows, N)) * (size_t)N;
}

// masked column sums of dC [M,N] (dbias): two deterministic stages, partials in `part`
void run_colsum(const float* dC, int64_t ldc, const int32_t* rowmask, int64_t M, int64_t N, float* dbias,
                float* part, cudaStream_t s) {
  const int rpc = colsum_rows(M > 0 ? M : 1, N);
  const int chunks = (int)ceil_div(M > 0 ? M : 1, rpc);
  const bool vec = !(N & 3) && !(ldc & 3) && aligned16(dC);
  if (vec && N > 64) {
    dim3 g2((unsigned)ceil_div(N, 128), (unsigned)chunks);
    K(colsum_partial_vec_kernel<32><<<g2, 256, 0, s>>>(dC, ldc, rowmask, M, (int)N, rpc, part));
  } else if (vec) {
    dim3 g2((unsigned)ceil_div(N, 64), (unsigned)chunks);
    K(colsum_partial_vec_kernel<16><<<g2, 256, 0, s>>>(dC, ldc, rowmask, M, (int)N, rpc, part));
  } else {
    dim3 g2((unsigned)ceil_div(N, 32), (unsigned)chunks);
    K(colsum_partial_kernel<<<g2, 256, 0, s>>>(dC, ldc, rowmask, M, (int)N, rpc, part));
  }
  K(colsum_final_kernel<<<(unsigned)ceil_div(N, 32), 1024, 0, s>>>(part, chunks, (int)N, dbias));
}

// column sums of a matrix given as operand planes (dbias of the hoisted first layer: dP exists as planes only):
// stage 1 as colsum_partial_vec_kernel on plane values (still scaled by 2^s), stage 2 applies 2^exp.
__global__ void __launch_bounds__(256) colsum_planes_partial_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                                    int64_t ldp, int64_t M, int N, int rows_per_chunk,
                                                                    float* __restrict__ part) {
  constexpr int CL = 32, RL = 8;
  __shared__ float4 sm[RL][CL];
  const int tx = threadIdx.x % CL, ty = threadIdx.x / CL;
  const int n0 = (blockIdx.x * CL + tx) * 4;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t m1 = min(M, m0 + rows_per_chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n0 < N) {
    for (int64_t m = m0 + ty; m < m1; m += 4 * RL) {
      uint2 h[4], l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t mm = m + u * RL;
        h[u] = l[u] = make_uint2(0u, 0u);
        if (mm < m1) {
          h[u] = *reinterpret_cast<const uint2*>(hi + mm * ldp + n0);
          if (lo != nullptr) l[u] = *reinterpret_cast<const uint2*>(lo + mm * ldp + n0);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half* hh = reinterpret_cast<const __half*>(&h[u]);
        const __half* ll = reinterpret_cast<const __half*>(&l[u]);
        acc.x += plane_value(hh[0], ll[0]); acc.y += plane_value(hh[1], ll[1]);
        acc.z += plane_value(hh[2], ll[2]); acc.w += plane_value(hh[3], ll[3]);
      }
    }
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n0 < N) {
    float4 t = sm[0][tx];
#pragma unroll
    for (int y = 1; y < RL; ++y) { const float4 q = sm[y][tx]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
    *reinterpret_cast<float4*>(part + (int64_t)blockIdx.y * N + n0) = t;
  }
}
__global__ void __launch_bounds__(1024) colsum_planes_final_kernel(const float* __restrict__ part, int chunks, int N,
                                                                  const int32_t* __restrict__ exp, float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float t = 0.f;
  if (n < N)
    for (int c = ty; c < chunks; c += 32) t += part[(int64_t)c * N + n];
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && n < N) {
    float r = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) r += sm[y][tx];
    out[n] = ldexpf(r, __ldg(exp));
  }
}

// out[n] = sum_m x[m,n] of a matrix held as operand planes; `part` holds colsum_part_floats(M, N) floats
void run_colsum_planes(const __half* hi, const __half* lo, int64_t ldp, const int32_t* exp, int64_t M, int64_t N, float* out,
                       float* part, cudaStream_t s) {
  const int rpc = colsum_rows(M > 0 ? M : 1, N);
  const int chunks = (int)ceil_div(M > 0 ? M : 1, rpc);
  dim3 g2((unsigned)ceil_div(N, 128), (unsigned)chunks);
  K(colsum_planes_partial_kernel<<<g2, 256, 0, s>>>(hi, lo, ldp, M, (int)N, rpc, part));
  K(colsum_planes_final_kernel<<<(unsigned)ceil_div(N, 32), 1024, 0, s>>>(part, chunks, (int)N, exp, out));
}

// ---------------------------------------------------------------------------------------------------------------
// operand planes for the fp16 tensor-core modes (see include/stinet_b200.h, "dense layers on operand PLANES")

// max |x| as a bit pattern: non-negative floats order like their bit patterns (NaN above inf, so NaN propagates).
// One atomicMax per CTA, skipped when the CTA cannot raise the value (a stale read can only under-estimate it).
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols,
                                                   unsigned* __restrict__ amax) {
  __shared__ unsigned sm[8];
  unsigned m = 0;
  if ((cols & 3) == 0 && (ldx & 3) == 0 && aligned16(x)) {
    const int cpr = cols >> 2;
    const int64_t total = rows * cpr;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = idx / cpr;
      const int c = (int)(idx - r * cpr) << 2;
      const float4 v = ld_stream(reinterpret_cast<const float4*>(x + r * ldx + c));
      m = max(max(m, __float_as_uint(v.x) & 0x7FFFFFFFu), __float_as_uint(v.y) & 0x7FFFFFFFu);
      m = max(max(m, __float_as_uint(v.z) & 0x7FFFFFFFu), __float_as_uint(v.w) & 0x7FFFFFFFu);
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = idx / cols;
      m = max(m, __float_as_uint(x[r * ldx + (idx - r * cols)]) & 0x7FFFFFFFu);
    }
  }
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = max(m, sm[w]);
    if (m > *reinterpret_cast<volatile unsigned*>(amax)) atomicMax(amax, m);
  }
}

// x [rows, cols] fp32 -> hi / lo fp16 planes (pitch ldp) of x 2^s; exp_out = -s.  Vector form: 8 elements per thread.
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols,
                                                        const unsigned* __restrict__ amax, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, int64_t ldp, int32_t* __restrict__ exp_out) {
  const int sft = plane_shift(__ldg(amax));
  const float scale = plane_scale(sft);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = -sft;
  if ((cols & 7) == 0 && (ldx & 3) == 0 && aligned16(x)) {
    const int cpr = cols >> 3;
    const int64_t total = rows * cpr;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = idx / cpr;
      const int c = (int)(idx - r * cpr) << 3;
      const float4 a = ld_stream(reinterpret_cast<const float4*>(x + r * ldx + c));
      const float4 b = ld_stream(reinterpret_cast<const float4*>(x + r * ldx + c + 4));
      __half h[8], l[8];
      split_one(a.x, scale, h[0], l[0]); split_one(a.y, scale, h[1], l[1]);
      split_one(a.z, scale, h[2], l[2]); split_one(a.w, scale, h[3], l[3]);
      split_one(b.x, scale, h[4], l[4]); split_one(b.y, scale, h[5], l[5]);
      split_one(b.z, scale, h[6], l[6]); split_one(b.w, scale, h[7], l[7]);
      *reinterpret_cast<uint4*>(hi + r * ldp + c) = *reinterpret_cast<const uint4*>(h);
      if (lo != nullptr) *reinterpret_cast<uint4*>(lo + r * ldp + c) = *reinterpret_cast<const uint4*>(l);
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = idx / cols;
      const int64_t c = idx - r * cols;
      __half h, l;
      split_one(x[r * ldx + c], scale, h, l);
      hi[r * ldp + c] = h;
      if (lo != nullptr) lo[r * ldp + c] = l;
    }
  }
}

// split + masked column sums in one pass over dC (the backward of a Linear with bias needs both: the planes for dgrad /
// wgrad and dbias): CTA = CL column lanes (8 columns each) x 256/CL row lanes over one row chunk; stage-1 partials as in
// colsum_partial_vec_kernel (same chunking, same second stage), so dbias keeps its fixed summation order.
template <int CL>
__global__ void __launch_bounds__(256) split_colsum_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int N,
                                                           const unsigned* __restrict__ amax, const int32_t* __restrict__ rowmask,
                                                           int rows_per_chunk, __half* __restrict__ hi, __half* __restrict__ lo,
                                                           int64_t ldp, int32_t* __restrict__ exp_out, float* __restrict__ part) {
  constexpr int RL = 256 / CL;
  __shared__ float4 sm[2][RL][CL];
  const int sft = plane_shift(__ldg(amax));
  const float scale = plane_scale(sft);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *exp_out = -sft;
  const int tx = threadIdx.x % CL, ty = threadIdx.x / CL;
  const int n0 = (blockIdx.x * CL + tx) * 8;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t m1 = min(M, m0 + rows_per_chunk);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  if (n0 < N) {
    for (int64_t m = m0 + ty; m < m1; m += 2 * RL) {
      float4 v[2][2];
      bool on[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t mm = m + u * RL;
        on[u] = false;
        v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mm < m1) {
          v[u][0] = ld_stream(reinterpret_cast<const float4*>(x + mm * ldx + n0));
          v[u][1] = ld_stream(reinterpret_cast<const float4*>(x + mm * ldx + n0 + 4));
          on[u] = rowmask == nullptr || rowmask[mm] > 0;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t mm = m + u * RL;
        if (mm < m1) {
          split_store4(v[u][0], scale, hi + mm * ldp + n0, lo != nullptr ? lo + mm * ldp + n0 : nullptr);
          split_store4(v[u][1], scale, hi + mm * ldp + n0 + 4, lo != nullptr ? lo + mm * ldp + n0 + 4 : nullptr);
          if (on[u]) {
            a0.x += v[u][0].x; a0.y += v[u][0].y; a0.z += v[u][0].z; a0.w += v[u][0].w;
            a1.x += v[u][1].x; a1.y += v[u][1].y; a1.z += v[u][1].z; a1.w += v[u][1].w;
          }
        }
      }
    }
  }
  sm[0][ty][tx] = a0;
  sm[1][ty][tx] = a1;
  __syncthreads();
  if (ty < 2 && n0 < N) {          // row lane 0 sums the first four columns of the group, row lane 1 the other four
    float4 t = sm[ty][0][tx];
#pragma unroll
    for (int y = 1; y < RL; ++y) { const float4 q = sm[ty][y][tx]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
    *reinterpret_cast<float4*>(part + (int64_t)blockIdx.y * N + n0 + 4 * ty) = t;
  }
}

// fp32 -> bf16 cast of a row-major matrix (bf16 mode: the tensor-core kernel reads bf16 operands through TMA).
// cols % 8 == 0, 16 B aligned rows on both sides; one thread converts 8 elements.
// With a `lo` plane (bf16x3 mode) the rounding residual x - float(hi) is stored too, so hi + lo carries 16 significand bits.
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols,
                                                        __nv_bfloat16* __restrict__ y, int64_t ldy,
                                                        __nv_bfloat16* __restrict__ lo) {
  const int cpr = cols >> 3;
  const int64_t total = rows * cpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cpr;
    const int c = (int)(idx % cpr) << 3;
    const float4 a = ld_stream(reinterpret_cast<const float4*>(x + r * ldx + c));
    const float4 b = ld_stream(reinterpret_cast<const float4*>(x + r * ldx + c + 4));
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(y + r * ldy + c) = o;
    if (lo != nullptr) {
      const float2 h0 = __bfloat1622float2(p0), h1 = __bfloat1622float2(p1), h2 = __bfloat1622float2(p2),
                   h3 = __bfloat1622float2(p3);
      __nv_bfloat162 q0 = __floats2bfloat162_rn(a.x - h0.x, a.y - h0.y), q1 = __floats2bfloat162_rn(a.z - h1.x, a.w - h1.y);
      __nv_bfloat162 q2 = __floats2bfloat162_rn(b.x - h2.x, b.y - h2.y), q3 = __floats2bfloat162_rn(b.z - h3.x, b.w - h3.y);
      uint4 l;
      l.x = *reinterpret_cast<uint32_t*>(&q0); l.y = *reinterpret_cast<uint32_t*>(&q1);
      l.z = *reinterpret_cast<uint32_t*>(&q2); l.w = *reinterpret_cast<uint32_t*>(&q3);
      *reinterpret_cast<uint4*>(lo + r * ldy + c) = l;
    }
  }
}

// fp32 -> (hi, lo) TF32 planes for the pre-split 3xTF32 GEMM: hi = x rounded to TF32 (10 explicit significand bits,
// the same integer rounding the in-kernel split uses), lo = x - hi (exact in fp32).  Packed planes, pitch = cols.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols,
                                                         float* __restrict__ hi, float* __restrict__ lo) {
  const int cpr = cols >> 2;
  const int64_t total = rows * cpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cpr;
    const int c = (int)(idx % cpr) << 2;
    const float4 v = ld_stream(reinterpret_cast<const float4*>(x + r * ldx + c));
    float4 h, l;
    h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u); l.x = v.x - h.x;
    h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u); l.y = v.y - h.y;
    h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u); l.z = v.z - h.z;
    h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi + r * cols + c) = h;
    *reinterpret_cast<float4*>(lo + r * cols + c) = l;
  }
}

static int pick_splits(int64_t I, int64_t J, int64_t T) {
  int64_t tiles = ceil_div(I, BM) * ceil_div(J, BN);
  int64_t want = ceil_div(2 * kSMs, tiles);
  int64_t max_splits = ceil_div(T, BK * 16);
  int64_t s = want < max_splits ? want : max_splits;
  return (int)(s < 1 ? 1 : s);
}

// Split of the reduction dimension for the persistent tensor-core kernel (one CTA per SM walks the work units).
// Enough output tiles to fill the SMs: no split.  Otherwise pick the split count S that minimises the number of
// unit rounds per unit of work, ceil(tiles*S / 148) / S, with at least 256 reduction elements per split and whole
// 64-element k-blocks.  wgrad (few tiles, very long reduction) may use many splits, fwd / dgrad at most 8.
struct TcSplit { int splits; int64_t t_per; };
static TcSplit tc_split(int64_t I, int64_t J, int64_t T, int max_splits) {
  const int64_t bn = J <= 64 ? 64 : 128;
  const int64_t tiles = ceil_div(I, 128) * ceil_div(J, bn);
  int64_t best = 1;
  if (tiles < (kSMs * 7) / 8) {
    int64_t cap = T / 256;
    if (cap > max_splits) cap = max_splits;
    double best_cost = (double)ceil_div(tiles, kSMs);
    for (int64_t sp = 2; sp <= cap; ++sp) {
      const double cost = (double)ceil_div(tiles * sp, kSMs) / (double)sp;
      if (cost < best_cost * 0.97) { best_cost = cost; best = sp; }
    }
  }
  int64_t t_per = ceil_div(ceil_div(T, best), 64) * 64;
  if (t_per < 64) t_per = 64;
  return TcSplit{(int)ceil_div(T, t_per), t_per};
}
constexpr int kWgradMaxSplits = 2 * kSMs, kFwdMaxSplits = 8;

static int tc_mode(int precision) {
  switch (precision) {
    case STINET_PREC_FP32: return tc::MODE_TF32X3;
    case STINET_PREC_BF16: return tc::MODE_BF16;
    case STINET_PREC_TF32: return tc::MODE_TF32X1;
    case STINET_PREC_BF16X3: return tc::MODE_BF16X3;
    default: return -1;
  }
}
static bool valid_precision(int p) { return p >= STINET_PREC_FP32 && p <= STINET_PREC_BF16X3; }
static bool is_bf16_mode(int mode) { return mode == tc::MODE_BF16 || mode == tc::MODE_BF16X3; }

// fp32 mode, C[I,J] over T: split the operands in HBM first (MODE_TF32X3P) when the GEMM is compute-heavy enough that
// the in-kernel split (shared-memory bandwidth) is the limiter and the two extra HBM passes are cheap next to it:
// flops per operand byte = I*J / (2*(I+J)) above the threshold STINET_TC_PRESPLIT (unset or 0: disabled).
static bool tc_presplit(int precision, int64_t I, int64_t J, int64_t T) {
  // Measured on B200 (scripts/presplit_ab.sh, profiles/r1_h_gemm.md): the kernel itself runs 1.3-1.4x faster on planes,
  // but the two split passes per call eat most of it (fwd 4096x1024 +13 %, wgrad -6 %), so the mode is OFF unless asked
  // for; it pays once the planes are produced by the operand's producer or cached across fwd / dgrad / wgrad.
  static const int thr = [] { const char* e = getenv("STINET_TC_PRESPLIT"); return e ? atoi(e) : 0; }();
  if (precision != STINET_PREC_FP32 || thr <= 0) return false;
  if ((I | J | T) & 3) return false;                       // packed planes must keep 16-byte row pitches
  return I * J >= (int64_t)thr * 2 * (I + J) && T >= 256;
}

struct GemmWs {
  float *a32, *w32, *c32;                  // fp32 pre-split mode: hi planes of A [M,K], W [N,K], dC [M,N] ...
  float *a32lo, *w32lo, *c32lo;            // ... and their lo planes (nullptr when no entry point of this shape pre-splits)
  float *splitk, *colsum;
  __nv_bfloat16 *a16, *w16, *c16;          // bf16 copies of A [M,K], W [N,K], dC [M,N]
  __nv_bfloat16 *a16lo, *w16lo, *c16lo;    // bf16x3 mode: the residual planes (nullptr otherwise)
  size_t bytes;
};
static GemmWs carve_gemm(void* base, int64_t M, int64_t N, int64_t K, int precision) {
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  GemmWs w;
  const int64_t rows = M > 0 ? M : 1;
  int splits = pick_splits(N, K, rows);
  const int tcs = tc_split(N, K, rows, kWgradMaxSplits).splits;
  if (tcs > splits) splits = tcs;
  size_t sk_elems = (size_t)splits * N * K;                                                   // wgrad partials
  const int sf = tc_split(rows, N, K, kFwdMaxSplits).splits, sd = tc_split(rows, K, N, kFwdMaxSplits).splits;
  if (sf > 1 && (size_t)sf * rows * N > sk_elems) sk_elems = (size_t)sf * rows * N;          // fwd partials
  if (sd > 1 && (size_t)sd * rows * K > sk_elems) sk_elems = (size_t)sd * rows * K;          // dgrad partials
  const size_t sk = up(sizeof(float) * sk_elems);
  const size_t cs = up(sizeof(float) * (size_t)ceil_div(rows, colsum_rows(rows, N)) * N);
  const bool x3 = precision == STINET_PREC_BF16X3;
  const bool b16 = precision == STINET_PREC_BF16 || x3;
  const size_t planes = x3 ? 2 : 1;
  const size_t a1 = up(2 * (size_t)rows * K), w1 = up(2 * (size_t)N * K), c1 = up(2 * (size_t)rows * N);
  const size_t a16 = b16 ? planes * a1 : 0, w16 = b16 ? planes * w1 : 0, c16 = b16 ? planes * c1 : 0;
  char* p = static_cast<char*>(base);
  w.splitk = reinterpret_cast<float*>(p);
  w.colsum = reinterpret_cast<float*>(p + sk);
  w.a16 = reinterpret_cast<__nv_bfloat16*>(p + sk + cs);
  w.w16 = reinterpret_cast<__nv_bfloat16*>(p + sk + cs + a16);
  w.c16 = reinterpret_cast<__nv_bfloat16*>(p + sk + cs + a16 + w16);
  w.a16lo = x3 ? reinterpret_cast<__nv_bfloat16*>(p + sk + cs + a1) : nullptr;
  w.w16lo = x3 ? reinterpret_cast<__nv_bfloat16*>(p + sk + cs + a16 + w1) : nullptr;
  w.c16lo = x3 ? reinterpret_cast<__nv_bfloat16*>(p + sk + cs + a16 + w16 + c1) : nullptr;
  w.bytes = sk + cs + a16 + w16 + c16;
  // fp32 pre-split planes (any of fwd / dgrad / wgrad of this shape may ask for them)
  w.a32 = w.w32 = w.c32 = w.a32lo = w.w32lo = w.c32lo = nullptr;
  if (M > 0 && (tc_presplit(precision, rows, N, K) || tc_presplit(precision, rows, K, N) || tc_presplit(precision, N, K, rows))) {
    const size_t pa = up(4 * (size_t)rows * K), pw = up(4 * (size_t)N * K), pc = up(4 * (size_t)rows * N);
    char* q = p + w.bytes;
    w.a32 = reinterpret_cast<float*>(q);            w.a32lo = reinterpret_cast<float*>(q + pa);
    w.w32 = reinterpret_cast<float*>(q + 2 * pa);   w.w32lo = reinterpret_cast<float*>(q + 2 * pa + pw);
    w.c32 = reinterpret_cast<float*>(q + 2 * pa + 2 * pw);
    w.c32lo = reinterpret_cast<float*>(q + 2 * pa + 2 * pw + pc);
    w.bytes += 2 * (pa + pw + pc);
  }
  return w;
}

// fp32 matrix the cast kernel (and TMA) can address with 16-byte accesses
static bool castable(const float* x, int64_t ld, int64_t cols) { return aligned16(x) && ld % 4 == 0 && cols % 8 == 0; }
static void cast_bf16(const float* x, int64_t ld, int64_t rows, int64_t cols, __nv_bfloat16* y, __nv_bfloat16* lo,
                      cudaStream_t s) {
  K(cast_bf16_kernel<<<wave_grid(rows * (cols / 8), 256, 8), 256, 0, s>>>(x, ld, rows, (int)cols, y, cols, lo));
}

static void split_tf32(const float* x, int64_t ld, int64_t rows, int64_t cols, float* hi, float* lo, cudaStream_t s) {
  K(split_tf32_kernel<<<wave_grid(rows * (cols / 4), 256, 8), 256, 0, s>>>(x, ld, rows, (int)cols, hi, lo));
}
// fp32 matrix the split kernel can read with 16-byte accesses
static bool splittable(const float* x, int64_t ld, int64_t cols) { return aligned16(x) && ld % 4 == 0 && cols % 4 == 0; }

}  // namespace stinet

using namespace stinet;

extern "C" size_t stinet_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int precision) {
  if (M < 0 || N <= 0 || K <= 0) return 0;
  return carve_gemm(nullptr, M, N, K, precision).bytes;
}

extern "C" int stinet_linear_fwd(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                                 const int32_t* rowmask, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                 int precision, void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(A && W && C, STINET_ERR_ARG, "linear_fwd: null pointer");
  STINET_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, STINET_ERR_ARG, "linear_fwd: bad shape");
  STINET_REQUIRE(valid_precision(precision), STINET_ERR_ARG, "linear_fwd: unknown precision %d", precision);
  if (M == 0) return STINET_OK;
  const int mode = tc_mode(precision);
  if (mode >= 0) {
    tc::Problem p{A, lda, false, W, ldw, false, C, ldc, bias, rowmask, M, N, K, 1, K, mode};
    bool ok = true;
    if (is_bf16_mode(mode)) {
      GemmWs w = carve_gemm(workspace, M, N, K, precision);
      STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "linear_fwd: workspace %zu < %zu",
                     workspace_bytes, w.bytes);
      ok = castable(A, lda, K) && castable(W, ldw, K);
      if (ok) {
        cast_bf16(A, lda, M, K, w.a16, w.a16lo, s);
        cast_bf16(W, ldw, N, K, w.w16, w.w16lo, s);
        p.A = w.a16; p.lda = K; p.B = w.w16; p.ldb = K; p.A_lo = w.a16lo; p.B_lo = w.w16lo;
      }
    } else if (tc_presplit(precision, M, N, K) && splittable(A, lda, K) && splittable(W, ldw, K)) {
      GemmWs w = carve_gemm(workspace, M, N, K, precision);
      if (workspace && workspace_bytes >= w.bytes && w.a32) {
        split_tf32(A, lda, M, K, w.a32, w.a32lo, s);
        split_tf32(W, ldw, N, K, w.w32, w.w32lo, s);
        p.A = w.a32; p.lda = K; p.B = w.w32; p.ldb = K; p.A_lo = w.a32lo; p.B_lo = w.w32lo;
        p.mode = tc::MODE_TF32X3P;
      }
    }
    if (ok && tc::eligible(p)) {
      const TcSplit sp = tc_split(M, N, K, kFwdMaxSplits);
      if (sp.splits > 1) {
        GemmWs w = carve_gemm(workspace, M, N, K, precision);
        if (workspace && workspace_bytes >= w.bytes) {     // without a workspace the unsplit kernel is still correct
          p.C = w.splitk; p.ldc = N; p.bias = nullptr; p.rowmask = nullptr;
          p.splits = sp.splits; p.t_per_split = sp.t_per;
          int rc = tc::run(p, s);
          if (rc) return rc;
          K(splitk_reduce_kernel<<<wave_grid(M * N, 256, 8), 256, 0, s>>>(w.splitk, sp.splits, M * N, (int)N, C, ldc,
                                                                        bias, rowmask));
          return check_launch("linear_fwd");
        }
      }
      return tc::run(p, s);
    }
  }
  GemmArgs g{A, lda, W, ldw, C, ldc, bias, rowmask, (int)M, (int)N, (int)K, (int)K};
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM), 1);
  K(gemm_kernel<true, true><<<grid, kGemmThreads, 0, s>>>(g));
  return check_launch("linear_fwd");
}

extern "C" int stinet_linear_dgrad(const float* dC, int64_t ldc, const float* W, int64_t ldw, float* dA, int64_t lda,
                                   int64_t M, int64_t N, int64_t K, int precision, void* workspace,
                                   size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(dC && W && dA, STINET_ERR_ARG, "linear_dgrad: null pointer");
  STINET_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, STINET_ERR_ARG, "linear_dgrad: bad shape");
  STINET_REQUIRE(valid_precision(precision), STINET_ERR_ARG, "linear_dgrad: unknown precision %d", precision);
  if (M == 0) return STINET_OK;
  // dA[i=m, j=k] = sum_{t=n} dC[m,n] * W[n,k]
  const int mode = tc_mode(precision);
  if (mode >= 0) {
    tc::Problem p{dC, ldc, false, W, ldw, true, dA, lda, nullptr, nullptr, M, K, N, 1, N, mode};
    bool ok = true;
    if (is_bf16_mode(mode)) {
      GemmWs w = carve_gemm(workspace, M, N, K, precision);
      STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "linear_dgrad: workspace %zu < %zu",
                     workspace_bytes, w.bytes);
      ok = castable(dC, ldc, N) && castable(W, ldw, K);
      if (ok) {
        cast_bf16(dC, ldc, M, N, w.c16, w.c16lo, s);
        cast_bf16(W, ldw, N, K, w.w16, w.w16lo, s);
        p.A = w.c16; p.lda = N; p.B = w.w16; p.ldb = K; p.A_lo = w.c16lo; p.B_lo = w.w16lo;
      }
    } else if (tc_presplit(precision, M, K, N) && splittable(dC, ldc, N) && splittable(W, ldw, K)) {
      GemmWs w = carve_gemm(workspace, M, N, K, precision);
      if (workspace && workspace_bytes >= w.bytes && w.c32) {
        split_tf32(dC, ldc, M, N, w.c32, w.c32lo, s);
        split_tf32(W, ldw, N, K, w.w32, w.w32lo, s);
        p.A = w.c32; p.lda = N; p.B = w.w32; p.ldb = K; p.A_lo = w.c32lo; p.B_lo = w.w32lo;
        p.mode = tc::MODE_TF32X3P;
      }
    }
    if (ok && tc::eligible(p)) {
      const TcSplit sp = tc_split(M, K, N, kFwdMaxSplits);
      if (sp.splits > 1) {
        GemmWs w = carve_gemm(workspace, M, N, K, precision);
        if (workspace && workspace_bytes >= w.bytes) {
          p.C = w.splitk; p.ldc = K; p.splits = sp.splits; p.t_per_split = sp.t_per;
          int rc = tc::run(p, s);
          if (rc) return rc;
          K(splitk_reduce_kernel<<<wave_grid(M * K, 256, 8), 256, 0, s>>>(w.splitk, sp.splits, M * K, (int)K, dA, lda,
                                                                        nullptr, nullptr));
          return check_launch("linear_dgrad");
        }
      }
      return tc::run(p, s);
    }
  }
  GemmArgs g{dC, ldc, W, ldw, dA, lda, nullptr, nullptr, (int)M, (int)K, (int)N, (int)N};
  dim3 grid((unsigned)ceil_div(K, BN), (unsigned)ceil_div(M, BM), 1);
  K(gemm_kernel<true, false><<<grid, kGemmThreads, 0, s>>>(g));
  return check_launch("linear_dgrad");
}

extern "C" int stinet_linear_wgrad(const float* dC, int64_t ldc, const float* A, int64_t lda, const int32_t* rowmask,
                                   float* dW, int64_t ldw, float* dbias, int64_t M, int64_t N, int64_t K,
                                   int precision, void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(dC && A && dW, STINET_ERR_ARG, "linear_wgrad: null pointer");
  STINET_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, STINET_ERR_ARG, "linear_wgrad: bad shape");
  STINET_REQUIRE(valid_precision(precision), STINET_ERR_ARG, "linear_wgrad: unknown precision %d", precision);
  GemmWs w = carve_gemm(workspace, M, N, K, precision);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "linear_wgrad: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  // dW[i=n, j=k] = sum_{t=m} dC[m,n] * A[m,k]
  bool done = false;
  const int mode = tc_mode(precision);
  if (mode >= 0 && M > 0) {
    const TcSplit sp = tc_split(N, K, M, kWgradMaxSplits);
    tc::Problem p{dC, ldc, true, A, lda, true, sp.splits == 1 ? dW : w.splitk, sp.splits == 1 ? ldw : K,
                  nullptr, nullptr, N, K, M, sp.splits, sp.t_per, mode};
    bool ok = true;
    if (is_bf16_mode(mode)) {
      ok = castable(dC, ldc, N) && castable(A, lda, K);
      if (ok) {
        cast_bf16(dC, ldc, M, N, w.c16, w.c16lo, s);
        cast_bf16(A, lda, M, K, w.a16, w.a16lo, s);
        p.A = w.c16; p.lda = N; p.B = w.a16; p.ldb = K; p.A_lo = w.c16lo; p.B_lo = w.a16lo;
      }
    } else if (tc_presplit(precision, N, K, M) && w.c32 && splittable(dC, ldc, N) && splittable(A, lda, K)) {
      split_tf32(dC, ldc, M, N, w.c32, w.c32lo, s);
      split_tf32(A, lda, M, K, w.a32, w.a32lo, s);
      p.A = w.c32; p.lda = N; p.B = w.a32; p.ldb = K; p.A_lo = w.c32lo; p.B_lo = w.a32lo;
      p.mode = tc::MODE_TF32X3P;
    }
    if (ok && tc::eligible(p)) {
      int rc = tc::run(p, s);
      if (rc) return rc;
      if (sp.splits >= 16 && !(K & 3) && !(ldw & 3) && aligned16(dW))
        K(splitk_reduce_wide_kernel<<<(unsigned)ceil_div(N * K, 128), 256, 0, s>>>(w.splitk, sp.splits, N * K, (int)K, dW, ldw));
      else if (sp.splits > 1)
        K(splitk_reduce_kernel<<<wave_grid(N * K, 256, 8), 256, 0, s>>>(w.splitk, sp.splits, N * K, (int)K, dW, ldw, nullptr,
                                                                        nullptr));
      done = true;
    }
  }
  if (!done) {
    const int splits = pick_splits(N, K, M);
    const int t_per = (int)(ceil_div(ceil_div(M > 0 ? M : 1, splits), BK) * BK);
    dim3 grid((unsigned)ceil_div(K, BN), (unsigned)ceil_div(N, BM), (unsigned)splits);
    if (splits == 1) {
      GemmArgs g{dC, ldc, A, lda, dW, ldw, nullptr, nullptr, (int)N, (int)K, (int)M, t_per};
      K(gemm_kernel<false, false><<<grid, kGemmThreads, 0, s>>>(g));
    } else {
      GemmArgs g{dC, ldc, A, lda, w.splitk, K, nullptr, nullptr, (int)N, (int)K, (int)M, t_per};
      K(gemm_kernel<false, false><<<grid, kGemmThreads, 0, s>>>(g));
      K(splitk_reduce_kernel<<<wave_grid(N * K, 256, 8), 256, 0, s>>>(w.splitk, splits, N * K, (int)K, dW, ldw, nullptr,
                                                                      nullptr));
    }
  }
  if (dbias) run_colsum(dC, ldc, rowmask, M, N, dbias, w.colsum, s);
  return check_launch("linear_wgrad");
}

// ---------------------------------------------------------------------------------------------------------------
// dense layers on fp16 operand planes

extern "C" int stinet_f16_amax(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* amax, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && amax, STINET_ERR_ARG, "f16_amax: null pointer");
  STINET_REQUIRE(rows >= 0 && cols > 0 && ldx >= cols && cols < (1ll << 31), STINET_ERR_ARG, "f16_amax: bad shape");
  cudaError_t e = cudaMemsetAsync(amax, 0, sizeof(float), s);
  STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "f16_amax: cudaMemsetAsync: %s", cudaGetErrorString(e));
  if (rows == 0) return STINET_OK;
  K(amax_kernel<<<wave_grid(rows * ceil_div(cols, 4), 256 * 4, 8), 256, 0, s>>>(x, ldx, rows, (int)cols,
                                                                               reinterpret_cast<unsigned*>(amax)));
  return check_launch("f16_amax");
}

extern "C" int stinet_f16_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* amax, void* hi, void* lo,
                                int64_t ldp, int32_t* exp_out, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && amax && hi && exp_out, STINET_ERR_ARG, "f16_split: null pointer");
  STINET_REQUIRE(rows >= 0 && cols > 0 && ldx >= cols && ldp >= cols && ldp % 8 == 0 && cols < (1ll << 31), STINET_ERR_ARG,
                 "f16_split: bad shape (plane pitch must be a multiple of 8 elements)");
  STINET_REQUIRE(aligned16(hi) && (lo == nullptr || aligned16(lo)), STINET_ERR_ARG, "f16_split: planes must be 16-byte aligned");
  // rows == 0 still has to publish the exponent: one CTA does
  K(split_f16_kernel<<<wave_grid(rows * ceil_div(cols, 8), 256 * 2, 8), 256, 0, s>>>(
      x, ldx, rows, (int)cols, reinterpret_cast<const unsigned*>(amax), static_cast<__half*>(hi), static_cast<__half*>(lo),
      ldp, exp_out));
  return check_launch("f16_split");
}

extern "C" int stinet_colsum(const float* dC, int64_t ldc, const int32_t* rowmask, int64_t M, int64_t N, float* dbias,
                             void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(dC && dbias, STINET_ERR_ARG, "colsum: null pointer");
  STINET_REQUIRE(M >= 0 && N > 0 && ldc >= N, STINET_ERR_ARG, "colsum: bad shape");
  GemmWs w = carve_gemm(workspace, M, N, 1, STINET_PREC_FP32);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "colsum: workspace %zu < %zu", workspace_bytes,
                 w.bytes);
  run_colsum(dC, ldc, rowmask, M, N, dbias, w.colsum, s);
  return check_launch("colsum");
}

static int f16_mode(int passes) { return passes == 3 ? tc::MODE_F16X3 : passes == 1 ? tc::MODE_F16X1 : -1; }

// one planes GEMM  C[I,J] = sum_t A'(i,t) B'(t,j) 2^(a_exp + b_exp), split over t when the tiles do not fill the SMs
static int run_planes(tc::Problem p, int max_splits, const float* bias, const int32_t* rowmask, float* C, int64_t ldc,
                      float* amax_out, const GemmWs& w, bool have_ws, const char* what, cudaStream_t s) {
  unsigned* amax_bits = reinterpret_cast<unsigned*>(amax_out);
  if (amax_bits != nullptr) {
    cudaError_t e = cudaMemsetAsync(amax_bits, 0, sizeof(unsigned), s);
    STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "%s: cudaMemsetAsync: %s", what, cudaGetErrorString(e));
  }
  STINET_REQUIRE(tc::eligible(p), STINET_ERR_UNSUPPORTED,
                 "%s: operand planes / output not addressable by TMA (16-byte pitches, output width %% 4)", what);
  const TcSplit sp = tc_split(p.I, p.J, p.T, max_splits);
  if (sp.splits > 1 && have_ws) {
    p.C = w.splitk; p.ldc = p.J; p.bias = nullptr; p.rowmask = nullptr;
    p.splits = sp.splits; p.t_per_split = sp.t_per;
    int rc = tc::run(p, s);
    if (rc) return rc;
    const int64_t IJ = p.I * p.J;
    if (sp.splits >= 16 && bias == nullptr && amax_bits == nullptr && !(p.J & 3) && !(ldc & 3) && aligned16(C))
      K(splitk_reduce_wide_kernel<<<(unsigned)ceil_div(IJ, 128), 256, 0, s>>>(w.splitk, sp.splits, IJ, (int)p.J, C, ldc));
    else
      K(splitk_reduce_kernel<<<wave_grid(IJ, 256, 8), 256, 0, s>>>(w.splitk, sp.splits, IJ, (int)p.J, C, ldc, bias, rowmask,
                                                                   amax_bits));
    return check_launch(what);
  }
  p.C = C; p.ldc = ldc; p.bias = bias; p.rowmask = rowmask; p.splits = 1; p.t_per_split = p.T; p.amax_out = amax_bits;
  return tc::run(p, s);
}

extern "C" int stinet_linear_fwd_f16(const void* A_hi, const void* A_lo, int64_t lda, const int32_t* a_exp, const void* W_hi,
                                     const void* W_lo, int64_t ldw, const int32_t* w_exp, const float* bias,
                                     const int32_t* rowmask, float* C, int64_t ldc, float* amax_out, int64_t M, int64_t N,
                                     int64_t K, int passes, void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int mode = f16_mode(passes);
  STINET_REQUIRE(mode >= 0, STINET_ERR_ARG, "linear_fwd_f16: passes must be 1 or 3");
  STINET_REQUIRE(A_hi && W_hi && C && a_exp && w_exp && (passes == 1 || (A_lo && W_lo)), STINET_ERR_ARG, "linear_fwd_f16: null pointer");
  STINET_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, STINET_ERR_ARG, "linear_fwd_f16: bad shape");
  if (M == 0) return STINET_OK;
  GemmWs w = carve_gemm(workspace, M, N, K, STINET_PREC_FP32);
  tc::Problem p{A_hi, lda, false, W_hi, ldw, false, C, ldc, bias, rowmask, M, N, K, 1, K, mode, A_lo, W_lo, a_exp, w_exp};
  return run_planes(p, kFwdMaxSplits, bias, rowmask, C, ldc, amax_out, w, workspace && workspace_bytes >= w.bytes, "linear_fwd_f16", s);
}

extern "C" int stinet_linear_dgrad_f16(const void* dC_hi, const void* dC_lo, int64_t ldc, const int32_t* c_exp, const void* W_hi,
                                       const void* W_lo, int64_t ldw, const int32_t* w_exp, float* dA, int64_t lda,
                                       float* amax_out, int64_t M, int64_t N, int64_t K, int passes, void* workspace,
                                       size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int mode = f16_mode(passes);
  STINET_REQUIRE(mode >= 0, STINET_ERR_ARG, "linear_dgrad_f16: passes must be 1 or 3");
  STINET_REQUIRE(dC_hi && W_hi && dA && c_exp && w_exp && (passes == 1 || (dC_lo && W_lo)), STINET_ERR_ARG, "linear_dgrad_f16: null pointer");
  STINET_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, STINET_ERR_ARG, "linear_dgrad_f16: bad shape");
  if (M == 0) return STINET_OK;
  GemmWs w = carve_gemm(workspace, M, N, K, STINET_PREC_FP32);
  // dA[i=m, j=k] = sum_{t=n} dC[m,n] * W[n,k]
  tc::Problem p{dC_hi, ldc, false, W_hi, ldw, true, dA, lda, nullptr, nullptr, M, K, N, 1, N, mode, dC_lo, W_lo, c_exp, w_exp};
  return run_planes(p, kFwdMaxSplits, nullptr, nullptr, dA, lda, amax_out, w, workspace && workspace_bytes >= w.bytes, "linear_dgrad_f16", s);
}

extern "C" int stinet_linear_wgrad_f16(const void* dC_hi, const void* dC_lo, int64_t ldc, const int32_t* c_exp, const void* A_hi,
                                       const void* A_lo, int64_t lda, const int32_t* a_exp, float* dW, int64_t ldw, int64_t M,
                                       int64_t N, int64_t K, int passes, void* workspace, size_t workspace_bytes,
                                       stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int mode = f16_mode(passes);
  STINET_REQUIRE(mode >= 0, STINET_ERR_ARG, "linear_wgrad_f16: passes must be 1 or 3");
  STINET_REQUIRE(dC_hi && A_hi && dW && c_exp && a_exp && (passes == 1 || (dC_lo && A_lo)), STINET_ERR_ARG, "linear_wgrad_f16: null pointer");
  STINET_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, STINET_ERR_ARG, "linear_wgrad_f16: bad shape");
  GemmWs w = carve_gemm(workspace, M, N, K, STINET_PREC_FP32);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "linear_wgrad_f16: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  // dW[i=n, j=k] = sum_{t=m} dC[m,n] * A[m,k]
  tc::Problem p{dC_hi, ldc, true, A_hi, lda, true, dW, ldw, nullptr, nullptr, N, K, M, 1, M, mode, dC_lo, A_lo, c_exp, a_exp};
  return run_planes(p, kWgradMaxSplits, nullptr, nullptr, dW, ldw, nullptr, w, true, "linear_wgrad_f16", s);
}

extern "C" int stinet_colsum_planes(const void* hi, const void* lo, int64_t ldp, const int32_t* exp, int64_t M, int64_t N,
                                    float* out, void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(hi && exp && out, STINET_ERR_ARG, "colsum_planes: null pointer");
  STINET_REQUIRE(M >= 0 && N > 0 && N % 4 == 0 && ldp >= N && ldp % 4 == 0, STINET_ERR_ARG, "colsum_planes: bad shape");
  STINET_REQUIRE((reinterpret_cast<uintptr_t>(hi) & 7u) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7u) == 0, STINET_ERR_ARG,
                 "colsum_planes: planes must be 8-byte aligned");
  GemmWs w = carve_gemm(workspace, M, N, 1, STINET_PREC_FP32);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "colsum_planes: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  run_colsum_planes(static_cast<const __half*>(hi), static_cast<const __half*>(lo), ldp, exp, M, N, out, w.colsum, s);
  return check_launch("colsum_planes");
}

extern "C" int stinet_f16_split_colsum(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* amax,
                                       const int32_t* rowmask, void* hi, void* lo, int64_t ldp, int32_t* exp_out,
                                       float* colsum, void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && amax && hi && exp_out && colsum, STINET_ERR_ARG, "f16_split_colsum: null pointer");
  STINET_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && ldx >= cols && ldx % 4 == 0 && ldp >= cols && ldp % 8 == 0 &&
                     cols < (1ll << 31) && aligned16(x) && aligned16(hi) && (lo == nullptr || aligned16(lo)),
                 STINET_ERR_UNSUPPORTED, "f16_split_colsum: needs cols %% 8 == 0 and 16-byte aligned rows");
  GemmWs w = carve_gemm(workspace, rows, cols, 1, STINET_PREC_FP32);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "f16_split_colsum: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  const int rpc = colsum_rows(rows, cols);
  const int chunks = (int)ceil_div(rows, rpc);
  const unsigned* am = reinterpret_cast<const unsigned*>(amax);
  if (cols > 128) {
    dim3 g2((unsigned)ceil_div(cols, 256), (unsigned)chunks);
    K(split_colsum_kernel<32><<<g2, 256, 0, s>>>(x, ldx, rows, (int)cols, am, rowmask, rpc, static_cast<__half*>(hi),
                                                 static_cast<__half*>(lo), ldp, exp_out, w.colsum));
  } else {
    dim3 g2((unsigned)ceil_div(cols, 64), (unsigned)chunks);
    K(split_colsum_kernel<8><<<g2, 256, 0, s>>>(x, ldx, rows, (int)cols, am, rowmask, rpc, static_cast<__half*>(hi),
                                                static_cast<__half*>(lo), ldp, exp_out, w.colsum));
  }
  K(colsum_final_kernel<<<(unsigned)ceil_div(cols, 32), 1024, 0, s>>>(w.colsum, chunks, (int)cols, colsum));
  return check_launch("f16_split_colsum");
}
