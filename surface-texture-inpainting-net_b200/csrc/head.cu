// The tail of the network and the trainer's loss, each as one pass:
//   head     : out = tanh(h W2^T + b2), W2 [3, C] -- the 64 -> 3 Linear is done in registers (a row of h is read once by
//              C/8 lanes, 8 channels each; three dot products reduced with shuffles) instead of a zero-padded GEMM + a
//              separate tanh pass (reference models/surfacetextureinpaintingnet.py:466-469);
//   masked L1: loss = mean(|where(mask > 0, out, color) - color| * 0.99^mask) and its gradient
//              (trainers/inpainting3d_trainer.py:127-137: torch.where + L1Loss(reduction='none') + pow + mean), instead of
//              ~10 elementwise / reduction launches over [N, 3].
// Reductions (dW2, db2, the loss) are two-stage with per-CTA partials added in a fixed order: deterministic, no atomics.
#include "common.cuh"

namespace stinet {

constexpr int kHeadThreads = 256;
constexpr int kHeadOut = 3;

// LPR = lanes per row = C / 8 (1, 2, 4, 8, 16 or 32)
template <int LPR>
__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ W, const float* __restrict__ b,
                int64_t n_rows, float* __restrict__ out) {
  constexpr int RPW = 32 / LPR;                       // rows per warp and iteration
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, grp = lane / LPR;
  float w[kHeadOut][8];
#pragma unroll
  for (int o = 0; o < kHeadOut; ++o)
#pragma unroll
    for (int k = 0; k < 8; ++k) w[o][k] = W[o * (LPR * 8) + sub * 8 + k];
  const float b0 = b ? b[0] : 0.f, b1 = b ? b[1] : 0.f, b2 = b ? b[2] : 0.f;
  const int64_t warp = (int64_t)blockIdx.x * (kHeadThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kHeadThreads / 32);
  for (int64_t r0 = warp * RPW; r0 < n_rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + grp;
    float a[kHeadOut] = {0.f, 0.f, 0.f};
    if (r < n_rows) {
      const float4 x0 = *reinterpret_cast<const float4*>(h + r * ldh + sub * 8);
      const float4 x1 = *reinterpret_cast<const float4*>(h + r * ldh + sub * 8 + 4);
      const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int o = 0; o < kHeadOut; ++o)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[o] = fmaf(x[k], w[o][k], a[o]);
    }
#pragma unroll
    for (int o = 0; o < kHeadOut; ++o)
#pragma unroll
      for (int off = LPR / 2; off >= 1; off >>= 1) a[o] += __shfl_xor_sync(0xffffffffu, a[o], off);
    if (sub == 0 && r < n_rows) {
      out[r * 3 + 0] = tanhf(a[0] + b0);
      out[r * 3 + 1] = tanhf(a[1] + b1);
      out[r * 3 + 2] = tanhf(a[2] + b2);
    }
  }
}

// dpre = dout (1 - out^2);  dh = dpre W2;  per-CTA partials of dW2 [3, C] and db2 [3] in part[cta][3 * C + 3]
template <int LPR>
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ W, const float* __restrict__ out,
                const float* __restrict__ dout, int64_t n_rows, float* __restrict__ dh, int64_t lddh, float* __restrict__ part) {
  constexpr int C = LPR * 8;
  constexpr int RPW = 32 / LPR;
  __shared__ float sm[kHeadThreads / 32][kHeadOut * C + kHeadOut];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  float w[kHeadOut][8], dw[kHeadOut][8], db[kHeadOut] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < kHeadOut; ++o)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      w[o][k] = W[o * C + sub * 8 + k];
      dw[o][k] = 0.f;
    }
  const int64_t warp = (int64_t)blockIdx.x * (kHeadThreads / 32) + wid;
  const int64_t nwarps = (int64_t)gridDim.x * (kHeadThreads / 32);
  for (int64_t r0 = warp * RPW; r0 < n_rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + grp;
    if (r < n_rows) {
      float dp[kHeadOut];
#pragma unroll
      for (int o = 0; o < kHeadOut; ++o) {
        const float y = out[r * 3 + o];
        dp[o] = dout[r * 3 + o] * (1.f - y * y);
      }
      const float4 x0 = *reinterpret_cast<const float4*>(h + r * ldh + sub * 8);
      const float4 x1 = *reinterpret_cast<const float4*>(h + r * ldh + sub * 8 + 4);
      const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      float g[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g[k] = dp[0] * w[0][k] + dp[1] * w[1][k] + dp[2] * w[2][k];
#pragma unroll
        for (int o = 0; o < kHeadOut; ++o) dw[o][k] = fmaf(dp[o], x[k], dw[o][k]);
      }
      if (sub == 0) {
#pragma unroll
        for (int o = 0; o < kHeadOut; ++o) db[o] += dp[o];
      }
      if (dh != nullptr) {
        *reinterpret_cast<float4*>(dh + r * lddh + sub * 8) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(dh + r * lddh + sub * 8 + 4) = make_float4(g[4], g[5], g[6], g[7]);
      }
    }
  }
  // rows of the warp's row groups -> one sum per (output, channel): lanes with the same `sub` across the groups
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
    for (int o = 0; o < kHeadOut; ++o) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dw[o][k] += __shfl_xor_sync(0xffffffffu, dw[o][k], off);
      db[o] += __shfl_xor_sync(0xffffffffu, db[o], off);
    }
  }
  if (grp == 0) {
#pragma unroll
    for (int o = 0; o < kHeadOut; ++o) {
#pragma unroll
      for (int k = 0; k < 8; ++k) sm[wid][o * C + sub * 8 + k] = dw[o][k];
      if (sub == 0) sm[wid][kHeadOut * C + o] = db[o];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHeadOut * C + kHeadOut; i += kHeadThreads) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < kHeadThreads / 32; ++y) t += sm[y][i];
    part[(int64_t)blockIdx.x * (kHeadOut * C + kHeadOut) + i] = t;
  }
}

// out[i] = sum over rows of part[rows][n] (fixed order: 32 row lanes, then a tree over the lanes)
__global__ void __launch_bounds__(1024) head_partials_sum_kernel(const float* __restrict__ part, int rows, int n, float* __restrict__ out0,
                                                                 int n0, float* __restrict__ out1) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float t = 0.f;
  if (i < n)
    for (int c = ty; c < rows; c += 32) t += part[(int64_t)c * n + i];
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && i < n) {
    float r = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) r += sm[y][tx];
    if (i < n0) out0[i] = r;
    else if (out1) out1[i - n0] = r;
  }
}

// ---- masked, discounted L1
__device__ __forceinline__ float l1_weight(float m) { return m > 0.f ? powf(0.99f, m) : 0.f; }

__global__ void __launch_bounds__(256) masked_l1_fwd_kernel(const float* __restrict__ out, const float* __restrict__ color,
                                                            const float* __restrict__ mask, int64_t n_rows, int channels,
                                                            float* __restrict__ part) {
  __shared__ float sm[8];
  float acc = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float wgt = l1_weight(mask[r]);
    if (wgt != 0.f) {
      float s = 0.f;
      for (int c = 0; c < channels; ++c) s += fabsf(out[r * channels + c] - color[r * channels + c]) * wgt;
      acc += s;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) masked_l1_final_kernel(const float* __restrict__ part, int n_part, float inv_count,
                                                               float* __restrict__ loss) {
  __shared__ float sm[32];
  float t = 0.f;
  for (int i = threadIdx.x; i < n_part; i += 1024) t += part[i];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = sm[threadIdx.x];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) *loss = r * inv_count;
  }
}
// dout[r,c] = g * [mask_r > 0] * 0.99^mask_r * sign(out - color) / count
__global__ void __launch_bounds__(256) masked_l1_bwd_kernel(const float* __restrict__ out, const float* __restrict__ color,
                                                            const float* __restrict__ mask, const float* __restrict__ gloss,
                                                            int64_t n_rows, int channels, float inv_count, float* __restrict__ dout) {
  const float g = __ldg(gloss) * inv_count;
  const int64_t total = n_rows * channels;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / channels;
    const float wgt = l1_weight(mask[r]);
    const float d = out[i] - color[i];
    dout[i] = wgt == 0.f ? 0.f : g * wgt * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
  }
}

static int head_grid(int64_t n_rows, int lpr) {
  return wave_grid(ceil_div(n_rows, 32 / lpr), kHeadThreads / 32, 4, 2);
}
static bool head_ok(int64_t channels) {
  return channels == 8 || channels == 16 || channels == 32 || channels == 64 || channels == 128 || channels == 256;
}

}  // namespace stinet

using namespace stinet;

#define HEAD_DISPATCH(LPR_, CALL)   \
  switch (LPR_) {                   \
    case 1: { constexpr int L = 1; CALL; break; }    \
    case 2: { constexpr int L = 2; CALL; break; }    \
    case 4: { constexpr int L = 4; CALL; break; }    \
    case 8: { constexpr int L = 8; CALL; break; }    \
    case 16: { constexpr int L = 16; CALL; break; }  \
    default: { constexpr int L = 32; CALL; break; }  \
  }

extern "C" size_t stinet_head_workspace_bytes(int64_t n_rows, int64_t channels) {
  if (n_rows < 0 || !head_ok(channels)) return 0;
  return sizeof(float) * (size_t)head_grid(n_rows > 0 ? n_rows : 1, (int)(channels / 8)) * (size_t)(kHeadOut * channels + kHeadOut);
}

extern "C" int stinet_head_fwd(const float* h, int64_t ldh, const float* W, const float* b, int64_t n_rows, int64_t channels,
                               float* out, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(h && W && out, STINET_ERR_ARG, "head_fwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && ldh >= channels, STINET_ERR_ARG, "head_fwd: bad shape");
  STINET_REQUIRE(head_ok(channels) && ldh % 4 == 0 && aligned16(h), STINET_ERR_UNSUPPORTED,
                 "head_fwd: channels must be 8..256 (power of two), rows 16-byte aligned");
  if (n_rows == 0) return STINET_OK;
  const int lpr = (int)(channels / 8);
  HEAD_DISPATCH(lpr, K(head_fwd_kernel<L><<<head_grid(n_rows, lpr), kHeadThreads, 0, s>>>(h, ldh, W, b, n_rows, out)));
  return check_launch("head_fwd");
}

extern "C" int stinet_head_bwd(const float* h, int64_t ldh, const float* W, const float* out, const float* dout, int64_t n_rows,
                               int64_t channels, float* dh, int64_t lddh, float* dW, float* db, void* workspace,
                               size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(h && W && out && dout && dW, STINET_ERR_ARG, "head_bwd: null pointer");
  STINET_REQUIRE(n_rows >= 0 && ldh >= channels && (!dh || lddh >= channels), STINET_ERR_ARG, "head_bwd: bad shape");
  STINET_REQUIRE(head_ok(channels) && ldh % 4 == 0 && aligned16(h) && (!dh || (lddh % 4 == 0 && aligned16(dh))),
                 STINET_ERR_UNSUPPORTED, "head_bwd: channels must be 8..256 (power of two), rows 16-byte aligned");
  const size_t need = stinet_head_workspace_bytes(n_rows, channels);
  STINET_REQUIRE(workspace && workspace_bytes >= need, STINET_ERR_WORKSPACE, "head_bwd: workspace %zu < %zu", workspace_bytes, need);
  const int lpr = (int)(channels / 8);
  const int grid = head_grid(n_rows > 0 ? n_rows : 1, lpr);
  float* part = static_cast<float*>(workspace);
  HEAD_DISPATCH(lpr, K(head_bwd_kernel<L><<<grid, kHeadThreads, 0, s>>>(h, ldh, W, out, dout, n_rows, dh, lddh, part)));
  const int n = (int)(kHeadOut * channels + kHeadOut);
  K(head_partials_sum_kernel<<<(unsigned)ceil_div(n, 32), 1024, 0, s>>>(part, grid, n, dW, (int)(kHeadOut * channels), db));
  return check_launch("head_bwd");
}

extern "C" size_t stinet_masked_l1_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return 0;
  return sizeof(float) * (size_t)wave_grid(n_rows > 0 ? n_rows : 1, 256, 4, 1);
}

extern "C" int stinet_masked_l1_fwd(const float* out, const float* color, const float* mask, int64_t n_rows, int64_t channels,
                                    float* loss, void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(out && color && mask && loss, STINET_ERR_ARG, "masked_l1_fwd: null pointer");
  STINET_REQUIRE(n_rows > 0 && channels > 0, STINET_ERR_ARG, "masked_l1_fwd: bad shape");
  const size_t need = stinet_masked_l1_workspace_bytes(n_rows);
  STINET_REQUIRE(workspace && workspace_bytes >= need, STINET_ERR_WORKSPACE, "masked_l1_fwd: workspace %zu < %zu", workspace_bytes, need);
  const int grid = wave_grid(n_rows, 256, 4, 1);
  float* part = static_cast<float*>(workspace);
  K(masked_l1_fwd_kernel<<<grid, 256, 0, s>>>(out, color, mask, n_rows, (int)channels, part));
  K(masked_l1_final_kernel<<<1, 1024, 0, s>>>(part, grid, 1.f / (float)(n_rows * channels), loss));
  return check_launch("masked_l1_fwd");
}

extern "C" int stinet_masked_l1_bwd(const float* out, const float* color, const float* mask, const float* gloss, int64_t n_rows,
                                    int64_t channels, float* dout, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(out && color && mask && gloss && dout, STINET_ERR_ARG, "masked_l1_bwd: null pointer");
  STINET_REQUIRE(n_rows > 0 && channels > 0, STINET_ERR_ARG, "masked_l1_bwd: bad shape");
  K(masked_l1_bwd_kernel<<<wave_grid(n_rows * channels, 256, 8), 256, 0, s>>>(out, color, mask, gloss, n_rows, (int)channels,
                                                                             1.f / (float)(n_rows * channels), dout));
  return check_launch("masked_l1_bwd");
}
