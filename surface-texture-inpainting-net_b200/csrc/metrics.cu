// Per-step graph metrics of the 3D trainer (reference utils/metrics/graph_metrics.py:6-72, called after every
// training / validation step at trainers/inpainting3d_trainer.py:254-263) on the level-0 CSR the forward pass
// already built: graph Laplacian (aggr='add' propagate), its variance on the grey channel, graph total variation and
// (masked) PSNR without materialising [E,3] gathers or boolean-indexed copies.
// Scalars are reduced deterministically: per-CTA double partials (fixed tree) + one fixed-order finalize CTA.
#include "common.cuh"

namespace stinet {

constexpr int kMetThreads = 256;
constexpr int kMetMaxCtas = kSMs * 8;
constexpr int kMetK = 2;  // doubles per partial

inline int met_grid(int64_t n) {
  int64_t g = ceil_div(n > 0 ? n : 1, kMetThreads);
  return (int)(g < kMetMaxCtas ? g : kMetMaxCtas);
}

__device__ __forceinline__ void block_sum2(double a, double b, double* part) {
  __shared__ double sm[2][kMetThreads];
  sm[0][threadIdx.x] = a;
  sm[1][threadIdx.x] = b;
  __syncthreads();
  for (int off = kMetThreads / 2; off >= 1; off >>= 1) {
    if ((int)threadIdx.x < off) {
      sm[0][threadIdx.x] += sm[0][threadIdx.x + off];
      sm[1][threadIdx.x] += sm[1][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[blockIdx.x * kMetK + 0] = sm[0][0];
    part[blockIdx.x * kMetK + 1] = sm[1][0];
  }
}

__device__ __forceinline__ float grey(const float* __restrict__ x, int64_t ldx, int64_t i) {
  const float* r = x + i * ldx;
  return 0.299f * r[0] + 0.587f * r[1] + 0.114f * r[2];   // graph_metrics.py:25
}

// out[i,c] = sum_{j->i} x[j,c] - deg_i * x[i,c]; neighbours are summed in original edge order
__global__ void __launch_bounds__(kMetThreads)
laplace_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
               const int32_t* __restrict__ col, int64_t n, int channels, float* __restrict__ out, int64_t ldo) {
  const int64_t total = n * channels;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / channels;
    const int c = (int)(idx - i * channels);
    const int beg = rowptr[i], end = rowptr[i + 1];
    float s = 0.f;
    for (int k = beg; k < end; ++k) s += x[(int64_t)col[k] * ldx + c];
    out[i * ldo + c] = s - (float)(end - beg) * x[i * ldx + c];
  }
}

// partial (sum l, sum l^2) of l_i = laplace(grey(x))_i
__global__ void __launch_bounds__(kMetThreads)
lapvar_partial_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                      const int32_t* __restrict__ col, int64_t n, double* __restrict__ part) {
  double a = 0.0, b = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int beg = rowptr[i], end = rowptr[i + 1];
    float s = 0.f;
    for (int k = beg; k < end; ++k) s += grey(x, ldx, col[k]);
    const float l = s - (float)(end - beg) * grey(x, ldx, i);
    a += (double)l;
    b += (double)l * (double)l;
  }
  block_sum2(a, b, part);
}

// partial sum over in-edges (every edge exactly once) of sum_c |x[j,c] - x[i,c]|
__global__ void __launch_bounds__(kMetThreads)
tv_partial_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                  const int32_t* __restrict__ col, int64_t n, int channels, double* __restrict__ part) {
  double a = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float* xi = x + i * ldx;
    for (int k = beg; k < end; ++k) {
      const float* xj = x + (int64_t)col[k] * ldx;
      float t = 0.f;
      for (int c = 0; c < channels; ++c) t += fabsf(xj[c] - xi[c]);
      a += (double)t;
    }
  }
  block_sum2(a, 0.0, part);
}

// partial (sum ((x-y)/range)^2, rows used) over rows with mask > 0 (mask NULL: every row)
__global__ void __launch_bounds__(kMetThreads)
sqerr_partial_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ y, int64_t ldy,
                     const float* __restrict__ mask, int64_t n, int channels, float range,
                     double* __restrict__ part) {
  double a = 0.0, b = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (mask && !(mask[i] > 0.f)) continue;
    for (int c = 0; c < channels; ++c) {
      const float d = x[i * ldx + c] / range - y[i * ldy + c] / range;   // graph_metrics.py:61-62,70
      a += (double)d * (double)d;
    }
    b += 1.0;
  }
  block_sum2(a, b, part);
}

enum { FIN_LAPVAR = 0, FIN_TV = 1, FIN_PSNR = 2 };

__global__ void __launch_bounds__(kMetThreads)
metric_finalize_kernel(const double* __restrict__ part, int nparts, int mode, double denom, float* __restrict__ out) {
  __shared__ double sm[2][kMetThreads];
  double a = 0.0, b = 0.0;
  for (int p = threadIdx.x; p < nparts; p += kMetThreads) {
    a += part[p * kMetK + 0];
    b += part[p * kMetK + 1];
  }
  sm[0][threadIdx.x] = a;
  sm[1][threadIdx.x] = b;
  __syncthreads();
  for (int off = kMetThreads / 2; off >= 1; off >>= 1) {
    if ((int)threadIdx.x < off) {
      sm[0][threadIdx.x] += sm[0][threadIdx.x + off];
      sm[1][threadIdx.x] += sm[1][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double s0 = sm[0][0], s1 = sm[1][0];
    if (mode == FIN_LAPVAR) {            // biased variance (torch.var(unbiased=False), :30)
      const double m = s0 / denom;
      out[0] = (float)(s1 / denom - m * m);
    } else if (mode == FIN_TV) {         // / (h*w), :36
      out[0] = (float)(s0 / denom);
    } else {                             // -10 log10(mse + 1e-8), :70-71; denom = channels, s1 = rows used
      const double mse = s0 / (s1 * denom);
      out[0] = (float)(-10.0 * log10(mse + 1e-8));
      out[1] = (float)s1;
    }
  }
}

}  // namespace stinet

using namespace stinet;

extern "C" size_t stinet_metrics_workspace_bytes(int64_t n) {
  return sizeof(double) * kMetK * (size_t)met_grid(n);
}

extern "C" int stinet_graph_laplace(const float* x, int64_t ldx, const int32_t* rowptr_t, const int32_t* col_t,
                                    int64_t n, int64_t channels, float* out, int64_t ldo, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(x && rowptr_t && out, STINET_ERR_ARG, "graph_laplace: null pointer");
  STINET_REQUIRE(n >= 0 && channels > 0 && ldx >= channels && ldo >= channels, STINET_ERR_ARG, "graph_laplace: bad shape");
  if (n == 0) return STINET_OK;
  K(laplace_kernel<<<wave_grid(n * channels, kMetThreads, 8), kMetThreads, 0, s>>>(x, ldx, rowptr_t, col_t, n, (int)channels, out, ldo));
  return check_launch("graph_laplace");
}

#define MET_COMMON(what)                                                                                      \
  cudaStream_t s = static_cast<cudaStream_t>(stream_);                                                        \
  STINET_REQUIRE(x && out && workspace, STINET_ERR_ARG, what ": null pointer");                               \
  STINET_REQUIRE(n > 0, STINET_ERR_ARG, what ": empty graph");                                                \
  STINET_REQUIRE(workspace_bytes >= stinet_metrics_workspace_bytes(n), STINET_ERR_WORKSPACE, what ": workspace"); \
  double* part = static_cast<double*>(workspace);                                                             \
  const int grid = met_grid(n)

extern "C" int stinet_graph_laplace_variance(const float* x, int64_t ldx, const int32_t* rowptr_t,
                                             const int32_t* col_t, int64_t n, float* out, void* workspace,
                                             size_t workspace_bytes, stinet_stream_t stream_) {
  MET_COMMON("graph_laplace_variance");
  STINET_REQUIRE(rowptr_t && ldx >= 3, STINET_ERR_ARG, "graph_laplace_variance: needs RGB rows");
  K(lapvar_partial_kernel<<<grid, kMetThreads, 0, s>>>(x, ldx, rowptr_t, col_t, n, part));
  K(metric_finalize_kernel<<<1, kMetThreads, 0, s>>>(part, grid, FIN_LAPVAR, (double)n, out));
  return check_launch("graph_laplace_variance");
}

extern "C" int stinet_graph_total_variation(const float* x, int64_t ldx, const int32_t* rowptr_t,
                                            const int32_t* col_t, int64_t n, int64_t channels, float* out,
                                            void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  MET_COMMON("graph_total_variation");
  STINET_REQUIRE(rowptr_t && channels > 0 && ldx >= channels, STINET_ERR_ARG, "graph_total_variation: bad shape");
  K(tv_partial_kernel<<<grid, kMetThreads, 0, s>>>(x, ldx, rowptr_t, col_t, n, (int)channels, part));
  K(metric_finalize_kernel<<<1, kMetThreads, 0, s>>>(part, grid, FIN_TV, (double)n * (double)channels, out));
  return check_launch("graph_total_variation");
}

extern "C" int stinet_psnr(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* mask, int64_t n,
                           int64_t channels, float data_range, float* out, void* workspace, size_t workspace_bytes,
                           stinet_stream_t stream_) {
  MET_COMMON("psnr");
  STINET_REQUIRE(y && channels > 0 && ldx >= channels && ldy >= channels && data_range > 0.f, STINET_ERR_ARG, "psnr: bad shape");
  K(sqerr_partial_kernel<<<grid, kMetThreads, 0, s>>>(x, ldx, y, ldy, mask, n, (int)channels, data_range, part));
  K(metric_finalize_kernel<<<1, kMetThreads, 0, s>>>(part, grid, FIN_PSNR, (double)channels, out));
  return check_launch("psnr");
}
