// Hoisted first-layer parameters of EdgeConv (models/modules/edge_conv_filter.py:46-57 in the reference):
//   nn.0([x_i || x_j - x_i]) = (Wa - Wb) x_i + Wb x_j + b     ->  Wcat = [Wa - Wb ; Wb],  bcat = [b ; 0]
//   nn.0(x_j - x_i)          = (-W) x_i + W x_j + b           ->  Wcat = [-W ; W]          (EdgeConvTransInv)
// One elementwise kernel each way instead of the slice / sub / neg / cat chain (and its autograd mirror) per block.
#include "common.cuh"

namespace stinet {

__global__ void __launch_bounds__(256)
hoist_fwd_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ b, int64_t H, int64_t din,
                 int trans_inv, float* __restrict__ Wcat, float* __restrict__ bcat) {
  const int64_t total = H * din;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / din, c = idx - r * din;
    float p, q;
    if (trans_inv) {
      q = W[r * ldw + c];
      p = -q;
    } else {
      const float wa = W[r * ldw + c];
      q = W[r * ldw + din + c];
      p = wa - q;
    }
    Wcat[idx] = p;
    Wcat[total + idx] = q;
    if (bcat && idx < H) {
      bcat[idx] = b[idx];
      bcat[H + idx] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(256)
hoist_bwd_kernel(const float* __restrict__ dWcat, const float* __restrict__ dbcat, int64_t H, int64_t din,
                 int trans_inv, float* __restrict__ dW, int64_t ldw, float* __restrict__ db) {
  const int64_t total = H * din;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / din, c = idx - r * din;
    const float gp = dWcat[idx], gq = dWcat[total + idx];
    if (trans_inv) {
      dW[r * ldw + c] = gq - gp;
    } else {
      dW[r * ldw + c] = gp;
      dW[r * ldw + din + c] = gq - gp;
    }
    if (db && idx < H) db[idx] = dbcat[idx];
  }
}

}  // namespace stinet

using namespace stinet;

extern "C" int stinet_edgeconv_hoist_fwd(const float* W, int64_t ldw, const float* b, int64_t hidden, int64_t din,
                                         int trans_inv, float* Wcat, float* bcat, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(W && Wcat && ((b == nullptr) == (bcat == nullptr)), STINET_ERR_ARG, "edgeconv_hoist_fwd: null pointer");
  STINET_REQUIRE(hidden > 0 && din > 0 && ldw >= (trans_inv ? din : 2 * din), STINET_ERR_ARG, "edgeconv_hoist_fwd: bad shape");
  K(hoist_fwd_kernel<<<wave_grid(hidden * din, 256 * 4, 8), 256, 0, s>>>(W, ldw, b, hidden, din, trans_inv, Wcat, bcat));
  return check_launch("edgeconv_hoist_fwd");
}

extern "C" int stinet_edgeconv_hoist_bwd(const float* dWcat, const float* dbcat, int64_t hidden, int64_t din,
                                         int trans_inv, float* dW, int64_t ldw, float* db, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(dWcat && dW && (!db || dbcat), STINET_ERR_ARG, "edgeconv_hoist_bwd: null pointer");
  STINET_REQUIRE(hidden > 0 && din > 0 && ldw >= (trans_inv ? din : 2 * din), STINET_ERR_ARG, "edgeconv_hoist_bwd: bad shape");
  K(hoist_bwd_kernel<<<wave_grid(hidden * din, 256 * 4, 8), 256, 0, s>>>(dWcat, dbcat, hidden, din, trans_inv, dW, ldw, db));
  return check_launch("edgeconv_hoist_bwd");
}
