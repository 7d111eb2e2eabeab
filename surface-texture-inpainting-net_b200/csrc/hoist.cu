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

// ---------------------------------------------------------------------------------------------------------------
// Operand planes of ALL dense-layer weights of a network in two launches per step (instead of amax + split [+ hoist] per
// weight): a device table lists the weights, each CTA works on one 4096-element chunk of one weight.
//   kind 0: planes of W [rows, cols] itself (second Linear of a conv, shortcut, head)
//   kind 1: planes of the hoisted first layer  Wcat = [Wa - Wb ; Wb]  of EdgeConv, W = [Wa | Wb] [H, 2 din]   (+ bcat = [b ; 0])
//   kind 2: planes of  Wcat = [-W ; W]  of EdgeConvTransInv, W [H, din]                                        (+ bcat = [b ; 0])
// Pass 1 takes max|.| of the values that will be split (atomicMax of bit patterns into the entry's slot, zeroed by a
// memset), pass 2 reads it, scales and splits (csrc/common.cuh: split_one) and publishes the exponent.
namespace stinet {

struct WeightEntry {
  const float* w;        // source weight, row pitch ldw
  const float* b;        // kind 1/2: bias of the first Linear (nullable)
  __half* hi;            // planes [out_rows, ldp]
  __half* lo;            // nullable (one-pass mode)
  float* bcat;           // kind 1/2: [2H] (nullable)
  unsigned* amax;        // slot
  int32_t* exp;          // slot
  int64_t ldw, ldp;
  int32_t rows, cols;    // of the SOURCE view that is walked: kind 0: W; kind 1/2: H x din
  int32_t kind, chunk0;  // first global chunk index of this entry
};
constexpr int kWChunk = 4096;

__device__ __forceinline__ int find_entry(const WeightEntry* __restrict__ tab, int n_entries, int chunk) {
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {                         // last entry with chunk0 <= chunk
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].chunk0 <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// the (up to two) values a source element contributes: kind 0: (w, -); kind 1: (wa - wb, wb); kind 2: (-w, w)
__device__ __forceinline__ void weight_values(const WeightEntry& e, int64_t r, int64_t c, float& p, float& q) {
  if (e.kind == 0) {
    p = e.w[r * e.ldw + c];
    q = 0.f;
  } else if (e.kind == 1) {
    const float wa = e.w[r * e.ldw + c];
    q = e.w[r * e.ldw + e.cols + c];
    p = wa - q;
  } else {
    q = e.w[r * e.ldw + c];
    p = -q;
  }
}

// four consecutive source elements per thread when the row width and the pitches allow 16-byte accesses
__device__ __forceinline__ bool entry_vec(const WeightEntry& e) {
  return (e.cols & 3) == 0 && (e.ldw & 3) == 0 && aligned16(e.w) && (e.kind != 1 || ((e.cols & 3) == 0));
}
__device__ __forceinline__ void weight_values4(const WeightEntry& e, int64_t r, int64_t c, float4& p, float4& q) {
  const float4 a = *reinterpret_cast<const float4*>(e.w + r * e.ldw + c);
  if (e.kind == 0) {
    p = a;
    q = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (e.kind == 1) {
    q = *reinterpret_cast<const float4*>(e.w + r * e.ldw + e.cols + c);
    p = make_float4(a.x - q.x, a.y - q.y, a.z - q.z, a.w - q.w);
  } else {
    q = a;
    p = make_float4(-a.x, -a.y, -a.z, -a.w);
  }
}

__global__ void __launch_bounds__(256) wplanes_amax_kernel(const WeightEntry* __restrict__ tab, int n_entries) {
  const int ei = find_entry(tab, n_entries, blockIdx.x);
  const WeightEntry e = tab[ei];
  const int64_t total = (int64_t)e.rows * e.cols;
  const int64_t base = (int64_t)(blockIdx.x - e.chunk0) * kWChunk;
  const int64_t stop = min(total, base + kWChunk);
  unsigned m = 0u;
  if (entry_vec(e)) {
    for (int64_t idx = base + 4 * threadIdx.x; idx < stop; idx += 4 * 256) {
      const int64_t r = idx / e.cols, c = idx - r * e.cols;
      float4 p, q;
      weight_values4(e, r, c, p, q);
      m = amax4(amax4(m, p), q);
    }
  } else {
    for (int64_t idx = base + threadIdx.x; idx < stop; idx += 256) {
      const int64_t r = idx / e.cols, c = idx - r * e.cols;
      float p, q;
      weight_values(e, r, c, p, q);
      m = max(max(m, __float_as_uint(p) & 0x7FFFFFFFu), __float_as_uint(q) & 0x7FFFFFFFu);
    }
  }
  amax_publish(m, e.amax);
}

__global__ void __launch_bounds__(256) wplanes_split_kernel(const WeightEntry* __restrict__ tab, int n_entries) {
  const int ei = find_entry(tab, n_entries, blockIdx.x);
  const WeightEntry e = tab[ei];
  const int64_t total = (int64_t)e.rows * e.cols;
  const int64_t base = (int64_t)(blockIdx.x - e.chunk0) * kWChunk;
  const int64_t stop = min(total, base + kWChunk);
  const int sft = plane_shift(*e.amax);
  const float scale = plane_scale(sft);
  if (blockIdx.x == e.chunk0 && threadIdx.x == 0) *e.exp = -sft;
  if (entry_vec(e)) {                                  // ldp is a multiple of 8: 8-byte plane stores are aligned
    for (int64_t idx = base + 4 * threadIdx.x; idx < stop; idx += 4 * 256) {
      const int64_t r = idx / e.cols, c = idx - r * e.cols;
      float4 p, q;
      weight_values4(e, r, c, p, q);
      split_store4(p, scale, e.hi + r * e.ldp + c, e.lo ? e.lo + r * e.ldp + c : nullptr);
      if (e.kind != 0) {
        split_store4(q, scale, e.hi + (r + e.rows) * e.ldp + c, e.lo ? e.lo + (r + e.rows) * e.ldp + c : nullptr);
        if (e.bcat && c == 0) {
          e.bcat[r] = e.b ? e.b[r] : 0.f;
          e.bcat[e.rows + r] = 0.f;
        }
      }
    }
    return;
  }
  for (int64_t idx = base + threadIdx.x; idx < stop; idx += 256) {
    const int64_t r = idx / e.cols, c = idx - r * e.cols;
    float p, q;
    weight_values(e, r, c, p, q);
    __half h, l;
    split_one(p, scale, h, l);
    e.hi[r * e.ldp + c] = h;
    if (e.lo) e.lo[r * e.ldp + c] = l;
    if (e.kind != 0) {
      split_one(q, scale, h, l);
      e.hi[(r + e.rows) * e.ldp + c] = h;
      if (e.lo) e.lo[(r + e.rows) * e.ldp + c] = l;
      if (e.bcat && c == 0) {
        e.bcat[r] = e.b ? e.b[r] : 0.f;
        e.bcat[e.rows + r] = 0.f;
      }
    }
  }
}

}  // namespace stinet

extern "C" size_t stinet_weight_entry_bytes(void) { return sizeof(stinet::WeightEntry); }

// Fill one table entry on the HOST (the caller copies the table to the device once and reuses it every step).
// Returns the number of 4096-element chunks the entry occupies.
extern "C" long long stinet_weight_entry_fill(void* host_entry, const float* w, int64_t ldw, const float* b, int64_t rows,
                                            int64_t cols, int kind, void* hi, void* lo, int64_t ldp, float* bcat, float* amax,
                                            int32_t* exp, int64_t chunk0) {
  stinet::WeightEntry* e = static_cast<stinet::WeightEntry*>(host_entry);
  e->w = w; e->b = b; e->hi = static_cast<__half*>(hi); e->lo = static_cast<__half*>(lo); e->bcat = bcat;
  e->amax = reinterpret_cast<unsigned*>(amax); e->exp = exp; e->ldw = ldw; e->ldp = ldp;
  e->rows = (int32_t)rows; e->cols = (int32_t)cols; e->kind = kind; e->chunk0 = (int32_t)chunk0;
  return ceil_div(rows * cols > 0 ? rows * cols : 1, stinet::kWChunk);
}

extern "C" int stinet_weight_planes_refresh(const void* device_table, int n_entries, int64_t n_chunks, float* amax_slots,
                                            stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_entries >= 0 && n_chunks >= 0 && n_chunks < (1ll << 31), STINET_ERR_ARG, "weight_planes_refresh: bad size");
  if (n_entries == 0) return STINET_OK;
  STINET_REQUIRE(device_table && amax_slots, STINET_ERR_ARG, "weight_planes_refresh: null pointer");
  cudaError_t e = cudaMemsetAsync(amax_slots, 0, sizeof(float) * (size_t)n_entries, s);
  STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "weight_planes_refresh: cudaMemsetAsync: %s", cudaGetErrorString(e));
  const stinet::WeightEntry* tab = static_cast<const stinet::WeightEntry*>(device_table);
  K(stinet::wplanes_amax_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(tab, n_entries));
  K(stinet::wplanes_split_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(tab, n_entries));
  return check_launch("weight_planes_refresh");
}
