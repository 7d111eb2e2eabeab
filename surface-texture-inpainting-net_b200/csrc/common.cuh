// Shared helpers for the stinet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/stinet_b200.h"

namespace stinet {

void set_error(const char* fmt, ...);
void count_launch();  // process-wide count of kernels this library has launched (stinet_launch_count)

// K(kernel<<<grid, block, smem, stream>>>(args...));  -- every launch goes through this so it is counted
#define K(...)                 \
  do {                         \
    __VA_ARGS__;               \
    ::stinet::count_launch();  \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return STINET_ERR_CUDA;
  }
  return STINET_OK;
}

#define STINET_REQUIRE(cond, code, ...)  \
  do {                                   \
    if (!(cond)) {                       \
      ::stinet::set_error(__VA_ARGS__);  \
      return (code);                     \
    }                                    \
  } while (0)

// deterministic two-stage masked column sums of a row-major fp32 matrix (gemm.cu): out[n] = sum_m x[m,n] * (rowmask ?
// rowmask[m] > 0 : 1); `part` holds colsum_part_floats(M, N) floats
void run_colsum(const float* x, int64_t ldx, const int32_t* rowmask, int64_t M, int64_t N, float* out, float* part,
                cudaStream_t s);
size_t colsum_part_floats(int64_t M, int64_t N);
// the same over a matrix held as fp16 operand planes (hi + lo 2^-11) 2^exp
void run_colsum_planes(const __half* hi, const __half* lo, int64_t ldp, const int32_t* exp, int64_t M, int64_t N, float* out,
                       float* part, cudaStream_t s);

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid for a grid-stride kernel: enough CTAs to cover `work` items at `per_cta` each, rounded to whole waves of
// `ctas_per_sm` CTAs on 148 SMs and capped at `max_waves` waves.
inline int wave_grid(int64_t work, int64_t per_cta, int ctas_per_sm, int max_waves = 8) {
  int64_t need = ceil_div(work > 0 ? work : 1, per_cta);
  int64_t wave = (int64_t)kSMs * ctas_per_sm;
  int64_t waves = ceil_div(need, wave);
  if (waves > max_waves) waves = max_waves;
  if (need < wave) return (int)need;
  return (int)(waves * wave);
}

// streaming (read-once) 128-bit load that does not pollute L1, and the matching store
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

// ---- fp16 operand planes of the dense layers (include/stinet_b200.h, "dense layers on operand PLANES")
// s with bound 2^s in [2^14, 2^15) (bound = 0, inf or NaN: s = 0), clamped so that 2^s and 2^-s stay normal floats;
// `bound_bits` is the fp32 bit pattern of max|x| or of any upper bound of it
__device__ __forceinline__ int plane_shift(unsigned bound_bits) {
  const int e = (int)((bound_bits >> 23) & 0xFFu);
  if ((bound_bits & 0x7FFFFFFFu) == 0u || e == 255) return 0;
  const int s = 15 - (max(e, 1) - 126);
  return max(-110, min(110, s));
}
__device__ __forceinline__ float plane_scale(int shift) { return __uint_as_float((uint32_t)(127 + shift) << 23); }
// hi = fp16(x 2^s), lo = fp16((x 2^s - hi) 2^11)
__device__ __forceinline__ void split_one(float x, float scale, __half& hi, __half& lo) {
  const float xs = x * scale;
  hi = __float2half_rn(xs);
  lo = __float2half_rn((xs - __half2float(hi)) * 2048.f);
}
// four consecutive elements -> 8 bytes of each plane
__device__ __forceinline__ void split_store4(float4 v, float scale, __half* hi, __half* lo) {
  __half h[4], l[4];
  split_one(v.x, scale, h[0], l[0]); split_one(v.y, scale, h[1], l[1]);
  split_one(v.z, scale, h[2], l[2]); split_one(v.w, scale, h[3], l[3]);
  *reinterpret_cast<uint2*>(hi) = *reinterpret_cast<const uint2*>(h);
  if (lo != nullptr) *reinterpret_cast<uint2*>(lo) = *reinterpret_cast<const uint2*>(l);
}
// the value a plane pair stands for, still scaled by 2^s
__device__ __forceinline__ float plane_value(__half hi, __half lo) { return __half2float(hi) + __half2float(lo) * (1.f / 2048.f); }
// max |v| of a float4 folded into a running bit-pattern maximum
__device__ __forceinline__ unsigned amax4(unsigned m, float4 v) {
  m = max(max(m, __float_as_uint(v.x) & 0x7FFFFFFFu), __float_as_uint(v.y) & 0x7FFFFFFFu);
  return max(max(m, __float_as_uint(v.z) & 0x7FFFFFFFu), __float_as_uint(v.w) & 0x7FFFFFFFu);
}
// one atomicMax per warp, skipped when the warp cannot raise the value (a stale read can only under-estimate it)
__device__ __forceinline__ void amax_publish(unsigned m, unsigned* slot) {
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31) == 0 && m > *reinterpret_cast<volatile unsigned*>(slot)) atomicMax(slot, m);
}

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }
__device__ __forceinline__ float elu1_grad(float v) { return v > 0.f ? 1.f : expf(v); }

}  // namespace stinet
