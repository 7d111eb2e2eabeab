// Shared helpers for the stinet_b200 kernels (sm_100a only).
#pragma once
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

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }
__device__ __forceinline__ float elu1_grad(float v) { return v > 0.f ? 1.f : expf(v); }

}  // namespace stinet
