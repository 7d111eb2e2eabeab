// Interface between the dense-layer entry points (gemm.cu) and the tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stinet {
namespace tc {

enum { MODE_TF32X3 = 0, MODE_TF32X1 = 1, MODE_BF16 = 2, MODE_BF16X3 = 3, MODE_TF32X3P = 4, MODE_F16X3 = 5, MODE_F16X1 = 6 };

// C[i,j] = sum_t A'(i,t) B'(t,j);  a_mn: A'(i,t) = A[t*lda + i] (else A[i*lda + t]); same for B with j.
// A and B are fp32 (TF32 modes) or bf16 (MODE_BF16, MODE_BF16X3); C is fp32.  MODE_BF16X3: every operand is given as
// two bf16 planes, x = hi + lo (A/B = hi, A_lo/B_lo = lo, same pitch), and the kernel evaluates hi*hi + hi*lo + lo*hi.
// MODE_TF32X3P: the same with two fp32 planes (hi = x rounded to TF32, lo = x - hi).
// MODE_F16X3 (the fp32-parity mode): every operand is given as two fp16 planes of the SCALED matrix x * 2^s (s chosen
// from the operand's amax so that the largest element sits just below 2^15): hi = fp16(x 2^s), lo = fp16((x 2^s - hi) 2^11);
// hi + lo 2^-11 carries 22 significand bits.  The kernel evaluates hi*hi in one TMEM accumulator and hi*lo + lo*hi in a
// second one on kind::f16 (the full 16-bit tensor rate, twice kind::tf32), promotes partial sums to round-to-nearest
// registers every 128 reduction elements and multiplies the result by 2^(a_exp + b_exp), a_exp = -s_A, b_exp = -s_B read
// from device memory (the planes' producers write them).  MODE_F16X1: the hi planes only (one pass, 11 significand bits).
// With splits > 1, split z covers t in [z*t_per_split, (z+1)*t_per_split) and writes its partial to C + z*I*ldc.
struct Problem {
  const void* A; int64_t lda; bool a_mn;
  const void* B; int64_t ldb; bool b_mn;
  float* C; int64_t ldc;
  const float* bias; const int32_t* rowmask;
  int64_t I, J, T;
  int splits; int64_t t_per_split;
  int mode;
  const void* A_lo = nullptr;
  const void* B_lo = nullptr;
  const int32_t* a_exp = nullptr;   // F16 modes: device scalars, result *= 2^(*a_exp + *b_exp)
  const int32_t* b_exp = nullptr;
  unsigned* amax_out = nullptr;     // nullable, unsplit launches only: atomicMax of |C| bit patterns into a zeroed slot
};

bool eligible(const Problem& p);              // TMA alignment / size rules
int run(const Problem& p, cudaStream_t s);    // STINET_OK or an error code (message via set_error)

}  // namespace tc
}  // namespace stinet
