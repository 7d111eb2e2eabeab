// tcgen05 + TMA GEMM for the dense layers (sm_100a only).
//
//   C[i,j] = sum_t A'(i,t) * B'(t,j)      (+ bias[j] on rows with rowmask > 0)
//
// Each operand is either "K-major" (t contiguous in HBM:  A'(i,t) = A[i*ld + t]) or "MN-major" (i / j contiguous:
// A'(i,t) = A[t*ld + i]); that covers  fwd (K,K),  dgrad (K,MN)  and  wgrad (MN,MN, split over t with a fixed-order
// second-stage reduction done by the caller).  One CTA computes one 128 x BN output tile:
//
//   warp 0   : TMA producer   -- cp.async.bulk.tensor loads 128B-swizzled operand tiles into a kStages-deep smem ring
//   warp 1   : MMA issuer     -- one elected thread issues tcgen05.mma (M=128, N=BN), accumulator in TMEM;
//                                tcgen05.commit releases smem stages and finally signals the epilogue
//   warps 2-5: (fp32 mode) in-smem operand split  x = hi + lo  (both rounded to TF32), then the epilogue:
//              tcgen05.ld TMEM -> registers -> (+bias) -> global
//
// Arithmetic modes
//   TF32X3 : fp32 storage, every product evaluated as  hi*hi + hi*lo + lo*hi  on kind::tf32 with fp32 accumulation --
//            error ~2^-22 per product, i.e. fp32-class, which the 1e-5 parity bar against the reference needs;
//   TF32X1 : fp32 storage, one kind::tf32 pass (operands truncated by the tensor core);
//   BF16   : bf16 storage (the caller casts), kind::f16 with fp32 accumulation -- the 2e-2 bar of the bf16 mode.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "gemm_tc.cuh"

namespace stinet {
namespace tc {

constexpr int BM = 128;
constexpr int kThreads = 192;
constexpr uint32_t kRowBytes = 128;  // one swizzle row: 32 fp32 or 64 bf16 along the contiguous dimension

struct TcArgs {
  float* C;
  int64_t ldc;
  const float* bias;
  const int32_t* rowmask;
  int I, J, T;
  int t_per_split;
};

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a pipeline bug must surface as a launch failure (trap), never as a hung device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && globaltimer_ns() - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// 32 lanes x 32 consecutive fp32 columns of the accumulator: thread `lane` of the warp gets its row's 32 values
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor (sm_100 format): start address, leading / stride byte offsets (16 B units),
// version 1, swizzle layout type in bits [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
constexpr uint32_t kLayoutSW128 = 2;        // Swizzle<3,4,3>: 16 B chunks, pattern repeats every 8 rows
constexpr uint32_t kLayoutSW128Base32 = 1;  // Swizzle<2,5,2>: 32 B chunks, every 4 rows (MN-major tf32 operands)

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// ---------------------------------------------------------------------------------------------------------------

template <int BN, int MODE>
struct Cfg {
  static constexpr bool kBf16 = MODE == MODE_BF16;
  static constexpr bool kSplit = MODE == MODE_TF32X3;
  static constexpr int kElemBytes = kBf16 ? 2 : 4;
  static constexpr int BKE = kRowBytes / kElemBytes;  // reduction elements per stage: 32 fp32 / 64 bf16
  static constexpr int MNE = kRowBytes / kElemBytes;  // MN elements per swizzle row of an MN-major operand
  static constexpr int UMMA_K = 32 / kElemBytes;      // 8 (tf32) / 16 (bf16)
  static constexpr uint32_t kABytes = BM * kRowBytes;
  static constexpr uint32_t kBBytes = BN * kRowBytes;
  static constexpr uint32_t kLoadBytes = kABytes + kBBytes;
  static constexpr uint32_t kStageBytes = (kSplit ? 2u : 1u) * kLoadBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;  // BN is 64 / 128 / 256: already a power of two
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 8 * (3 * kStages + 1) + 16 + 1024;
};

template <int BN, bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcArgs g) {
  using C_ = Cfg<BN, MODE>;
  constexpr bool kBf16 = C_::kBf16;
  constexpr int kStages = C_::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bar_base = base + kStages * C_::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * kStages);
  const uint32_t tmem_slot = accum_bar + 8u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * BN, i0 = blockIdx.y * BM;
  const int t_begin = blockIdx.z * g.t_per_split;
  const int t_end = min(g.T, t_begin + g.t_per_split);
  const int nkb = (t_end - t_begin + C_::BKE - 1) / C_::BKE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C_::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), C_::kLoadBytes);
        const int t0 = t_begin + kb * C_::BKE;
        const uint32_t a_dst = base + s * C_::kStageBytes;
        const uint32_t b_dst = a_dst + C_::kABytes;
        if (!A_MN) {
          tma_load_2d(a_dst, &tmA, t0, i0, full_bar(s));
        } else {
#pragma unroll
          for (int b = 0; b < BM / C_::MNE; ++b)
            tma_load_2d(a_dst + b * (C_::BKE * kRowBytes), &tmA, i0 + b * C_::MNE, t0, full_bar(s));
        }
        if (!B_MN) {
          tma_load_2d(b_dst, &tmB, t0, j0, full_bar(s));
        } else {
#pragma unroll
          for (int b = 0; b < BN / C_::MNE; ++b)
            tma_load_2d(b_dst + b * (C_::BKE * kRowBytes), &tmB, j0 + b * C_::MNE, t0, full_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor: D fp32, A/B tf32 or bf16, majors, N>>3, M>>4
      const uint32_t fmt = kBf16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // per-operand descriptor geometry
      constexpr uint32_t kMnAtomStride = C_::BKE * kRowBytes;              // bytes between MN atoms (one TMA box)
      constexpr uint32_t kMnLayout = kBf16 ? kLayoutSW128 : kLayoutSW128Base32;
      constexpr uint32_t kMnSbo = kBf16 ? 1024u : 512u;                    // bytes between k-groups inside one MMA
      constexpr uint32_t kMnKStep = C_::UMMA_K * kRowBytes;                // k rows consumed per MMA * 128 B
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(C_::kSplit ? conv_bar(s) : full_bar(s), ph);
        tcgen05_fence_after();
        const uint32_t a_hi = base + s * C_::kStageBytes;
        const uint32_t b_hi = a_hi + C_::kABytes;
#pragma unroll
        for (int k = 0; k < C_::BKE / C_::UMMA_K; ++k) {
          const uint32_t a_off = A_MN ? k * kMnKStep : k * 32u;
          const uint32_t b_off = B_MN ? k * kMnKStep : k * 32u;
          auto adesc = [&](uint32_t addr) {
            return A_MN ? make_smem_desc(addr + a_off, kMnAtomStride, kMnSbo, kMnLayout)
                        : make_smem_desc(addr + a_off, 16u, 1024u, kLayoutSW128);
          };
          auto bdesc = [&](uint32_t addr) {
            return B_MN ? make_smem_desc(addr + b_off, kMnAtomStride, kMnSbo, kMnLayout)
                        : make_smem_desc(addr + b_off, 16u, 1024u, kLayoutSW128);
          };
          const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
          if (C_::kSplit) {
            const uint32_t a_lo = a_hi + C_::kLoadBytes, b_lo = b_hi + C_::kLoadBytes;
            umma<false>(tmem_base, adesc(a_lo), bdesc(b_hi), idesc, first);
            umma<false>(tmem_base, adesc(a_hi), bdesc(b_lo), idesc, 1u);
            umma<false>(tmem_base, adesc(a_hi), bdesc(b_hi), idesc, 1u);
          } else {
            umma<kBf16>(tmem_base, adesc(a_hi), bdesc(b_hi), idesc, first);
          }
        }
        umma_commit(empty_bar(s));  // implies tcgen05.fence::before_thread_sync
      }
      umma_commit(accum_bar);
    }
  } else {
    // ===== warps 2..5: operand split (fp32 mode), then epilogue =====
    const int ct = threadIdx.x - 64;
    if (C_::kSplit) {
      constexpr int kChunks = C_::kLoadBytes / 16;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(full_bar(s), ph);
        uint4* hi = reinterpret_cast<uint4*>(base_ptr + (size_t)s * C_::kStageBytes);
        uint4* lo = hi + kChunks;
#pragma unroll 4
        for (int idx = ct; idx < kChunks; idx += 128) {
          const uint4 v = hi[idx];
          uint4 h, l;
          h.x = to_tf32(__uint_as_float(v.x)); l.x = to_tf32(__uint_as_float(v.x) - __uint_as_float(h.x));
          h.y = to_tf32(__uint_as_float(v.y)); l.y = to_tf32(__uint_as_float(v.y) - __uint_as_float(h.y));
          h.z = to_tf32(__uint_as_float(v.z)); l.z = to_tf32(__uint_as_float(v.z) - __uint_as_float(h.z));
          h.w = to_tf32(__uint_as_float(v.w)); l.w = to_tf32(__uint_as_float(v.w) - __uint_as_float(h.w));
          hi[idx] = h;
          lo[idx] = l;
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(conv_bar(s));
      }
    }
    mbar_wait(accum_bar, 0);
    tcgen05_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = i0 + q * 32 + lane;
    const bool row_ok = row < g.I;
    const bool add_bias = g.bias != nullptr && row_ok && (g.rowmask == nullptr || g.rowmask[row] > 0);
    float* crow = g.C + (int64_t)blockIdx.z * g.I * g.ldc + (int64_t)row * g.ldc;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      if (row_ok) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const int j = j0 + c0 + c;
          if (j < g.J) {
            float4 v = make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]),
                                   __uint_as_float(r[c + 3]));
            if (add_bias) {
              const float4 b = *reinterpret_cast<const float4*>(g.bias + j);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            *reinterpret_cast<float4*>(crow + j) = v;
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C_::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D tensor map over a row-major matrix [outer, inner] with row pitch ld (elements); box = [box_outer, box_inner].
static int make_map(CUtensorMap* m, const void* ptr, int64_t inner, int64_t outer, int64_t ld, bool bf16,
                    uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  STINET_REQUIRE(fn != nullptr, STINET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t esz = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  STINET_REQUIRE(r == CUDA_SUCCESS, STINET_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner %lld outer %lld ld %lld",
                 (int)r, (long long)inner, (long long)outer, (long long)ld);
  return STINET_OK;
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static int launch(const Problem& p, cudaStream_t s) {
  using C_ = Cfg<BN, MODE>;
  constexpr bool bf16 = C_::kBf16;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_map(&tmA, p.A, p.T, p.I, p.lda, bf16, C_::BKE, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  else rc = make_map(&tmA, p.A, p.I, p.T, p.lda, bf16, C_::MNE, C_::BKE,
                     bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  if (!B_MN) rc = make_map(&tmB, p.B, p.T, p.J, p.ldb, bf16, C_::BKE, BN, CU_TENSOR_MAP_SWIZZLE_128B);
  else rc = make_map(&tmB, p.B, p.J, p.T, p.ldb, bf16, C_::MNE, C_::BKE,
                     bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, MODE>;
  static bool attr_set = false;  // per instantiation; idempotent, so a race only repeats the call
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::kSmemBytes);
    STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  TcArgs g{p.C, p.ldc, p.bias, p.rowmask, (int)p.I, (int)p.J, (int)p.T, (int)p.t_per_split};
  dim3 grid((unsigned)ceil_div(p.J, BN), (unsigned)ceil_div(p.I, BM), (unsigned)p.splits);
  K(kern<<<grid, kThreads, C_::kSmemBytes, s>>>(tmA, tmB, g));
  return check_launch("gemm_tc");
}

template <bool A_MN, bool B_MN, int MODE>
static int launch_bn(const Problem& p, cudaStream_t s) {
  if (p.J <= 64) return launch<64, A_MN, B_MN, MODE>(p, s);
  return launch<128, A_MN, B_MN, MODE>(p, s);
}

template <int MODE>
static int launch_major(const Problem& p, cudaStream_t s) {
  if (!p.a_mn && !p.b_mn) return launch_bn<false, false, MODE>(p, s);
  if (!p.a_mn && p.b_mn) return launch_bn<false, true, MODE>(p, s);
  if (p.a_mn && p.b_mn) return launch_bn<true, true, MODE>(p, s);
  set_error("gemm_tc: (MN,K) operand combination is not instantiated");
  return STINET_ERR_UNSUPPORTED;
}

bool eligible(const Problem& p) {
  const int esz = p.mode == MODE_BF16 ? 2 : 4;
  const int64_t align_elems = 16 / esz;
  if (p.I <= 0 || p.J <= 0 || p.T <= 0) return false;
  if (!aligned16(p.A) || !aligned16(p.B) || !aligned16(p.C)) return false;
  if (p.lda % align_elems || p.ldb % align_elems) return false;
  if (p.ldc % 4 || p.J % 4) return false;
  if (p.bias && !aligned16(p.bias)) return false;
  if (p.I >= (1ll << 31) || p.J >= (1ll << 31) || p.T >= (1ll << 31)) return false;
  return true;
}

int run(const Problem& p, cudaStream_t s) {
  STINET_REQUIRE(eligible(p), STINET_ERR_UNSUPPORTED, "gemm_tc: operands not TMA-eligible");
  STINET_REQUIRE(p.splits >= 1 && (p.splits == 1 || p.t_per_split % 64 == 0), STINET_ERR_ARG, "gemm_tc: bad split");
  switch (p.mode) {
    case MODE_TF32X3: return launch_major<MODE_TF32X3>(p, s);
    case MODE_TF32X1: return launch_major<MODE_TF32X1>(p, s);
    case MODE_BF16: return launch_major<MODE_BF16>(p, s);
  }
  set_error("gemm_tc: unknown mode %d", p.mode);
  return STINET_ERR_ARG;
}

}  // namespace tc
}  // namespace stinet
