// tcgen05 + TMA GEMM for the dense layers (sm_100a only).
//
//   C[i,j] = sum_t A'(i,t) * B'(t,j)      (+ bias[j] on rows with rowmask > 0)
//
// Each operand is either "K-major" (t contiguous in HBM:  A'(i,t) = A[i*ld + t]) or "MN-major" (i / j contiguous:
// A'(i,t) = A[t*ld + i]); that covers  fwd (K,K),  dgrad (K,MN)  and  wgrad (MN,MN, split over t with a fixed-order
// second-stage reduction done by the caller).  Persistent kernel: one CTA per SM walks the 128 x BN output tiles
// (x reduction splits), 10 warps with fixed roles (14 in the one mode that splits its operands in shared memory):
//
//   warp 0    : TMA producer  -- cp.async.bulk.tensor loads 128B-swizzled operand tiles into a kStages-deep smem ring
//   warp 1    : MMA issuer    -- one elected thread issues tcgen05.mma (M=128, N=BN) into one of TWO TMEM accumulator
//                                buffers; tcgen05.commit releases smem stages and hands finished buffers to the epilogue
//   warps 2-5 : (TF32X3 only) in-smem operand split  x = hi + lo  between the TMA and the MMA
//   last 8    : epilogue      -- tcgen05.ld TMEM -> registers; in fp32 mode the partial sums of every 128 reduction
//                                elements are added round-to-nearest into register accumulators (the tensor core's own
//                                fp32 accumulation truncates); then (+bias) -> swizzled smem transpose -> coalesced
//                                128-bit stores.  The epilogue of tile n overlaps the main loop of tile n+1.
//
// Arithmetic modes
//   TF32X3 : fp32 storage, every product evaluated as  hi*hi + hi*lo + lo*hi  on kind::tf32 with fp32 accumulation --
//            error ~2^-22 per product, i.e. fp32-class, which the 1e-5 parity bar against the reference needs;
//   TF32X1 : fp32 storage, one kind::tf32 pass (operands truncated by the tensor core);
//   BF16   : bf16 storage (the caller casts), kind::f16 with fp32 accumulation -- 2e-2 per operator;
//   TF32X3P: as TF32X3, but the caller hands over the hi / lo planes (fp32 storage, split once in HBM): no in-smem
//            split, which costs 96 KB of shared-memory traffic per k-block next to the 96 KB the three MMAs read --
//            the kernel is shared-memory-bandwidth bound, so compute-heavy shapes run ~1.4x faster pre-split;
//   BF16X3 : the caller splits every fp32 operand into two bf16 planes x = hi + lo in HBM; TMA loads all four tiles
//            and the products are evaluated as  hi*hi + hi*lo + lo*hi  on kind::f16 (error ~2^-16 per product) --
//            the mode that keeps a whole deep network within 2e-2 while running bf16 tiles only.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_tc.cuh"

namespace stinet {
namespace tc {

constexpr int BM = 128;
constexpr int kConvWarps = 4;                            // operand-split warps (fp32 mode)
constexpr int kEpiWarps = 8;                             // epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - first) / 4
constexpr uint32_t kRowBytes = 128;  // one swizzle row: 32 fp32 or 64 bf16 along the contiguous dimension
constexpr uint32_t kStagingBytes = kEpiWarps * 32 * 32 * 4;  // one swizzled 32x32 fp32 transpose buffer per epilogue warp

struct TcArgs {
  float* C;
  int64_t ldc;
  const float* bias;
  const int32_t* rowmask;
  int I, J, T;
  int t_per_split;
  int tiles_j, tiles_ij, n_units;  // work units = (split z, row tile, column tile), column tile fastest
  int promote;                     // k-blocks accumulated in TMEM before the partial sum is promoted to registers
  int split_acc;                   // fp32 mode: keep the hi*lo + lo*hi correction terms in their own TMEM accumulator
  const int32_t* a_exp;            // F16 modes: the result is multiplied by 2^(*a_exp + *b_exp) (operand plane scales)
  const int32_t* b_exp;
  unsigned* amax_out;              // nullable: max |C| (after bias) as an fp32 bit pattern, atomicMax into a zeroed slot
};

#ifdef STINET_TC_DEBUG
// cycles CTA 0 spends waiting at each barrier family (build with -DSTINET_TC_DEBUG; read with stinet_tc_debug_read):
// 0 producer<-empty, 1 split<-full, 2 mma<-operands, 3 mma<-tmem_empty, 4 epilogue<-tmem_full, 5 epilogue store
// phase, 6 total kernel cycles, 7 split work, 8 units processed by CTA 0
__device__ unsigned long long g_tc_dbg[16];
#define DBG_T0() const long long dbg_t0 = clock64()
#define DBG_ADD(slot) dbg_acc[slot] += clock64() - dbg_t0
#define DBG_DECL() long long dbg_acc[16] = {0}
#define DBG_FLUSH(cond, lo, hi) \
  if ((cond) && blockIdx.x == 0) { for (int d_ = lo; d_ <= hi; ++d_) g_tc_dbg[d_] = (unsigned long long)dbg_acc[d_]; }
#else
#define DBG_T0()
#define DBG_ADD(slot)
#define DBG_DECL()
#define DBG_FLUSH(cond, lo, hi)
#endif

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

// smem -> global tile store through the tensor map (coordinates: column, row, split plane); rows / columns outside the
// tensor are clipped by the TMA unit.  bulk_group completion: wait_read = the smem source may be overwritten,
// wait_all = the global writes are done.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

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
template <bool K16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (K16) {
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
  static constexpr bool kF16 = MODE == MODE_F16X3 || MODE == MODE_F16X1;      // scaled fp16 planes, result descaled
  static constexpr bool kBf16 = MODE == MODE_BF16 || MODE == MODE_BF16X3 || kF16;  // 16-bit operands on kind::f16
  static constexpr bool kConv = MODE == MODE_TF32X3;               // operand split done in smem by the split warps
  static constexpr bool kPre = MODE == MODE_BF16X3 || MODE == MODE_TF32X3P || MODE == MODE_F16X3;  // split done by the caller in HBM
  static constexpr bool kFp32 = MODE == MODE_TF32X3 || MODE == MODE_TF32X3P || MODE == MODE_F16X3; // fp32-class result: promotion + correction accumulator
  static constexpr bool kSplit = kConv || kPre;                    // stage holds hi and lo tiles, three MMAs per k-step
  // warps: 0 TMA producer, 1 MMA issuer, [2, kEpiWarp0) operand split (only the mode that splits in shared memory has
  // them: registers are handed out per 4 warps, 10 warps leave 170 per thread where 14 leave 128 -- the epilogue's
  // 64 + 32 + 32 live accumulator values then fit without spills), then 8 epilogue warps
  static constexpr int kEpiWarp0 = 2 + (kConv ? kConvWarps : 0);
  static constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
  // F16X3: the B planes sit back to back in the stage ([A_hi][A_lo][B_hi][B_lo]) and the main / correction accumulators
  // back to back in TMEM, so ONE MMA of width 2*BN evaluates A_hi * [B_hi ; B_lo]^T = [hi*hi | hi*lo] and a second one
  // adds A_lo * B_hi^T to the correction half: A_hi is read from shared memory once instead of twice per k-step
  // (20 KB instead of 24 KB of operand reads -- the kernel is shared-memory-bandwidth bound)
  static constexpr bool kStacked = MODE == MODE_F16X3 && BN <= 128;
  static constexpr int kElemBytes = kBf16 ? 2 : 4;
  static constexpr int BKE = kRowBytes / kElemBytes;  // reduction elements per stage: 32 fp32 / 64 bf16
  static constexpr int MNE = kRowBytes / kElemBytes;  // MN elements per swizzle row of an MN-major operand
  static constexpr int UMMA_K = 32 / kElemBytes;      // 8 (tf32) / 16 (bf16)
  static constexpr uint32_t kABytes = BM * kRowBytes;
  static constexpr uint32_t kBBytes = BN * kRowBytes;
  static constexpr uint32_t kLoadBytes = kABytes + kBBytes;
  static constexpr uint32_t kStageBytes = (kSplit ? 2u : 1u) * kLoadBytes;
  static constexpr uint32_t kBarBytes = 8 * (3 * 8 + 4) + 16;
  static constexpr int kStagesRaw = (227 * 1024 - 1024 - (int)kStagingBytes - (int)kBarBytes) / (int)kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  // two accumulator buffers of BN columns; fp32 mode doubles that for the separate correction-term accumulators
  static constexpr uint32_t kTmemCols = (kFp32 ? 4 : 2) * BN;
  static constexpr int kEpiCols = BN / 2;             // accumulator columns owned by one epilogue warp
  // fp32 mode: the tensor core adds into its fp32 accumulator with truncation, so the error of a long reduction
  // grows linearly with K (measured: ~5e-9 * K relative).  Every kPromote k-blocks (128 reduction elements) the
  // partial sum is moved out of TMEM and added, round-to-nearest, to accumulators held in the epilogue warps' registers.
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;
};

template <int BN, bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__((Cfg<BN, MODE>::kThreads), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmC, TcArgs g) {
  using C_ = Cfg<BN, MODE>;
  constexpr bool kBf16 = C_::kBf16;
  constexpr int kStages = C_::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t staging_off = kStages * C_::kStageBytes;
  const uint32_t bar_base = base + staging_off + kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (24 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (26 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 28;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (C_::kPre) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 32 * kConvWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C_::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // decode of a work unit: split z, row tile, column tile -> origin of the output tile and the k-block range
  struct Unit { int i0, j0, z, t_begin, nkb; };
  auto decode = [&](int u) {
    Unit w;
    w.z = u / g.tiles_ij;
    const int r = u - w.z * g.tiles_ij;
    const int ti = r / g.tiles_j;
    w.i0 = ti * BM;
    w.j0 = (r - ti * g.tiles_j) * BN;
    w.t_begin = w.z * g.t_per_split;
    const int t_end = min(g.T, w.t_begin + g.t_per_split);
    w.nkb = (t_end - w.t_begin + C_::BKE - 1) / C_::BKE;
    return w;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int s = 0;
      uint32_t ph = 0;
      DBG_DECL();
      for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
        const Unit w = decode(u);
        for (int kb = 0; kb < w.nkb; ++kb) {
          { DBG_T0(); mbar_wait(empty_bar(s), ph ^ 1u); DBG_ADD(0); }
          mbar_expect_tx(full_bar(s), (C_::kPre ? 2u : 1u) * C_::kLoadBytes);
          const int t0 = w.t_begin + kb * C_::BKE;
#pragma unroll
          for (int plane = 0; plane < (C_::kPre ? 2 : 1); ++plane) {
            const CUtensorMap* mA = plane ? &tmA2 : &tmA;
            const CUtensorMap* mB = plane ? &tmB2 : &tmB;
            const uint32_t a_dst = C_::kStacked ? base + s * C_::kStageBytes + plane * C_::kABytes
                                                : base + s * C_::kStageBytes + plane * C_::kLoadBytes;
            const uint32_t b_dst = C_::kStacked ? base + s * C_::kStageBytes + 2 * C_::kABytes + plane * C_::kBBytes
                                                : a_dst + C_::kABytes;
            if (!A_MN) {
              tma_load_2d(a_dst, mA, t0, w.i0, full_bar(s));
            } else {
#pragma unroll
              for (int b = 0; b < BM / C_::MNE; ++b)
                tma_load_2d(a_dst + b * (C_::BKE * kRowBytes), mA, w.i0 + b * C_::MNE, t0, full_bar(s));
            }
            if (!B_MN) {
              tma_load_2d(b_dst, mB, t0, w.j0, full_bar(s));
            } else {
#pragma unroll
              for (int b = 0; b < BN / C_::MNE; ++b)
                tma_load_2d(b_dst + b * (C_::BKE * kRowBytes), mB, w.j0 + b * C_::MNE, t0, full_bar(s));
            }
          }
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
      DBG_FLUSH(true, 0, 0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor: D fp32, A/B tf32 or bf16, majors, N>>3, M>>4
      const uint32_t fmt = C_::kF16 ? 0u : (kBf16 ? 1u : 2u);   // operand format: f16 / bf16 / tf32
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * BN) >> 3) << 17);   // the stacked MMA: N = 2 * BN
      // per-operand descriptor geometry
      constexpr uint32_t kMnAtomStride = C_::BKE * kRowBytes;              // bytes between MN atoms (one TMA box)
      constexpr uint32_t kMnLayout = kBf16 ? kLayoutSW128 : kLayoutSW128Base32;
      constexpr uint32_t kMnSbo = kBf16 ? 1024u : 512u;                    // bytes between k-groups inside one MMA
      constexpr uint32_t kMnKStep = C_::UMMA_K * kRowBytes;                // k rows consumed per MMA * 128 B
      int s = 0;
      uint32_t ph = 0;
      uint32_t acc_it = 0;
      DBG_DECL();
#ifdef STINET_TC_DEBUG
      const long long dbg_start = clock64();
#endif
      for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
        const Unit w = decode(u);
#ifdef STINET_TC_DEBUG
        dbg_acc[8] += 1;
#endif
        for (int kb0 = 0; kb0 < w.nkb; kb0 += g.promote) {
          const int kb1 = min(w.nkb, kb0 + g.promote);
          const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
          { DBG_T0(); mbar_wait(tempty_bar(buf), aph ^ 1u); DBG_ADD(3); }  // the epilogue has drained this accumulator buffer
          tcgen05_fence_after();
          const uint32_t tmem_d = C_::kStacked ? tmem_base + buf * 2 * BN : tmem_base + buf * BN;
          // correction-term accumulator: right behind the main one (stacked) or in the upper half of the allocation
          const uint32_t tmem_s = C_::kStacked ? tmem_d + BN : (g.split_acc ? tmem_d + 2 * BN : tmem_d);
          for (int kb = kb0; kb < kb1; ++kb) {
            { DBG_T0(); mbar_wait(C_::kConv ? conv_bar(s) : full_bar(s), ph); DBG_ADD(2); }
            tcgen05_fence_after();
            const uint32_t a_hi = base + s * C_::kStageBytes;
            const uint32_t b_hi = C_::kStacked ? a_hi + 2 * C_::kABytes : a_hi + C_::kABytes;
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
              const uint32_t accumulate = (kb > kb0 || k > 0) ? 1u : 0u;
              if (C_::kStacked) {
                const uint32_t a_lo = a_hi + C_::kABytes;
                umma<kBf16>(tmem_d, adesc(a_hi), bdesc(b_hi), idesc2, accumulate);   // [hi*hi | hi*lo] -> [main | corr]
                umma<kBf16>(tmem_s, adesc(a_lo), bdesc(b_hi), idesc, 1u);            // lo*hi -> corr
              } else if (C_::kSplit) {
                const uint32_t a_lo = a_hi + C_::kLoadBytes, b_lo = b_hi + C_::kLoadBytes;
                umma<kBf16>(tmem_s, adesc(a_lo), bdesc(b_hi), idesc, accumulate);
                umma<kBf16>(tmem_s, adesc(a_hi), bdesc(b_lo), idesc, 1u);
                umma<kBf16>(tmem_d, adesc(a_hi), bdesc(b_hi), idesc, g.split_acc ? accumulate : 1u);
              } else {
                umma<kBf16>(tmem_d, adesc(a_hi), bdesc(b_hi), idesc, accumulate);
              }
            }
            umma_commit(empty_bar(s));  // implies tcgen05.fence::before_thread_sync
            if (++s == kStages) { s = 0; ph ^= 1u; }
          }
          umma_commit(tfull_bar(buf));
          ++acc_it;
        }
      }
#ifdef STINET_TC_DEBUG
      dbg_acc[6] = clock64() - dbg_start;
#endif
      DBG_FLUSH(true, 2, 3);
      DBG_FLUSH(true, 6, 6);
      DBG_FLUSH(true, 8, 8);
    }
  } else if (warp < C_::kEpiWarp0) {
    // ===== operand split (fp32 mode): x = hi + lo, hi = x rounded to TF32 (written in place), lo = x - hi (exact in
    // fp32; the tensor core keeps its top 11 significand bits).  Element-wise, so the swizzled layouts are untouched.
    if (C_::kConv) {
      const int ct = threadIdx.x - 64;
      constexpr int kChunks = C_::kLoadBytes / 16;
      int s = 0;
      uint32_t ph = 0;
      DBG_DECL();
      for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
        const Unit w = decode(u);
        for (int kb = 0; kb < w.nkb; ++kb) {
          { DBG_T0(); mbar_wait(full_bar(s), ph); DBG_ADD(1); }
          DBG_T0();
          uint4* hi = reinterpret_cast<uint4*>(base_ptr + (size_t)s * C_::kStageBytes);
          uint4* lo = hi + kChunks;
#pragma unroll 4
          for (int idx = ct; idx < kChunks; idx += 32 * kConvWarps) {
            const uint4 v = hi[idx];
            uint4 h, l;
#ifdef STINET_TC_TRUNC_HI
            // EXPERIMENT (make -C csrc trunc; scripts/exp_trunc_hi.sh): leave the fp32 tile as TMA wrote it and let the
            // tensor core drop the low 13 bits itself, i.e. hi = trunc(x); only lo is written (64 instead of 96 KB of
            // split traffic per k-block).  The numpy model (tests/test_tf32x3_model.py) puts the error at 7e-7 .. 9e-7;
            // whether kind::tf32 really truncates its operands (rather than rounding) has to be seen on hardware.
            h.x = v.x & 0xFFFFE000u; l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
            h.y = v.y & 0xFFFFE000u; l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
            h.z = v.z & 0xFFFFE000u; l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
            h.w = v.w & 0xFFFFE000u; l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
            lo[idx] = l;
#else
            h.x = (v.x + 0x1000u) & 0xFFFFE000u; l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
            h.y = (v.y + 0x1000u) & 0xFFFFE000u; l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
            h.z = (v.z + 0x1000u) & 0xFFFFE000u; l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
            h.w = (v.w + 0x1000u) & 0xFFFFE000u; l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
            hi[idx] = h;
            lo[idx] = l;
#endif
          }
          fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
          mbar_arrive(conv_bar(s));
          DBG_ADD(7);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
      DBG_FLUSH(threadIdx.x == 64, 1, 1);
      DBG_FLUSH(threadIdx.x == 64, 7, 7);
    }
  } else {
    // ===== epilogue: TMEM -> registers (RN accumulation across promotion chunks) -> swizzled smem transpose ->
    // coalesced 128-bit global stores =====
    const int ew = warp - C_::kEpiWarp0;
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int h = ew >> 2;             // column half
    constexpr int EC = C_::kEpiCols;
    float4* stage = reinterpret_cast<float4*>(base_ptr + staging_off + (size_t)ew * 4096);
    float acc[EC];
    uint32_t acc_it = 0;
    unsigned amax_bits = 0u;           // max |C| over every tile of this warp, published once at the end
    DBG_DECL();
    for (int u = blockIdx.x; u < g.n_units; u += gridDim.x) {
      const Unit w = decode(u);
      for (int kb0 = 0; kb0 < w.nkb; kb0 += g.promote) {
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        { DBG_T0(); mbar_wait(tfull_bar(buf), aph); DBG_ADD(4); }
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (C_::kStacked ? 2 * BN : BN) + (uint32_t)(h * EC);
#pragma unroll
        for (int cc = 0; cc < EC; cc += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + cc, r);
          if (C_::kFp32 && g.split_acc) {
            uint32_t r2[32];
            tmem_ld32(taddr + (C_::kStacked ? BN : 2 * BN) + cc, r2);
#pragma unroll
            for (int c = 0; c < 32; ++c)   // F16X3: the lo planes carry a factor 2^11
              r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(r2[c]) * (C_::kF16 ? (1.f / 2048.f) : 1.f));
          }
          if (kb0 == 0) {
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[cc + c] = __uint_as_float(r[c]);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[cc + c] += __uint_as_float(r[c]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));
        ++acc_it;
      }
      // ---- write the tile out: (+bias) -> this warp's swizzled 32x32 staging block -> one TMA store per block
      DBG_T0();
      const int row0 = w.i0 + q * 32;                      // first row of this warp's lane quarter
      const int row = row0 + lane;                         // the row this thread holds
      const bool add_bias = g.bias != nullptr && row < g.I && (g.rowmask == nullptr || g.rowmask[row] > 0);
      if (C_::kF16) {
        // undo the operand scales: exact powers of two, applied in two factors so that neither leaves the fp32 range
        const int t = __ldg(g.a_exp) + __ldg(g.b_exp);
        const int t1 = t / 2;
        const float f1 = __uint_as_float((uint32_t)(127 + max(-126, min(127, t1))) << 23);
        const float f2 = __uint_as_float((uint32_t)(127 + max(-126, min(127, t - t1))) << 23);
#pragma unroll
        for (int c = 0; c < EC; ++c) acc[c] = acc[c] * f1 * f2;
      }
      if (row0 < g.I) {
#pragma unroll
        for (int cc = 0; cc < EC; cc += 32) {
          const int jb = w.j0 + h * EC + cc;               // first global column of this 32-column block
          if (jb >= g.J) break;
          if (lane == 0) bulk_wait_read();                 // the previous block has left the staging buffer
          __syncwarp();
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 v = make_float4(acc[cc + 4 * c4], acc[cc + 4 * c4 + 1], acc[cc + 4 * c4 + 2], acc[cc + 4 * c4 + 3]);
            if (add_bias && jb + 4 * c4 < g.J) {
              const float4 b = *reinterpret_cast<const float4*>(g.bias + jb + 4 * c4);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            amax_bits = max(max(amax_bits, __float_as_uint(v.x) & 0x7FFFFFFFu), __float_as_uint(v.y) & 0x7FFFFFFFu);
            amax_bits = max(max(amax_bits, __float_as_uint(v.z) & 0x7FFFFFFFu), __float_as_uint(v.w) & 0x7FFFFFFFu);
            stage[lane * 8 + (c4 ^ (lane & 7))] = v;       // the 128-byte swizzle of the tensor map, by hand
          }
          fence_proxy_async();                             // generic-proxy smem writes -> visible to the TMA unit
          __syncwarp();
          if (lane == 0) tma_store_3d(&tmC, smem_u32(stage), jb, row0, w.z);
        }
      }
      DBG_ADD(5);
    }
    if (g.amax_out != nullptr) {
      // rows / columns outside the matrix hold exact zeros (TMA zero fill, no bias), so they cannot raise the maximum
      amax_bits = __reduce_max_sync(0xFFFFFFFFu, amax_bits);
      if (lane == 0 && amax_bits > *reinterpret_cast<volatile unsigned*>(g.amax_out)) atomicMax(g.amax_out, amax_bits);
    }
    if (lane == 0) bulk_wait_all();
    DBG_FLUSH(threadIdx.x == 32 * C_::kEpiWarp0, 4, 5);
  }
  tcgen05_fence_before();
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
static int make_map(CUtensorMap* m, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int bf16,
                    uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {   // bf16: 0 fp32, 1 bf16, 2 fp16
  EncodeTiledFn fn = encode_fn();
  STINET_REQUIRE(fn != nullptr, STINET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t esz = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, bf16 == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  STINET_REQUIRE(r == CUDA_SUCCESS, STINET_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner %lld outer %lld ld %lld",
                 (int)r, (long long)inner, (long long)outer, (long long)ld);
  return STINET_OK;
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static int launch(const Problem& p, cudaStream_t s) {
  using C_ = Cfg<BN, MODE>;
  constexpr int bf16 = C_::kF16 ? 2 : (C_::kBf16 ? 1 : 0);
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
  CUtensorMap tmA2 = tmA, tmB2 = tmB;
  if (C_::kPre) {
    // the lo planes have the element type, pitch and layout of the hi planes
    STINET_REQUIRE(p.A_lo && p.B_lo, STINET_ERR_ARG, "gemm_tc: a pre-split mode needs the lo planes of both operands");
    const CUtensorMapSwizzle mn_swz = bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if (!A_MN) rc = make_map(&tmA2, p.A_lo, p.T, p.I, p.lda, bf16, C_::BKE, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    else rc = make_map(&tmA2, p.A_lo, p.I, p.T, p.lda, bf16, C_::MNE, C_::BKE, mn_swz);
    if (rc) return rc;
    if (!B_MN) rc = make_map(&tmB2, p.B_lo, p.T, p.J, p.ldb, bf16, C_::BKE, BN, CU_TENSOR_MAP_SWIZZLE_128B);
    else rc = make_map(&tmB2, p.B_lo, p.J, p.T, p.ldb, bf16, C_::MNE, C_::BKE, mn_swz);
    if (rc) return rc;
  }
  // output (or split-K partial planes [splits, I, J]) as a 3-D tensor: 32 x 32 fp32 boxes, 128-byte swizzle
  CUtensorMap tmC;
  {
    EncodeTiledFn fn = encode_fn();
    STINET_REQUIRE(fn != nullptr, STINET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {(cuuint64_t)p.J, (cuuint64_t)p.I, (cuuint64_t)p.splits};
    cuuint64_t strides[2] = {(cuuint64_t)p.ldc * 4, (cuuint64_t)p.I * (cuuint64_t)p.ldc * 4};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    STINET_REQUIRE(r == CUDA_SUCCESS, STINET_ERR_CUDA, "cuTensorMapEncodeTiled (C) failed (%d): I %lld J %lld ldc %lld",
                   (int)r, (long long)p.I, (long long)p.J, (long long)p.ldc);
  }
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, MODE>;
  static bool attr_set = false;  // per instantiation; idempotent, so a race only repeats the call
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::kSmemBytes);
    STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles_j = (int)ceil_div(p.J, BN), tiles_i = (int)ceil_div(p.I, BM);
  const int64_t units = (int64_t)tiles_i * tiles_j * p.splits;
  STINET_REQUIRE(units < (1ll << 31), STINET_ERR_UNSUPPORTED, "gemm_tc: too many tiles");
  // fp32 mode: promote every 4 k-blocks (128 reduction elements) and keep the correction terms in their own
  // accumulator.  Measured on B200 against fp64 (scripts/exp_promote.sh): 2.6e-7 .. 6.5e-7 max-norm relative error,
  // independent of K and at no cost in time (FFMA tiles: 2.7e-7 .. 1.4e-6; one shared accumulator, no promotion: up
  // to 3e-5 at K = 4096).  STINET_TC_PROMOTE / STINET_TC_SPLITACC override the two knobs for that experiment.
  static const int env_promote = [] { const char* e = getenv("STINET_TC_PROMOTE"); return e ? atoi(e) : 4; }();
  static const int env_split = [] { const char* e = getenv("STINET_TC_SPLITACC"); return e ? atoi(e) : 1; }();
  // 16-bit k-blocks hold 64 reduction elements: promote every 2 of them; the F16X3 correction accumulator is not optional
  // (its terms carry the factor 2^11 of the lo planes)
  static const int env_promote16 = [] { const char* e = getenv("STINET_TC_PROMOTE16"); return e ? atoi(e) : 2; }();
  const int promote = !C_::kFp32 ? (1 << 28) : C_::kBf16 ? (env_promote16 > 0 ? env_promote16 : 2) : (env_promote > 0 ? env_promote : 4);
  if (C_::kF16) STINET_REQUIRE(p.a_exp && p.b_exp, STINET_ERR_ARG, "gemm_tc: the fp16 modes need the operand scale exponents");
  TcArgs g{p.C, p.ldc, p.bias, p.rowmask, (int)p.I, (int)p.J, (int)p.T, (int)p.t_per_split,
           tiles_j, tiles_i * tiles_j, (int)units,
           promote, C_::kFp32 ? (C_::kF16 ? 1 : env_split) : 0, p.a_exp, p.b_exp, p.splits == 1 ? p.amax_out : nullptr};
  // persistent: one CTA per SM walks the work units.  STINET_TC_SMS < 148 leaves SMs to kernels that run beside the GEMMs
  // (NCCL's all-reduce CTAs during an overlapped backward: a persistent grid that cannot become fully resident at once
  // pays a second wave for its last CTAs)
  static const int env_sms = [] { const char* e = getenv("STINET_TC_SMS"); int v = e ? atoi(e) : kSMs; return v >= 1 && v <= kSMs ? v : kSMs; }();
  const unsigned grid = (unsigned)(units < env_sms ? units : env_sms);
  K(kern<<<grid, C_::kThreads, C_::kSmemBytes, s>>>(tmA, tmB, tmA2, tmB2, tmC, g));
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
  const int esz = (p.mode == MODE_BF16 || p.mode == MODE_BF16X3 || p.mode == MODE_F16X3 || p.mode == MODE_F16X1) ? 2 : 4;
  const int64_t align_elems = 16 / esz;
  if (p.I <= 0 || p.J <= 0 || p.T <= 0) return false;
  if (!aligned16(p.A) || !aligned16(p.B) || !aligned16(p.C)) return false;
  if ((p.mode == MODE_BF16X3 || p.mode == MODE_TF32X3P || p.mode == MODE_F16X3) &&
      (!p.A_lo || !p.B_lo || !aligned16(p.A_lo) || !aligned16(p.B_lo))) return false;
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
    case MODE_BF16X3: return launch_major<MODE_BF16X3>(p, s);
    case MODE_TF32X3P: return launch_major<MODE_TF32X3P>(p, s);
    case MODE_F16X3: return launch_major<MODE_F16X3>(p, s);
    case MODE_F16X1: return launch_major<MODE_F16X1>(p, s);
  }
  set_error("gemm_tc: unknown mode %d", p.mode);
  return STINET_ERR_ARG;
}

}  // namespace tc
}  // namespace stinet

#ifdef STINET_TC_DEBUG
extern "C" int stinet_tc_debug_read(unsigned long long* out16) {
  return cudaMemcpyFromSymbol(out16, stinet::tc::g_tc_dbg, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -3;
}
#endif
