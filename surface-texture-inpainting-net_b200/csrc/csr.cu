// CSR builder: stable grouping of positions by an int64 key (edge target / edge source / trace id).
//   count (atomic histogram, order-independent) -> exclusive scan -> atomic fill -> per-row sort by position.
// The final per-row sort makes the result independent of atomic arrival order: rows list their members in
// ascending original position, i.e. exactly `argsort(key, stable=True)` (oracle: csr_by_key).
#include "common.cuh"

namespace stinet {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;
constexpr int kSmallRow = 64;  // rows up to this length are sorted by one thread

__global__ void csr_count_kernel(const int64_t* __restrict__ key, int64_t n_items, int64_t n_rows,
                                 int32_t* __restrict__ cnt, int32_t* __restrict__ key32,
                                 int32_t* __restrict__ status) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_items;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = key[e];
    bool ok = (t >= 0) & (t < n_rows);
    if (key32) key32[e] = ok ? (int32_t)t : 0;
    if (ok) {
      atomicAdd(&cnt[t], 1);
    } else if (status) {
      atomicOr(status, 1);
    }
  }
}

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// inclusive scan of one value per thread across the block; returns the inclusive prefix, *total = block sum
__device__ __forceinline__ int block_incl_scan(int v, int* total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0;
    w = warp_incl_scan(w);
    wsum[lane] = w;
  }
  __syncthreads();
  int off = wid ? wsum[wid - 1] : 0;
  *total = wsum[(blockDim.x >> 5) - 1];
  __syncthreads();
  return inc + off;
}

__global__ void __launch_bounds__(kScanThreads) scan_chunk_sums(const int32_t* __restrict__ cnt, int64_t n,
                                                                 int32_t* __restrict__ chunk_sum) {
  int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) s += cnt[base + k];
  int total;
  block_incl_scan(s, &total);
  if (threadIdx.x == 0) chunk_sum[blockIdx.x] = total;
}

// rowptr[i] = sum_{k<i} cnt[k] for i in [0, n_rows]  (n = n_rows + 1 outputs; cnt[n_rows] is read as 0)
// (the exclusive offset of a chunk = sum of the chunk sums before it, added up by the CTA itself in a fixed order: a few
// hundred values at most, so no separate offsets kernel)
__global__ void __launch_bounds__(kScanThreads) scan_write(const int32_t* __restrict__ cnt, int64_t n_rows,
                                                           const int32_t* __restrict__ chunk_sum,
                                                           int32_t* __restrict__ rowptr) {
  __shared__ int chunk_off_s;
  {
    int part = 0;
    for (int c = threadIdx.x; c < (int)blockIdx.x; c += kScanThreads) part += chunk_sum[c];
    int total;
    block_incl_scan(part, &total);
    if (threadIdx.x == 0) chunk_off_s = total;
    __syncthreads();
  }
  int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n_rows) ? cnt[base + k] : 0;
    s += v[k];
  }
  int total;
  int inc = block_incl_scan(s, &total);
  int run = chunk_off_s + inc - s;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k <= n_rows) rowptr[base + k] = run;
    run += v[k];
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ key, int64_t n_items, int64_t n_rows,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ perm) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_items;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = key[e];
    if (t < 0 || t >= n_rows) continue;
    int pos = atomicAdd(&cursor[t], 1);
    perm[rowptr[t] + pos] = (int32_t)e;
  }
}

// one thread per row: insertion sort of short rows; long rows are queued for the block-level rank sort
__global__ void csr_sort_small_rows(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t* __restrict__ perm,
                                    int32_t* __restrict__ big_rows, int32_t* __restrict__ n_big) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    int beg = rowptr[r], end = rowptr[r + 1];
    int d = end - beg;
    if (d <= 1) continue;
    if (d > kSmallRow) {
      big_rows[atomicAdd(n_big, 1)] = (int32_t)r;
      continue;
    }
    for (int a = beg + 1; a < end; ++a) {
      int v = perm[a];
      int b = a - 1;
      while (b >= beg && perm[b] > v) {
        perm[b + 1] = perm[b];
        --b;
      }
      perm[b + 1] = v;
    }
  }
}

// rank sort of long rows (positions are distinct, so ranks are a permutation): tmp[beg + rank] = value
__global__ void csr_rank_big_rows(const int32_t* __restrict__ rowptr, int32_t* __restrict__ perm,
                                  const int32_t* __restrict__ big_rows, const int32_t* __restrict__ n_big,
                                  int32_t* __restrict__ tmp) {
  const int nb = *n_big;
  for (int q = blockIdx.x; q < nb; q += gridDim.x) {
    int r = big_rows[q];
    int beg = rowptr[r], end = rowptr[r + 1];
    for (int a = beg + threadIdx.x; a < end; a += blockDim.x) {
      int v = perm[a];
      int rank = 0;
      for (int b = beg; b < end; ++b) rank += (perm[b] < v);
      tmp[beg + rank] = v;
    }
    __syncthreads();                                   // the row is owned by this CTA: ranks done, copy back
    for (int a = beg + threadIdx.x; a < end; a += blockDim.x) perm[a] = tmp[a];
    __syncthreads();
  }
}
// Out-of-range keys are dropped by count / fill (status bit 0), so only the first rowptr[n_rows] entries of perm are
// written: the tail is filled here (perm = -1, col = 0; every later pass over perm skips negative entries) instead of being
// left to chance.
__global__ void csr_gather_col(const int64_t* __restrict__ other, int32_t* __restrict__ perm, int64_t n_items,
                               int64_t n_rows, const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                               int32_t* __restrict__ status) {
  const int64_t n_valid = rowptr[n_rows];
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_items;
       k += (int64_t)gridDim.x * blockDim.x) {
    if (k >= n_valid) {
      perm[k] = -1;
      if (col) col[k] = 0;
      continue;
    }
    if (!col) continue;
    int64_t v = other[perm[k]];
    bool ok = (v >= 0) & (v < n_rows);
    col[k] = ok ? (int32_t)v : 0;
    if (!ok && status) atomicOr(status, 1);
  }
}

// tpos_s[k_s] = position, in the by-target CSR, of the edge that sits at position k_s of the by-source CSR
// (both are permutations of the same original edge ids): inv[eid_t[k]] = k, then tpos_s[k_s] = inv[eid_s[k_s]].
__global__ void invert_perm_kernel(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t e = perm[k];
    if (e >= 0 && e < n) inv[e] = (int32_t)k;          // the tail left by out-of-range keys is marked -1
  }
}
__global__ void gather_i32_kernel(const int32_t* __restrict__ table, const int32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = idx[k];
    out[k] = i >= 0 ? table[i] : 0;
  }
}

// Block-diagonal batching of per-sample structures: part p of the destination is src[p][0..len) + add[p].
// Up to kConcatParts parts travel by value in the launch parameters (no device-side pointer table, no H2D copy).
constexpr int kConcatParts = 32;
struct ConcatArgs {
  const int32_t* src[kConcatParts];
  int64_t len[kConcatParts];
  int64_t dst_off[kConcatParts];
  int32_t add[kConcatParts];
  int n_parts;
};
__global__ void __launch_bounds__(256) concat_i32_kernel(const __grid_constant__ ConcatArgs a, int32_t* __restrict__ dst) {
  const int p = blockIdx.y;
  const int32_t* __restrict__ src = a.src[p];
  int32_t* __restrict__ out = dst + a.dst_off[p];
  const int64_t n = a.len[p];
  const int32_t add = a.add[p];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  // 128-bit body when both sides are 16-byte aligned, scalar otherwise (offsets are arbitrary vertex / edge counts)
  if ((((uintptr_t)src | (uintptr_t)out) & 15u) == 0) {
    const int64_t n4 = n >> 2;
    for (int64_t k = i; k < n4; k += stride) {
      int4 v = reinterpret_cast<const int4*>(src)[k];
      v.x += add; v.y += add; v.z += add; v.w += add;
      reinterpret_cast<int4*>(out)[k] = v;
    }
    for (int64_t k = (n4 << 2) + i; k < n; k += stride) out[k] = src[k] + add;
  } else {
    for (int64_t k = i; k < n; k += stride) out[k] = src[k] + add;
  }
}

struct CsrWorkspace {
  int32_t *cnt, *cursor, *chunk, *big_rows, *n_big, *tmp;
  size_t bytes;
};

static CsrWorkspace carve(void* base, int64_t n_rows, int64_t n_items) {
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  int64_t n_chunks = ceil_div(n_rows + 1, kScanChunk);
  size_t off = 0;
  CsrWorkspace w;
  char* p = static_cast<char*>(base);
  w.cnt = reinterpret_cast<int32_t*>(p + off);      off += up(sizeof(int32_t) * (n_rows + 1));
  w.cursor = reinterpret_cast<int32_t*>(p + off);   off += up(sizeof(int32_t) * (n_rows + 1));
  w.n_big = reinterpret_cast<int32_t*>(p + off);    off += up(sizeof(int32_t) * 64);
  // everything above is zero-initialised with ONE memset
  w.chunk = reinterpret_cast<int32_t*>(p + off);    off += up(sizeof(int32_t) * (n_chunks + 1));
  w.big_rows = reinterpret_cast<int32_t*>(p + off); off += up(sizeof(int32_t) * (n_items / kSmallRow + 1));
  w.tmp = reinterpret_cast<int32_t*>(p + off);      off += up(sizeof(int32_t) * (n_items + 1));
  w.bytes = off;
  return w;
}

}  // namespace stinet

using namespace stinet;

extern "C" size_t stinet_csr_workspace_bytes(int64_t n_rows, int64_t n_items) {
  if (n_rows < 0 || n_items < 0) return 0;
  return carve(nullptr, n_rows, n_items).bytes;
}

extern "C" int stinet_csr_build(const int64_t* key, const int64_t* other, int64_t n_items, int64_t n_rows,
                                int32_t* rowptr, int32_t* perm, int32_t* col, int32_t* key32, int32_t* status,
                                void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_items >= 0 && n_rows >= 0, STINET_ERR_ARG, "csr_build: negative size");
  STINET_REQUIRE(n_items < (int64_t(1) << 31) - 1 && n_rows < (int64_t(1) << 31) - 1, STINET_ERR_UNSUPPORTED,
                 "csr_build: sizes must fit int32");
  STINET_REQUIRE(rowptr && (n_items == 0 || (key && perm)), STINET_ERR_ARG, "csr_build: null pointer");
  STINET_REQUIRE(!(other && !col), STINET_ERR_ARG, "csr_build: `other` given without `col`");
  CsrWorkspace w = carve(workspace, n_rows, n_items);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "csr_build: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  size_t zero_bytes = reinterpret_cast<char*>(w.chunk) - reinterpret_cast<char*>(w.cnt);
  cudaMemsetAsync(w.cnt, 0, zero_bytes, stream);
  const int threads = 256;
  const int grid_items = wave_grid(n_items, threads * 4, 8);
  if (n_items > 0)
    K(csr_count_kernel<<<grid_items, threads, 0, stream>>>(key, n_items, n_rows, w.cnt, key32, status));
  const int n_chunks = (int)ceil_div(n_rows + 1, kScanChunk);
  K(scan_chunk_sums<<<n_chunks, kScanThreads, 0, stream>>>(w.cnt, n_rows, w.chunk));
  K(scan_write<<<n_chunks, kScanThreads, 0, stream>>>(w.cnt, n_rows, w.chunk, rowptr));
  if (n_items > 0) {
    K(csr_fill_kernel<<<grid_items, threads, 0, stream>>>(key, n_items, n_rows, rowptr, w.cursor, perm));
    K(csr_sort_small_rows<<<wave_grid(n_rows, threads, 8), threads, 0, stream>>>(rowptr, n_rows, perm, w.big_rows,
                                                                                 w.n_big));
    K(csr_rank_big_rows<<<kSMs * 2, 256, 0, stream>>>(rowptr, perm, w.big_rows, w.n_big, w.tmp));
    // (with `other` == NULL the kernel only has a tail to fill; meshes have none, but the launch decides that on the device)
    K(csr_gather_col<<<grid_items, threads, 0, stream>>>(other, perm, n_items, n_rows, rowptr, other ? col : nullptr, status));
  }
  return check_launch("csr_build");
}

extern "C" int stinet_concat_i32(const int32_t* const* src, const int64_t* len, const int64_t* dst_off,
                                 const int32_t* add, int n_parts, int32_t* dst, stinet_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_parts >= 0, STINET_ERR_ARG, "concat_i32: negative part count");
  if (n_parts == 0) return STINET_OK;
  STINET_REQUIRE(src && len && dst_off && add && dst, STINET_ERR_ARG, "concat_i32: null pointer");
  for (int p0 = 0; p0 < n_parts; p0 += kConcatParts) {
    ConcatArgs a;
    a.n_parts = n_parts - p0 < kConcatParts ? n_parts - p0 : kConcatParts;
    int64_t longest = 0;
    for (int p = 0; p < kConcatParts; ++p) {
      const bool on = p < a.n_parts;
      a.src[p] = on ? src[p0 + p] : nullptr;
      a.len[p] = on ? len[p0 + p] : 0;
      a.dst_off[p] = on ? dst_off[p0 + p] : 0;
      a.add[p] = on ? add[p0 + p] : 0;
      STINET_REQUIRE(!on || (a.len[p] >= 0 && a.dst_off[p] >= 0 && (a.len[p] == 0 || a.src[p])), STINET_ERR_ARG,
                     "concat_i32: bad part %d", p0 + p);
      if (a.len[p] > longest) longest = a.len[p];
    }
    if (longest == 0) continue;
    int gx = wave_grid(longest, 256 * 8, 4) / a.n_parts;     // ~4 CTAs per SM over all parts
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)a.n_parts);
    K(concat_i32_kernel<<<grid, 256, 0, stream>>>(a, dst));
  }
  return check_launch("concat_i32");
}

extern "C" int stinet_csr_cross_positions(const int32_t* eid_t, const int32_t* eid_s, int64_t n_items, int32_t* tpos_s,
                                          int32_t* scratch, stinet_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_items >= 0, STINET_ERR_ARG, "csr_cross_positions: negative size");
  if (n_items == 0) return STINET_OK;
  STINET_REQUIRE(eid_t && eid_s && tpos_s && scratch, STINET_ERR_ARG, "csr_cross_positions: null pointer");
  const int grid = wave_grid(n_items, 256 * 4, 8);
  // an edge that one of the two structures dropped (index out of range, status bit 0) has no by-target slot: it maps to
  // slot 0 instead of whatever the scratch buffer held
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(int32_t) * (size_t)n_items, stream);
  STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "csr_cross_positions: cudaMemsetAsync: %s", cudaGetErrorString(e));
  K(invert_perm_kernel<<<grid, 256, 0, stream>>>(eid_t, n_items, scratch));
  K(gather_i32_kernel<<<grid, 256, 0, stream>>>(scratch, eid_s, n_items, tpos_s));
  return check_launch("csr_cross_positions");
}
