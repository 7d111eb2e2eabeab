// Stable LSD radix sort of (uint64 key, int32 value) pairs, 8 bits per pass -- the integer workhorse of hierarchy
// construction (voxel keys -> clusters, coarse edge pairs -> edge set: preprocessing/graph_level_generation.py:194-244 in the
// reference, np.unique there) and of the degree-sorted row order of a CSR.  HBM-bound byte/integer work: no tensor cores.
//
// One pass = three kernels over tiles of 2048 items (256 threads, a warp owns 256 consecutive items):
//   hist    : per-tile digit histogram (shared-memory atomics: counts do not depend on arrival order)
//   scan    : per digit, exclusive prefix of the tile counts (one CTA per digit) + the digit totals
//   scatter : stable rank of every item = (items with the same digit in earlier tiles) + (in earlier warps of the tile)
//             + (earlier in the warp's own 256 items, walked 32 at a time with __match_any_sync), so equal keys keep their
//             input order by construction -- no atomics on the output side, results are run-to-run identical.
// Keys with `key_bits` significant bits take ceil(key_bits / 8) passes, ping-ponging between the output and a scratch
// buffer so that the last pass lands in the output.
#include "common.cuh"

namespace stinet {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;   // 2048
constexpr int kSortWarps = kSortThreads / 32;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist) {
  __shared__ int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[(int)((keys[i] >> shift) & 0xFFu)], 1);
  }
  __syncthreads();
  hist[(int64_t)blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

// CTA d: prefix[t][d] = sum_{t' < t} hist[t'][d]  (in place), total[d] = sum_t hist[t][d]
__global__ void __launch_bounds__(1024) radix_scan_kernel(int32_t* __restrict__ hist, int n_tiles, int32_t* __restrict__ total) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int d = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_tiles; t0 += 1024) {
    const int t = t0 + threadIdx.x;
    const int v = t < n_tiles ? hist[(int64_t)t * 256 + d] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (wid ? wsum[wid - 1] : 0) + inc - v;
    if (t < n_tiles) hist[(int64_t)t * 256 + d] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[d] = carry_s;
}

template <bool HAS_VAL_IN>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals, int64_t n, int shift,
                     const int32_t* __restrict__ prefix, const int32_t* __restrict__ total, uint64_t* __restrict__ keys_out,
                     int32_t* __restrict__ vals_out) {
  __shared__ int wcnt[kSortWarps][256];      // per warp: items of each digit (then: items in earlier warps)
  __shared__ int gbase[256];                 // first output slot of this tile's items of each digit
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcnt[0][0])[i] = 0;
  {
    // exclusive scan of the 256 digit totals (every CTA does its own: 256 values)
    __shared__ int ws[8];
    const int v = total[threadIdx.x];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += ws[k];
    gbase[threadIdx.x] = off + inc - v + prefix[(int64_t)blockIdx.x * 256 + threadIdx.x];
  }
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile + (int64_t)w * (32 * kSortItems);
  uint64_t key[kSortItems];
  int32_t val[kSortItems];
  int local[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = base + r * 32 + lane;
    const bool ok = i < n;
    key[r] = ok ? keys[i] : 0;
    val[r] = ok ? (HAS_VAL_IN ? vals[i] : (int32_t)i) : 0;
    const int d = ok ? (int)((key[r] >> shift) & 0xFFu) : 256 + lane;          // lanes past the end match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    int prev = 0;
    if (ok && lane == leader) {
      prev = wcnt[w][d];
      wcnt[w][d] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    local[r] = prev + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;               // items of digit d in earlier warps of this tile
    int run = 0;
#pragma unroll
    for (int k = 0; k < kSortWarps; ++k) {
      const int c = wcnt[k][d];
      wcnt[k][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = base + r * 32 + lane;
    if (i < n) {
      const int d = (int)((key[r] >> shift) & 0xFFu);
      const int64_t dst = (int64_t)gbase[d] + wcnt[w][d] + local[r];
      keys_out[dst] = key[r];
      if (vals_out) vals_out[dst] = val[r];
    }
  }
}

struct SortWs { uint64_t* keys_tmp; int32_t* vals_tmp; int32_t* hist; int32_t* total; size_t bytes; };
static SortWs carve_sort(void* base, int64_t n) {
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const int64_t tiles = ceil_div(n > 0 ? n : 1, kSortTile);
  SortWs w;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  w.keys_tmp = reinterpret_cast<uint64_t*>(p + off); off += up(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
  w.vals_tmp = reinterpret_cast<int32_t*>(p + off);  off += up(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  w.hist = reinterpret_cast<int32_t*>(p + off);      off += up(sizeof(int32_t) * (size_t)tiles * 256);
  w.total = reinterpret_cast<int32_t*>(p + off);     off += up(sizeof(int32_t) * 256);
  w.bytes = off;
  return w;
}

// rows of a CSR as sort keys: larger degree first, ties in ascending row order (the sort is stable)
__global__ void degree_key_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int max_deg, uint64_t* __restrict__ key) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int d = min(rowptr[r + 1] - rowptr[r], max_deg);
    key[r] = (uint64_t)(max_deg - d);
  }
}

}  // namespace stinet

using namespace stinet;

extern "C" size_t stinet_sort_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return carve_sort(nullptr, n).bytes;
}

extern "C" int stinet_sort_pairs_u64(const uint64_t* keys_in, const int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out,
                                     int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                                     stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n >= 0 && n < (int64_t(1) << 31) - kSortTile && key_bits >= 1 && key_bits <= 64, STINET_ERR_ARG,
                 "sort_pairs_u64: bad size / key_bits");
  if (n == 0) return STINET_OK;
  STINET_REQUIRE(keys_in && keys_out, STINET_ERR_ARG, "sort_pairs_u64: null pointer");
  SortWs w = carve_sort(workspace, n);
  STINET_REQUIRE(workspace && workspace_bytes >= w.bytes, STINET_ERR_WORKSPACE, "sort_pairs_u64: workspace %zu < %zu",
                 workspace_bytes, w.bytes);
  const int passes = (key_bits + 7) / 8;
  const int tiles = (int)ceil_div(n, kSortTile);
  const uint64_t* src_k = keys_in;
  const int32_t* src_v = vals_in;
  for (int p = 0; p < passes; ++p) {
    // the last pass must land in the caller's output: with an odd number of passes the first one goes there too
    const bool to_out = ((passes - 1 - p) & 1) == 0;
    uint64_t* dst_k = to_out ? keys_out : w.keys_tmp;
    int32_t* dst_v = vals_out ? (to_out ? vals_out : w.vals_tmp) : nullptr;
    K(radix_hist_kernel<<<tiles, kSortThreads, 0, s>>>(src_k, n, 8 * p, w.hist));
    K(radix_scan_kernel<<<256, 1024, 0, s>>>(w.hist, tiles, w.total));
    if (src_v != nullptr)
      K(radix_scatter_kernel<true><<<tiles, kSortThreads, 0, s>>>(src_k, src_v, n, 8 * p, w.hist, w.total, dst_k, dst_v));
    else
      K(radix_scatter_kernel<false><<<tiles, kSortThreads, 0, s>>>(src_k, nullptr, n, 8 * p, w.hist, w.total, dst_k, dst_v));
    src_k = dst_k;
    src_v = dst_v;             // after the first pass the payload travels with the keys (NULL: keys only)
  }
  return check_launch("sort_pairs_u64");
}

extern "C" int stinet_csr_degree_order(const int32_t* rowptr, int64_t n_rows, int32_t* order, void* workspace,
                                       size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_rows >= 0, STINET_ERR_ARG, "csr_degree_order: negative size");
  if (n_rows == 0) return STINET_OK;
  STINET_REQUIRE(rowptr && order, STINET_ERR_ARG, "csr_degree_order: null pointer");
  // keys in the first part of the workspace (in and out), the sort's own scratch behind them
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t kb = up(sizeof(uint64_t) * (size_t)n_rows);
  const size_t need = 2 * kb + carve_sort(nullptr, n_rows).bytes;
  STINET_REQUIRE(workspace && workspace_bytes >= need, STINET_ERR_WORKSPACE, "csr_degree_order: workspace %zu < %zu",
                 workspace_bytes, need);
  char* p = static_cast<char*>(workspace);
  uint64_t* k_in = reinterpret_cast<uint64_t*>(p);
  uint64_t* k_out = reinterpret_cast<uint64_t*>(p + kb);
  constexpr int kMaxDeg = 65535;     // rows of higher degree share the first bucket
  K(degree_key_kernel<<<wave_grid(n_rows, 256, 8), 256, 0, s>>>(rowptr, n_rows, kMaxDeg, k_in));
  return stinet_sort_pairs_u64(k_in, nullptr, k_out, order, n_rows, 16, p + 2 * kb, workspace_bytes - 2 * kb, stream_);
}

extern "C" size_t stinet_csr_degree_order_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return 0;
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  return 2 * up(sizeof(uint64_t) * (size_t)(n_rows > 0 ? n_rows : 1)) + carve_sort(nullptr, n_rows).bytes;
}
