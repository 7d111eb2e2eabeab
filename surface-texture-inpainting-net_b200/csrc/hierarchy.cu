// Hierarchy construction by vertex clustering: the integer / byte kernels around the radix sort of
// sort.cu that replace reference preprocessing/graph_level_generation.py:194-244 (`vertex_clustering`: Python loops over
// bins, points and neighbour sets) -- voxel binning with numpy's floor-division semantics, packed 64-bit cell keys,
// sort-unique + inverse (= the trace map), the coarse edge set, and the per-cluster centres of gravity summed in ascending
// member order in the input dtype (what `coords[members].mean(axis=0)` does), so every output is bit-identical to the
// reference's.  HBM-bound; one thread per item, 64-bit keys, no floating-point atomics.
#include "common.cuh"

namespace stinet {

// numpy's floor_divide for floating point (npy_divmod): fmod-based, then a correction toward floor
template <typename T>
__device__ __forceinline__ T np_floor_divide(T a, T b) {
  T mod = fmod(a, b);
  T div = (a - mod) / b;
  if (mod != T(0)) {
    if ((b < T(0)) != (mod < T(0))) div -= T(1);
  }
  if (div != T(0)) {
    T fl = floor(div);
    if (div - fl > T(0.5)) fl += T(1);
    return fl;
  }
  return copysign(T(0), a / b);
}

__global__ void minmax_init_kernel(long long* __restrict__ mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = 0x7FFFFFFFFFFFFFFFll;
  else if (threadIdx.x < 6) mm[threadIdx.x] = (long long)0x8000000000000000ull;
}

template <typename T>
__global__ void __launch_bounds__(256)
voxel_bins_kernel(const T* __restrict__ coords, int64_t n, T voxel, long long* __restrict__ bins, long long* __restrict__ mm) {
  long long lo[3] = {0x7FFFFFFFFFFFFFFFll, 0x7FFFFFFFFFFFFFFFll, 0x7FFFFFFFFFFFFFFFll};
  long long hi[3] = {(long long)0x8000000000000000ull, (long long)0x8000000000000000ull, (long long)0x8000000000000000ull};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const long long b = (long long)np_floor_divide<T>(coords[3 * i + a], voxel);
      bins[3 * i + a] = b;
      lo[a] = min(lo[a], b);
      hi[a] = max(hi[a], b);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&mm[a], lo[a]);
      atomicMax(&mm[3 + a], hi[a]);
    }
  }
}

// key = ((bx - lox) * span_y + (by - loy)) * span_z + (bz - loz): lexicographic (x, y, z) order = np.unique(axis=0)
__global__ void voxel_keys_kernel(const long long* __restrict__ bins, const long long* __restrict__ mm, int64_t n,
                                  uint64_t* __restrict__ keys) {
  const long long lox = mm[0], loy = mm[1], loz = mm[2];
  const uint64_t sy = (uint64_t)(mm[4] - loy + 1), sz = (uint64_t)(mm[5] - loz + 1);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ((uint64_t)(bins[3 * i] - lox) * sy + (uint64_t)(bins[3 * i + 1] - loy)) * sz + (uint64_t)(bins[3 * i + 2] - loz);
}

// ---- unique of a sorted key array: flags -> exclusive scan (two kernels, fixed order) -> ids
constexpr int kUThreads = 1024;
constexpr int kUItems = 4;
constexpr int kUChunk = kUThreads * kUItems;

__device__ __forceinline__ int u_block_incl_scan(int v, int* total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const int off = wid ? wsum[wid - 1] : 0;
  *total = wsum[31];
  __syncthreads();
  return inc + off;
}

// flag(i) = key[i] < limit && (i == 0 || key[i] != key[i-1])
__device__ __forceinline__ int uflag(const uint64_t* __restrict__ keys, int64_t i, int64_t n, uint64_t limit) {
  if (i >= n) return 0;
  const uint64_t k = keys[i];
  return (k < limit && (i == 0 || k != keys[i - 1])) ? 1 : 0;
}

__global__ void __launch_bounds__(kUThreads) unique_chunk_sums_kernel(const uint64_t* __restrict__ keys, int64_t n, uint64_t limit,
                                                                     int32_t* __restrict__ chunk_sum) {
  const int64_t base = (int64_t)blockIdx.x * kUChunk + (int64_t)threadIdx.x * kUItems;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kUItems; ++k) s += uflag(keys, base + k, n, limit);
  int total;
  u_block_incl_scan(s, &total);
  if (threadIdx.x == 0) chunk_sum[blockIdx.x] = total;
}

// ids[i] = (number of flags up to and including i) - 1 = index of item i's unique value; count = total number of uniques
__global__ void __launch_bounds__(kUThreads) unique_ids_kernel(const uint64_t* __restrict__ keys, int64_t n, uint64_t limit,
                                                              const int32_t* __restrict__ chunk_sum, int n_chunks,
                                                              int32_t* __restrict__ ids, int32_t* __restrict__ count) {
  __shared__ int off_s;
  {
    int part = 0;
    for (int c = threadIdx.x; c < (int)blockIdx.x; c += kUThreads) part += chunk_sum[c];
    int total;
    u_block_incl_scan(part, &total);
    if (threadIdx.x == 0) off_s = total;
    __syncthreads();
  }
  const int64_t base = (int64_t)blockIdx.x * kUChunk + (int64_t)threadIdx.x * kUItems;
  int f[kUItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kUItems; ++k) {
    f[k] = uflag(keys, base + k, n, limit);
    s += f[k];
  }
  int total;
  const int inc = u_block_incl_scan(s, &total);
  int run = off_s + inc - s;
#pragma unroll
  for (int k = 0; k < kUItems; ++k) {
    run += f[k];
    if (base + k < n) ids[base + k] = run - 1;
  }
  if (blockIdx.x == (unsigned)(n_chunks - 1) && threadIdx.x == kUThreads - 1) *count = off_s + total;
}

// trace[idx_sorted[i]] = ids[i];  start[c] = first sorted position of cluster c;  start[n_coarse] = n
__global__ void cluster_finish_kernel(const int32_t* __restrict__ idx_sorted, const int32_t* __restrict__ ids, int64_t n,
                                      int64_t n_coarse, int64_t* __restrict__ trace, int32_t* __restrict__ start) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = ids[i];
    trace[idx_sorted[i]] = c;
    if (i == 0 || ids[i - 1] != c) start[c] = (int32_t)i;
    if (i == n - 1) start[n_coarse] = (int32_t)n;
  }
}

// centre of gravity of every cluster: members in ascending vertex id (the sort is stable and its payload is the vertex id),
// summed sequentially in the input dtype, divided by the count, stored as float32 -- numpy's coords[members].mean(axis=0)
template <typename T>
__global__ void cluster_centroids_kernel(const T* __restrict__ coords, const int32_t* __restrict__ idx_sorted,
                                         const int32_t* __restrict__ start, int64_t n_coarse, float* __restrict__ out) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_coarse; c += (int64_t)gridDim.x * blockDim.x) {
    const int b = start[c], e = start[c + 1];
    T sx = T(0), sy = T(0), sz = T(0);
    for (int k = b; k < e; ++k) {
      const int64_t v = idx_sorted[k];
      sx += coords[3 * v];
      sy += coords[3 * v + 1];
      sz += coords[3 * v + 2];
    }
    const T cnt = (T)(e - b);
    out[3 * c] = (float)(sx / cnt);
    out[3 * c + 1] = (float)(sy / cnt);
    out[3 * c + 2] = (float)(sz / cnt);
  }
}

// coarse edge candidates: (trace[v], trace[w]) of every fine edge as key a * n_coarse + b; self loops get the key `limit`
__global__ void coarse_edge_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n_edges,
                                        const int64_t* __restrict__ trace, int64_t n_fine, uint64_t n_coarse,
                                        uint64_t* __restrict__ keys, int32_t* __restrict__ status) {
  const uint64_t limit = n_coarse * n_coarse;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = src[e], w = dst[e];
    if (v < 0 || v >= n_fine || w < 0 || w >= n_fine) {
      keys[e] = limit;
      if (status) atomicOr(status, 1);
      continue;
    }
    const uint64_t a = (uint64_t)trace[v], b = (uint64_t)trace[w];
    keys[e] = a != b ? a * n_coarse + b : limit;
  }
}

// the unique coarse edges, sorted by (vertex, neighbour): out[0][j] = key / n_coarse, out[1][j] = key % n_coarse
__global__ void coarse_edges_emit_kernel(const uint64_t* __restrict__ keys_sorted, const int32_t* __restrict__ ids, int64_t n_edges,
                                         uint64_t n_coarse, int64_t n_out, int64_t* __restrict__ out) {
  const uint64_t limit = n_coarse * n_coarse;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_edges; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys_sorted[i];
    if (k < limit && (i == 0 || k != keys_sorted[i - 1])) {
      const int64_t j = ids[i];
      out[j] = (int64_t)(k / n_coarse);
      out[n_out + j] = (int64_t)(k % n_coarse);
    }
  }
}

}  // namespace stinet

using namespace stinet;

extern "C" int stinet_voxel_bins(const void* coords, int is_f64, int64_t n, double voxel, int64_t* bins, int64_t* minmax,
                                 stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n >= 0 && voxel > 0.0, STINET_ERR_ARG, "voxel_bins: bad size / voxel");
  STINET_REQUIRE(minmax && (n == 0 || (coords && bins)), STINET_ERR_ARG, "voxel_bins: null pointer");
  K(minmax_init_kernel<<<1, 32, 0, s>>>(reinterpret_cast<long long*>(minmax)));
  if (n > 0) {
    const int grid = wave_grid(n, 256, 8);
    if (is_f64)
      K(voxel_bins_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(coords), n, voxel, reinterpret_cast<long long*>(bins),
                                                      reinterpret_cast<long long*>(minmax)));
    else
      K(voxel_bins_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(coords), n, (float)voxel, reinterpret_cast<long long*>(bins),
                                                     reinterpret_cast<long long*>(minmax)));
  }
  return check_launch("voxel_bins");
}

extern "C" int stinet_voxel_keys(const int64_t* bins, const int64_t* minmax, int64_t n, uint64_t* keys, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n >= 0, STINET_ERR_ARG, "voxel_keys: negative size");
  if (n == 0) return STINET_OK;
  STINET_REQUIRE(bins && minmax && keys, STINET_ERR_ARG, "voxel_keys: null pointer");
  K(voxel_keys_kernel<<<wave_grid(n, 256, 8), 256, 0, s>>>(reinterpret_cast<const long long*>(bins),
                                                          reinterpret_cast<const long long*>(minmax), n, keys));
  return check_launch("voxel_keys");
}

extern "C" size_t stinet_unique_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return sizeof(int32_t) * (size_t)(ceil_div(n > 0 ? n : 1, kUChunk) + 1);
}

extern "C" int stinet_unique_sorted_u64(const uint64_t* keys_sorted, int64_t n, uint64_t limit, int32_t* ids, int32_t* count,
                                        void* workspace, size_t workspace_bytes, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n >= 0 && count, STINET_ERR_ARG, "unique_sorted_u64: bad arguments");
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), s);
    STINET_REQUIRE(e == cudaSuccess, STINET_ERR_CUDA, "unique_sorted_u64: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return STINET_OK;
  }
  STINET_REQUIRE(keys_sorted && ids, STINET_ERR_ARG, "unique_sorted_u64: null pointer");
  STINET_REQUIRE(workspace && workspace_bytes >= stinet_unique_workspace_bytes(n), STINET_ERR_WORKSPACE,
                 "unique_sorted_u64: workspace too small");
  const int n_chunks = (int)ceil_div(n, kUChunk);
  int32_t* chunk = static_cast<int32_t*>(workspace);
  K(unique_chunk_sums_kernel<<<n_chunks, kUThreads, 0, s>>>(keys_sorted, n, limit, chunk));
  K(unique_ids_kernel<<<n_chunks, kUThreads, 0, s>>>(keys_sorted, n, limit, chunk, n_chunks, ids, count));
  return check_launch("unique_sorted_u64");
}

extern "C" int stinet_cluster_finish(const int32_t* idx_sorted, const int32_t* ids, int64_t n, int64_t n_coarse, int64_t* trace,
                                     int32_t* start, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n >= 0 && n_coarse >= 0, STINET_ERR_ARG, "cluster_finish: negative size");
  if (n == 0) return STINET_OK;
  STINET_REQUIRE(idx_sorted && ids && trace && start, STINET_ERR_ARG, "cluster_finish: null pointer");
  K(cluster_finish_kernel<<<wave_grid(n, 256, 8), 256, 0, s>>>(idx_sorted, ids, n, n_coarse, trace, start));
  return check_launch("cluster_finish");
}

extern "C" int stinet_cluster_centroids(const void* coords, int is_f64, const int32_t* idx_sorted, const int32_t* start,
                                        int64_t n_coarse, float* out, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_coarse >= 0, STINET_ERR_ARG, "cluster_centroids: negative size");
  if (n_coarse == 0) return STINET_OK;
  STINET_REQUIRE(coords && idx_sorted && start && out, STINET_ERR_ARG, "cluster_centroids: null pointer");
  const int grid = wave_grid(n_coarse, 128, 8);
  if (is_f64) K(cluster_centroids_kernel<double><<<grid, 128, 0, s>>>(static_cast<const double*>(coords), idx_sorted, start, n_coarse, out));
  else K(cluster_centroids_kernel<float><<<grid, 128, 0, s>>>(static_cast<const float*>(coords), idx_sorted, start, n_coarse, out));
  return check_launch("cluster_centroids");
}

extern "C" int stinet_coarse_edge_keys(const int64_t* src, const int64_t* dst, int64_t n_edges, const int64_t* trace,
                                       int64_t n_fine, int64_t n_coarse, uint64_t* keys, int32_t* status, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_edges >= 0 && n_fine >= 0 && n_coarse >= 0 && n_coarse < (int64_t(1) << 31), STINET_ERR_ARG,
                 "coarse_edge_keys: bad size");
  if (n_edges == 0) return STINET_OK;
  STINET_REQUIRE(src && dst && trace && keys, STINET_ERR_ARG, "coarse_edge_keys: null pointer");
  K(coarse_edge_keys_kernel<<<wave_grid(n_edges, 256, 8), 256, 0, s>>>(src, dst, n_edges, trace, n_fine, (uint64_t)n_coarse, keys, status));
  return check_launch("coarse_edge_keys");
}

extern "C" int stinet_coarse_edges_emit(const uint64_t* keys_sorted, const int32_t* ids, int64_t n_edges, int64_t n_coarse,
                                        int64_t n_out, int64_t* out, stinet_stream_t stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  STINET_REQUIRE(n_edges >= 0 && n_coarse >= 0 && n_out >= 0, STINET_ERR_ARG, "coarse_edges_emit: negative size");
  if (n_edges == 0 || n_out == 0) return STINET_OK;
  STINET_REQUIRE(keys_sorted && ids && out, STINET_ERR_ARG, "coarse_edges_emit: null pointer");
  K(coarse_edges_emit_kernel<<<wave_grid(n_edges, 256, 8), 256, 0, s>>>(keys_sorted, ids, n_edges, (uint64_t)n_coarse, n_out, out));
  return check_launch("coarse_edges_emit");
}
