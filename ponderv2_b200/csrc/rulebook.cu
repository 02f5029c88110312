// Rulebook (neighbour map) construction for submanifold and strided sparse convolutions.
//
// Replaces the indice-pair generation spconv runs the first time an `indice_key` is seen
// (reference call sites: ponder/models/sparse_unet/spconv_unet_v1m1_base.py:47-66,111-119
// for SubMConv3d, :135-142 for SparseConv3d k2 s2, :171-177 reuse by SparseInverseConv3d).
// Semantics follow SURVEY.md Appendix B; oracle = oracle/spconv_oracle.py.
//
// Pure integer work, bound by random 8-12 B probes into an L2-resident open-addressing table
// (2x load headroom, 12 B/slot: 3 MB for 100 k voxels, 25 MB for 1 M — inside the 126 MB L2)
// plus the coalesced 4*K*N B write of the map.  Algorithmic HBM bytes: 16 N + 4 K N.
#include "pv2_common.cuh"

namespace {

constexpr uint64_t kEmptyKey = 0xffffffffffffffffULL;
constexpr int kScanItems = 8;      // items per thread in the block scans
constexpr int kScanThreads = 256;
constexpr int kScanTile = kScanItems * kScanThreads;

struct HashTable {
  uint64_t* keys;
  int32_t* vals;
  uint32_t mask;
};

__host__ __device__ inline int64_t hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

struct WorkspaceLayout {
  size_t keys_off, vals_off, rank_off, bsum_off, total;
};

inline WorkspaceLayout layout_for(int64_t n) {
  WorkspaceLayout L;
  int64_t cap = hash_capacity(n);
  size_t off = 0;
  L.keys_off = off; off += (size_t)cap * 8;
  L.vals_off = off; off += (size_t)cap * 4;
  L.rank_off = off; off += ((size_t)n * 4 + 255) / 256 * 256;
  L.bsum_off = off; off += (((size_t)n + kScanTile - 1) / kScanTile + 1) * 4;
  L.total = (off + 255) / 256 * 256;
  return L;
}

__device__ __forceinline__ uint64_t linear_key(int b, int c0, int c1, int c2, int s0, int s1, int s2) {
  return (((uint64_t)b * (uint64_t)s0 + (uint64_t)c0) * (uint64_t)s1 + (uint64_t)c1) * (uint64_t)s2 + (uint64_t)c2;
}

__global__ void init_table_kernel(uint64_t* keys, int32_t* vals, int64_t cap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < cap; i += stride) { keys[i] = kEmptyKey; vals[i] = 0x7fffffff; }
}

// returns the slot the key lives in
__device__ __forceinline__ uint32_t table_insert(const HashTable& t, uint64_t key, int32_t row) {
  uint32_t slot = pv2_hash64(key) & t.mask;
  while (true) {
    unsigned long long prev = atomicCAS((unsigned long long*)&t.keys[slot], (unsigned long long)kEmptyKey,
                                        (unsigned long long)key);
    if (prev == kEmptyKey || prev == key) {
      atomicMin(&t.vals[slot], row);
      return slot;
    }
    slot = (slot + 1) & t.mask;
  }
}

__device__ __forceinline__ int32_t table_lookup(const HashTable& t, uint64_t key) {
  uint32_t slot = pv2_hash64(key) & t.mask;
  while (true) {
    uint64_t k = __ldg(&t.keys[slot]);
    if (k == key) return __ldg(&t.vals[slot]);
    if (k == kEmptyKey) return -1;
    slot = (slot + 1) & t.mask;
  }
}

__global__ void insert_subm_kernel(const int4* __restrict__ coords, int64_t n, int s0, int s1, int s2, HashTable t) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int4 c = __ldg(&coords[i]);
    table_insert(t, linear_key(c.x, c.y, c.z, c.w, s0, s1, s2), (int32_t)i);
  }
}

// One thread per (output row j, outer offsets k0,k1); the thread walks the innermost axis k2.
// Rows are the fastest-varying index so the K stores of a warp are 128 B coalesced lines.
template <int KS>
__global__ void probe_subm_kernel(const int4* __restrict__ coords, int64_t n, int s0, int s1, int s2, HashTable t,
                                  int32_t* __restrict__ nbr, unsigned long long* __restrict__ pair_count) {
  constexpr int R = KS / 2;
  const int64_t total = (int64_t)KS * KS * n;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int found = 0;
  for (; idx < total; idx += stride) {
    const int kk = (int)(idx / n);
    const int64_t j = idx - (int64_t)kk * n;
    const int k0 = kk / KS, k1 = kk - k0 * KS;
    const int4 c = __ldg(&coords[j]);
    const int a0 = c.y + k0 - R, a1 = c.z + k1 - R;
    const bool ok01 = (a0 >= 0) & (a0 < s0) & (a1 >= 0) & (a1 < s1);
    int32_t res[KS];
#pragma unroll
    for (int k2 = 0; k2 < KS; ++k2) {
      const int a2 = c.w + k2 - R;
      res[k2] = -1;
      if (ok01 && a2 >= 0 && a2 < s2) res[k2] = table_lookup(t, linear_key(c.x, a0, a1, a2, s0, s1, s2));
    }
#pragma unroll
    for (int k2 = 0; k2 < KS; ++k2) {
      nbr[(int64_t)(kk * KS + k2) * n + j] = res[k2];
      found += (res[k2] >= 0);
    }
  }
  if (pair_count != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) found += __shfl_xor_sync(0xffffffffu, found, o);
    if ((threadIdx.x & 31) == 0 && found) atomicAdd(pair_count, (unsigned long long)found);
  }
}

// ---------------------------------------------------------------- strided (k2 s2 p0) ----------

__global__ void insert_down_kernel(const int4* __restrict__ coords, int64_t n, int o0, int o1, int o2, HashTable t,
                                   int32_t* __restrict__ slot_of) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int4 c = __ldg(&coords[i]);
    int q0 = c.y >> 1, q1 = c.z >> 1, q2 = c.w >> 1;
    // a 2x2x2 window that does not fit inside the (unpadded) input is dropped, as in a dense k2 s2 p0 conv
    if (q0 >= o0 || q1 >= o1 || q2 >= o2) { slot_of[i] = -1; continue; }
    slot_of[i] = (int32_t)table_insert(t, linear_key(c.x, q0, q1, q2, o0, o1, o2), (int32_t)i);
  }
}

__device__ __forceinline__ int is_representative(const int32_t* slot_of, const int32_t* vals, int64_t i) {
  int32_t s = slot_of[i];
  return (s >= 0 && vals[s] == (int32_t)i) ? 1 : 0;
}

__global__ void __launch_bounds__(kScanThreads) flag_reduce_kernel(const int32_t* __restrict__ slot_of,
                                                                   const int32_t* __restrict__ vals, int64_t n,
                                                                   int32_t* __restrict__ bsum) {
  __shared__ int warp_sums[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int local = 0;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    int64_t i = base + (int64_t)it * kScanThreads + threadIdx.x;
    if (i < n) local += is_representative(slot_of, vals, i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) s += warp_sums[w];
    bsum[blockIdx.x] = s;
  }
}

// single block: exclusive scan of the per-tile sums, total -> n_out
__global__ void __launch_bounds__(1024) scan_bsum_kernel(int32_t* __restrict__ bsum, int nb, int32_t* __restrict__ n_out) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nb) ? bsum[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
      warp_tot[threadIdx.x] = w;  // inclusive
    }
    __syncthreads();
    int warp_off = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    int carry = carry_s;
    if (i < nb) bsum[i] = carry + warp_off + x - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = carry_s;
}

// rank[i] = number of representatives before row i (valid for representatives)
__global__ void __launch_bounds__(kScanThreads) rank_kernel(const int32_t* __restrict__ slot_of,
                                                            const int32_t* __restrict__ vals, int64_t n,
                                                            const int32_t* __restrict__ bsum, int32_t* __restrict__ rank) {
  __shared__ int warp_tot[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;  // blocked arrangement
  int f[kScanItems];
  int local = 0;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    int64_t i = base + it;
    f[it] = (i < n) ? is_representative(slot_of, vals, i) : 0;
    local += f[it];
  }
  int x = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  int warp_off = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_off += warp_tot[w];
  int run = bsum[blockIdx.x] + warp_off + x - local;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    int64_t i = base + it;
    if (i < n) rank[i] = run;
    run += f[it];
  }
}

__global__ void finalize_down_kernel(const int4* __restrict__ coords, int64_t n, const int32_t* __restrict__ vals,
                                     const int32_t* __restrict__ rank, int32_t* __restrict__ in2out /* holds slot_of on entry */,
                                     int32_t* __restrict__ koff, int4* __restrict__ out_coords) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int4 c = __ldg(&coords[i]);
    int32_t s = in2out[i];
    koff[i] = ((c.y & 1) * 2 + (c.z & 1)) * 2 + (c.w & 1);
    if (s < 0) { in2out[i] = -1; continue; }
    int32_t rep = vals[s];
    int32_t o = rank[rep];
    in2out[i] = o;
    if (rep == (int32_t)i) out_coords[o] = make_int4(c.x, c.y >> 1, c.z >> 1, c.w >> 1);
  }
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

__global__ void down_maps_kernel(const int32_t* __restrict__ in2out, const int32_t* __restrict__ koff, int64_t n,
                                 int64_t n_out, int32_t* __restrict__ nbr_down, int32_t* __restrict__ nbr_up) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int32_t o = in2out[i];
    int32_t k = koff[i];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) nbr_up[(int64_t)kk * n + i] = (kk == k) ? o : -1;
    // unsigned min: 0xffffffff (-1) means empty, duplicates resolve to the smallest fine row
    if (o >= 0) atomicMin((unsigned int*)&nbr_down[(int64_t)k * n_out + o], (unsigned int)i);
  }
}

__global__ void make_indices_kernel(const int64_t* __restrict__ grid_coord, const int64_t* __restrict__ offset,
                                    int64_t n, int batch, int4* __restrict__ indices) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    // batch id = number of cumulative offsets <= i (binary search; batch is tiny)
    int lo = 0, hi = batch;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(&offset[mid]) <= i) lo = mid + 1; else hi = mid; }
    indices[i] = make_int4(lo, (int)grid_coord[3 * i], (int)grid_coord[3 * i + 1], (int)grid_coord[3 * i + 2]);
  }
}

}  // namespace

extern "C" {

size_t pv2_rulebook_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return layout_for(n).total;
}

int pv2_rulebook_subm(const int32_t* coords, int64_t n, const int32_t* shape, int ksize, int32_t* nbr,
                      int64_t* pair_count, void* workspace, size_t workspace_bytes, void* stream_) {
  PV2_CHECK_ARG(n >= 0 && shape != nullptr && (ksize == 1 || ksize == 3 || ksize == 5));
  if (n == 0) return 0;
  PV2_CHECK_ARG(coords != nullptr && nbr != nullptr && workspace != nullptr);
  PV2_CHECK_ARG(((uintptr_t)coords & 15) == 0);
  WorkspaceLayout L = layout_for(n);
  if (workspace_bytes < L.total) return PV2_EWORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  int64_t cap = hash_capacity(n);
  HashTable t{(uint64_t*)((char*)workspace + L.keys_off), (int32_t*)((char*)workspace + L.vals_off), (uint32_t)(cap - 1)};
  init_table_kernel<<<pv2_grid_for(cap, 256), 256, 0, stream>>>(t.keys, t.vals, cap);
  insert_subm_kernel<<<pv2_grid_for(n, 256), 256, 0, stream>>>((const int4*)coords, n, shape[0], shape[1], shape[2], t);
  if (pair_count) cudaMemsetAsync(pair_count, 0, sizeof(int64_t), stream);
  int64_t work = (int64_t)ksize * ksize * n;
  int grid = pv2_grid_for(work, 256);
  if (ksize == 1)
    probe_subm_kernel<1><<<grid, 256, 0, stream>>>((const int4*)coords, n, shape[0], shape[1], shape[2], t, nbr, (unsigned long long*)pair_count);
  else if (ksize == 3)
    probe_subm_kernel<3><<<grid, 256, 0, stream>>>((const int4*)coords, n, shape[0], shape[1], shape[2], t, nbr, (unsigned long long*)pair_count);
  else
    probe_subm_kernel<5><<<grid, 256, 0, stream>>>((const int4*)coords, n, shape[0], shape[1], shape[2], t, nbr, (unsigned long long*)pair_count);
  PV2_DONE(3);
}

int pv2_rulebook_down(const int32_t* coords, int64_t n, const int32_t* shape, int32_t* out_coords, int32_t* in2out,
                      int32_t* koff, int32_t* n_out, void* workspace, size_t workspace_bytes, void* stream_) {
  PV2_CHECK_ARG(n >= 0 && shape != nullptr && n_out != nullptr);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) { cudaMemsetAsync(n_out, 0, 4, stream); return 0; }
  PV2_CHECK_ARG(coords && out_coords && in2out && koff && workspace);
  PV2_CHECK_ARG(((uintptr_t)coords & 15) == 0 && ((uintptr_t)out_coords & 15) == 0);
  WorkspaceLayout L = layout_for(n);
  if (workspace_bytes < L.total) return PV2_EWORKSPACE;
  int64_t cap = hash_capacity(n);
  HashTable t{(uint64_t*)((char*)workspace + L.keys_off), (int32_t*)((char*)workspace + L.vals_off), (uint32_t)(cap - 1)};
  int32_t* rank = (int32_t*)((char*)workspace + L.rank_off);
  int32_t* bsum = (int32_t*)((char*)workspace + L.bsum_off);
  int o0 = (shape[0] - 2) / 2 + 1, o1 = (shape[1] - 2) / 2 + 1, o2 = (shape[2] - 2) / 2 + 1;
  PV2_CHECK_ARG(o0 > 0 && o1 > 0 && o2 > 0);
  init_table_kernel<<<pv2_grid_for(cap, 256), 256, 0, stream>>>(t.keys, t.vals, cap);
  insert_down_kernel<<<pv2_grid_for(n, 256), 256, 0, stream>>>((const int4*)coords, n, o0, o1, o2, t, in2out);
  int nb = (int)((n + kScanTile - 1) / kScanTile);
  flag_reduce_kernel<<<nb, kScanThreads, 0, stream>>>(in2out, t.vals, n, bsum);
  scan_bsum_kernel<<<1, 1024, 0, stream>>>(bsum, nb, n_out);
  rank_kernel<<<nb, kScanThreads, 0, stream>>>(in2out, t.vals, n, bsum, rank);
  finalize_down_kernel<<<pv2_grid_for(n, 256), 256, 0, stream>>>((const int4*)coords, n, t.vals, rank, in2out, koff, (int4*)out_coords);
  PV2_DONE(6);
}

int pv2_rulebook_down_maps(const int32_t* in2out, const int32_t* koff, int64_t n, int64_t n_out, int32_t* nbr_down,
                           int32_t* nbr_up, void* stream_) {
  PV2_CHECK_ARG(n >= 0 && n_out >= 0);
  if (n == 0) return 0;
  PV2_CHECK_ARG(in2out && koff && nbr_up && (n_out == 0 || nbr_down));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_out > 0) fill_i32_kernel<<<pv2_grid_for(8 * n_out, 256), 256, 0, stream>>>(nbr_down, 8 * n_out, -1);
  down_maps_kernel<<<pv2_grid_for(n, 256), 256, 0, stream>>>(in2out, koff, n, n_out, nbr_down, nbr_up);
  PV2_DONE(n_out > 0 ? 2 : 1);
}

int pv2_make_indices(const int64_t* grid_coord, const int64_t* offset, int64_t n, int batch, int32_t* indices, void* stream_) {
  PV2_CHECK_ARG(n >= 0 && batch >= 1);
  if (n == 0) return 0;
  PV2_CHECK_ARG(grid_coord && offset && indices && ((uintptr_t)indices & 15) == 0);
  make_indices_kernel<<<pv2_grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(grid_coord, offset, n, batch, (int4*)indices);
  PV2_DONE(1);
}

}  // extern "C"
