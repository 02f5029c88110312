// Tile order for the output-stationary sparse convolution: rows sorted by their neighbour-presence mask.
//
// A 128-row tile of the gather-GEMM (spconv_umma.cu) has to run kernel offset k as soon as ONE of its rows has a
// neighbour at k.  In scan order almost every tile touches almost every offset (24.6 of 27 on the 100 k-voxel indoor
// scene) although a surface voxel has ~9 neighbours; grouping rows with equal masks makes the tiles homogeneous (9.6 of
// 27), which cuts gathered bytes, weight traffic and MMA work by the same factor.  spconv does the same for its
// implicit-GEMM kernels (mask sort of the indice pairs); results do not depend on the order.
//   mask[j] = OR_k (nbr[k][j] >= 0) << (k mod 32)  (exact for kvol <= 32; a folded signature up to kvol = 128)
//   order   = stable radix sort of rows by mask   (cub::DeviceRadixSort over kvol bits; deterministic)
#include "pv2_common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

__global__ void row_mask_kernel(const int32_t* __restrict__ nbr, int64_t n, int kvol, uint32_t* __restrict__ mask,
                                int32_t* __restrict__ iota) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; j < n; j += stride) {
    uint32_t m = 0;
    // K > 32 (the 5x5x5 stem): offsets are folded modulo 32 — a grouping signature, no longer the exact mask
    for (int k = 0; k < kvol; ++k) m |= (uint32_t)(__ldg(&nbr[(int64_t)k * n + j]) >= 0) << (k & 31);
    mask[j] = m;
    iota[j] = (int32_t)j;
  }
}

// nbr_sorted[k][pos] = nbr[k][order[pos]] (the map in tile order: coalesced for the conv kernels) and
// blk_active[k][b] = 1 iff one of the 32 rows of block b has a neighbour at offset k (wgrad skips the other stages)
__global__ void permute_map_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ order, int64_t n,
                                   int kvol, int64_t nblk, int32_t* __restrict__ nbr_sorted,
                                   uint8_t* __restrict__ blk_active) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp0; b < nblk; b += nwarps) {
    const int64_t pos = b * 32 + lane;
    const int64_t j = pos < n ? (int64_t)__ldg(&order[pos]) : -1;
    for (int k = 0; k < kvol; ++k) {
      const int32_t v = j >= 0 ? __ldg(&nbr[(int64_t)k * n + j]) : -1;
      if (pos < n && nbr_sorted != nullptr) nbr_sorted[(int64_t)k * n + pos] = v;
      const bool any = __any_sync(0xffffffffu, v >= 0);
      if (lane == 0 && blk_active != nullptr) blk_active[(int64_t)k * nblk + b] = any ? 1 : 0;
    }
  }
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" {

size_t pv2_rulebook_row_order_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  // mask in/out + iota + cub temporaries.  cub reports its need exactly when a device is present; without one (symbol
  // checks on a CPU-only box) a generous bound is returned.
  size_t temp = 0;
  uint32_t* k = nullptr; int32_t* v = nullptr;
  if (cub::DeviceRadixSort::SortPairs(nullptr, temp, k, k, v, v, (int)n, 0, 32, (cudaStream_t)0) != cudaSuccess) {
    (void)cudaGetLastError();
    temp = (size_t)n * 16 + (16u << 20);
  }
  return 3 * align256((size_t)n * 4) + align256(temp + 256);
}

int pv2_rulebook_row_order(const int32_t* nbr, int64_t n, int kvol, int32_t* order, int32_t* nbr_sorted,
                           uint8_t* blk_active, void* workspace, size_t workspace_bytes, void* stream_) {
  PV2_CHECK_ARG(n >= 0 && kvol >= 1);
  if (kvol > 128) return PV2_EUNSUPPORTED;
  const int key_bits = kvol < 32 ? kvol : 32;
  if (n == 0) return 0;
  PV2_CHECK_ARG(nbr && order && workspace && n < (int64_t)1 << 31);
  if (workspace_bytes < pv2_rulebook_row_order_workspace_bytes(n)) return PV2_EWORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t a = align256((size_t)n * 4);
  char* ws = (char*)workspace;
  uint32_t* mask_in = (uint32_t*)ws;
  uint32_t* mask_out = (uint32_t*)(ws + a);
  int32_t* iota = (int32_t*)(ws + 2 * a);
  void* temp = ws + 3 * a;
  size_t temp_have = workspace_bytes - 3 * a, temp_need = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp_need, mask_in, mask_out, iota, order, (int)n, 0, key_bits, stream);
  if (e != cudaSuccess) return (int)e;
  if (temp_need > temp_have) return PV2_EWORKSPACE;
  row_mask_kernel<<<pv2_grid_for(n, 256), 256, 0, stream>>>(nbr, n, kvol, mask_in, iota);
  e = cub::DeviceRadixSort::SortPairs(temp, temp_need, mask_in, mask_out, iota, order, (int)n, 0, key_bits, stream);
  if (e != cudaSuccess) return (int)e;
  int launches = 2;
  if (nbr_sorted != nullptr || blk_active != nullptr) {
    const int64_t nblk = (n + 31) / 32;
    permute_map_kernel<<<pv2_grid_for(nblk * 32, 256), 256, 0, stream>>>(nbr, order, n, kvol, nblk, nbr_sorted, blk_active);
    ++launches;
  }
  PV2_DONE(launches);
}

}  // extern "C"
