// Densify: scatter-mean of voxel features into the dense channels-last volume the renderer samples.
// Reference: ponder_indoor_base.py:177-216,332-342 (pooling branch) and ponder_outdoor_base.py:178-210
// (torch_scatter.scatter(reduce="mean", out=zeros) followed by view/permute/contiguous).
//
// The reference zero-fills a (cells, C) buffer, scatters, then permute-copies it to (C,Z,Y,X).  Here the
// caller passes the cell id already in OUTPUT memory order, so the volume is written once:
// algorithmic bytes = N*C*4 (read) + cells*C*4 (zero-fill/write).  HBM-bound.
#include "pv2_common.cuh"

namespace {

__global__ void zero_f4_kernel(float4* __restrict__ p, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (; i < n4; i += stride) p[i] = z;
}

__global__ void count_kernel(const int64_t* __restrict__ cell, int64_t n, int64_t cells, int32_t* __restrict__ count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t c = cell[i];
    if (c >= 0 && c < cells) atomicAdd(&count[c], 1);
  }
}

// one thread per (row, 4-channel group); the mean weight 1/count is applied before the atomic so no
// second pass over the volume is needed.  red.global.add.v4.f32 (sm_90+) moves 16 B per atomic.
__global__ void scatter_mean_v4_kernel(const float* __restrict__ feat, const int64_t* __restrict__ cell,
                                       const int32_t* __restrict__ count, int64_t n, int c4, int64_t cells,
                                       float* __restrict__ volume) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = n * c4;
  for (; idx < total; idx += stride) {
    int64_t i = idx / c4;
    int g = (int)(idx - i * c4);
    int64_t c = __ldg(&cell[i]);
    if (c < 0 || c >= cells) continue;
    float inv = 1.f / (float)__ldg(&count[c]);
    float4 v = __ldg(reinterpret_cast<const float4*>(feat) + idx);
    float* dst = volume + (c * c4 + g) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x * inv), "f"(v.y * inv),
                 "f"(v.z * inv), "f"(v.w * inv)
                 : "memory");
  }
}

__global__ void scatter_mean_scalar_kernel(const float* __restrict__ feat, const int64_t* __restrict__ cell,
                                           const int32_t* __restrict__ count, int64_t n, int ch, int64_t cells,
                                           float* __restrict__ volume) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = n * ch;
  for (; idx < total; idx += stride) {
    int64_t i = idx / ch;
    int g = (int)(idx - i * ch);
    int64_t c = __ldg(&cell[i]);
    if (c < 0 || c >= cells) continue;
    atomicAdd(&volume[c * ch + g], feat[idx] / (float)__ldg(&count[c]));
  }
}

__global__ void gather_mean_kernel(const float* __restrict__ dvolume, const int64_t* __restrict__ cell,
                                   const int32_t* __restrict__ count, int64_t n, int ch, int64_t cells,
                                   float* __restrict__ dfeat) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = n * ch;
  for (; idx < total; idx += stride) {
    int64_t i = idx / ch;
    int g = (int)(idx - i * ch);
    int64_t c = __ldg(&cell[i]);
    float v = 0.f;
    if (c >= 0 && c < cells) v = __ldg(&dvolume[c * ch + g]) / (float)__ldg(&count[c]);   // same bounds as the forward
    dfeat[idx] = v;
  }
}

}  // namespace

extern "C" {

int pv2_densify_fwd(const float* feat, const int64_t* cell, int64_t n, int c, int64_t cells, float* volume,
                    int32_t* count, void* stream_) {
  PV2_CHECK_ARG(n >= 0 && c > 0 && cells > 0 && volume && count);
  cudaStream_t stream = (cudaStream_t)stream_;
  int64_t vol_elems = cells * c;
  if ((vol_elems % 4) == 0 && ((uintptr_t)volume & 15) == 0)
    zero_f4_kernel<<<pv2_grid_for(vol_elems / 4, 256), 256, 0, stream>>>((float4*)volume, vol_elems / 4);
  else
    cudaMemsetAsync(volume, 0, vol_elems * 4, stream);
  cudaMemsetAsync(count, 0, cells * 4, stream);
  if (n == 0) { PV2_LAUNCH_OK(); return 0; }
  PV2_CHECK_ARG(feat && cell);
  count_kernel<<<pv2_grid_for(n, 256), 256, 0, stream>>>(cell, n, cells, count);
  if ((c % 4) == 0 && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)volume & 15) == 0)
    scatter_mean_v4_kernel<<<pv2_grid_for(n * (c / 4), 256), 256, 0, stream>>>(feat, cell, count, n, c / 4, cells, volume);
  else
    scatter_mean_scalar_kernel<<<pv2_grid_for(n * c, 256), 256, 0, stream>>>(feat, cell, count, n, c, cells, volume);
  PV2_DONE(3);
}

int pv2_densify_bwd(const float* dvolume, const int64_t* cell, const int32_t* count, int64_t n, int c, int64_t cells,
                    float* dfeat, void* stream_) {
  PV2_CHECK_ARG(n >= 0 && c > 0 && cells >= 0);
  if (n == 0) return 0;
  PV2_CHECK_ARG(dvolume && cell && count && dfeat);
  gather_mean_kernel<<<pv2_grid_for(n * c, 256), 256, 0, (cudaStream_t)stream_>>>(dvolume, cell, count, n, c, cells, dfeat);
  PV2_DONE(1);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Weight layout for the data gradient of a sparse convolution: out[ci][k][co] = w[co][flip ? K-1-k : k][ci]
// (the k-flip is the submanifold symmetry nbr[k][j] = i <=> nbr[K-1-k][i] = j, so dgrad reuses the forward map).
// One 32x32 shared-memory tile transpose per (k, tile): coalesced on both sides.  Replaces a torch flip + permute + copy.
namespace {
template <typename T>
__global__ void __launch_bounds__(256) dgrad_weights_kernel(const T* __restrict__ w, T* __restrict__ out, int cout, int kvol,
                                                            int cin, int flip) {
  __shared__ T tile[32][33];
  const int k = blockIdx.z;
  const int ks = flip ? kvol - 1 - k : k;
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty + 8 * i, ci = ci0 + tx;
    if (co < cout && ci < cin) tile[ty + 8 * i][tx] = w[((int64_t)co * kvol + ks) * cin + ci];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty + 8 * i, co = co0 + tx;
    if (ci < cin && co < cout) out[((int64_t)ci * kvol + k) * cout + co] = tile[tx][ty + 8 * i];
  }
}
}  // namespace

extern "C" int pv2_spconv_dgrad_weights(const void* w, void* out, int cout, int kvol, int cin, int flip, int dtype,
                                        void* stream_) {
  PV2_CHECK_ARG(cout > 0 && kvol > 0 && cin > 0 && kvol <= 65535 && w != nullptr && out != nullptr);
  dim3 grid((unsigned)((cin + 31) / 32), (unsigned)((cout + 31) / 32), (unsigned)kvol);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (dtype == PV2_F32)
    dgrad_weights_kernel<float><<<grid, 256, 0, stream>>>((const float*)w, (float*)out, cout, kvol, cin, flip);
  else if (dtype == PV2_BF16)
    dgrad_weights_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)out, cout, kvol, cin, flip);
  else
    return PV2_EUNSUPPORTED;
  PV2_DONE(1);
}
