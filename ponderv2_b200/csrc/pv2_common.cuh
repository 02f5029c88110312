// Shared device/host helpers for libpv2_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/pv2_b200.h"

#define PV2_SM_COUNT 148  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

#define PV2_CHECK_ARG(cond) do { if (!(cond)) return PV2_EINVAL; } while (0)
// launch statistics (relaxed atomic counter, read by bench.py through pv2_launch_count)
extern "C" void pv2_note_launches(int n);
#define PV2_DONE(nlaunch) do { pv2_note_launches(nlaunch); cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; return 0; } while (0)
#define PV2_LAUNCH_OK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

static inline int pv2_grid_for(int64_t work_items, int threads, int max_waves = 8) {
  // enough CTAs to cover the work, capped at max_waves * SMs * (2048/threads) resident CTAs
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)PV2_SM_COUNT * (2048 / threads) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ uint32_t pv2_hash64(uint64_t k) {
  // murmur3 fmix64 finaliser
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (uint32_t)k;
}

__device__ __forceinline__ float pv2_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> struct Pv2Vec4;
