// Sparse convolution as an output-stationary implicit GEMM on the Blackwell tensor cores (tcgen05 + TMEM).
//
//   Y[j, :] = bias + sum_k W_k . X[nbr[k][j], :]        (SubMConv3d / SparseConv3d / SparseInverseConv3d fwd + dgrad;
//                                                         reference call sites spconv_unet_v1m1_base.py:47-66,135-177)
//
// One CTA owns a tile of 128 output rows and all Cout (<= 256) columns.  Which rows form a tile is given by an optional
// `order` permutation (rows with equal neighbour masks adjacent, pv2_rulebook_row_order; `nbr` is then the map in tile
// order, so the 128 indices of an offset are one coalesced line): a (tile, offset) pair in which
// no row has a neighbour is skipped by every role, so mask-sorted tiles run ~K_present instead of K offsets.
// The contraction runs over the concatenated axis (kernel offset k, input channel ci) in chunks of 128 bytes per row:
//   * warps 0-7 are producers.  bf16: all 256 threads issue 16-byte cp.async (LDGSTS, zero-fill for missing neighbours)
//     straight into the 128B-swizzled K-major layout tcgen05.mma consumes.  fp32: the rows are gathered RAW (4 B per
//     element through L2, not a pre-split 8 B copy), split in registers into two TF32 halves and stored with 128-bit
//     st.shared; two groups of four warps alternate chunks so that one group's loads are in flight while the other
//     group splits/stores (measured on B200: 1.5x faster than cp.async + in-place split, which doubles the shared-memory
//     traffic of a kernel whose 3xTF32 MMAs already read 84 KB of operands per chunk).  3xTF32 (hi*hi + lo*hi + hi*lo,
//     ~2^-20 relative error) gives fp32-grade results from the tensor pipe;
//   * warp 8 issues tcgen05.mma (M = 128, N = Cout padded to 16) into a TMEM accumulator, releasing smem stages with
//     tcgen05.commit -> mbarrier;
//   * warps 0-7 then drain TMEM (tcgen05.ld 32x32b), add the bias and write every output row exactly once (no atomics).
//
// The kernel is bound by L2 -> SM traffic (B200: ~6.3 KB/clk chip-wide, ~42 B/clk/SM), hence raw operands, the skipped
// (tile, offset) pairs and the per-CTA chunk rotation that spreads co-resident CTAs over different weight lines.
// HBM-side algorithmic bytes per call: N*Cin*b + N*Cout*b + K*Cin*Cout*b + 4*K*N.
#include "pv2_common.cuh"
#include "umma.cuh"
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <stdlib.h>
#include <string.h>

extern "C" int pv2_get_option(const char* name);   // capi.cu

namespace {

using namespace pv2;

constexpr int kTileM = 128;
constexpr int kMaxProducerWarps = 12;   // fp32: kGroups (2 or 3) groups of 4 producer warps; bf16: 8 warps
constexpr int kABytes = kTileM * 128;  // one operand tile: 128 rows x 128 B
constexpr int kMaxStages = 6;

struct GGParams {
  const void* x;
  const void* w;
  int64_t w_sco, w_sk;
  const float* bias;
  const int32_t* nbr;
  const int32_t* order;  // optional [n_out]: tile position -> output row
  void* y;
  int64_t n_out;
  int cin, cout, kvol;
  int n_pad;       // cout rounded up to 16
  int num_chunks;  // ceil(kvol * cin / elems_per_row)
  int stages;
  uint32_t tmem_cols;
  int64_t x_row;     // elements between consecutive rows of x
  uint32_t x_row32, w_sco32, w_sk32;   // the same strides as 32-bit factors (one IMAD.WIDE.U32 per address)
  int64_t x_lo_off;  // fp32: != 0 -> x is split-precision, value = x[..] + x[.. + x_lo_off]
  // output: row stride, optional split-precision output (hi at y, lo at y + y_lo_off) and the fused epilogue
  int64_t y_row, y_lo_off;
  int y_split;       // fp32 only: write TF32 hi/lo halves instead of the plain value
  int act;           // 0: none; 1: y = softplus(beta=100, threshold=20)(v), y2 = sigmoid(100 v)   (SDF decoder, decoders.py:24)
                     // 2..4: backward epilogues reading y2 / the old y (see the epilogue)
  void* y2;
  int64_t y2_row, y2_lo_off;
  int ablate;        // development only (PV2_GG_ABLATE): 1 skip the global loads, 2 skip split + st.shared, 4 skip the MMAs
  int meta_bufs;     // persistent fp32 kernel: metadata buffers (2 or 3)
  int ksplit;        // > 1: gridDim.y CTAs share one row tile, each reduces a slice of the contraction and adds its
                     // partial result into the (pre-zeroed) output with red.global.add (small deep U-Net levels)
  // bf16x3 mode (persistent kernel): weights pre-split into bf16 hi / lo matrices [2][w2_rows][ktot64], fetched by TMA
  int w2_row0;       // first weight row (output channel) of this launch's column slice
  int w2_rows;       // rows of one half (hi rows [0, w2_rows), lo rows [w2_rows, 2 w2_rows))
};

template <bool kSplit>
struct ModeTraits;
template <>
struct ModeTraits<false> {  // bf16 storage, kind::f16
  static constexpr int kEltBytes = 2, kEPR = 64, kEPP = 8, kFmt = 1, kOperands = 1;
  using Elt = __nv_bfloat16;
};
template <>
struct ModeTraits<true> {  // fp32 storage, 3xTF32
  static constexpr int kEltBytes = 4, kEPR = 32, kEPP = 4, kFmt = 2, kOperands = 2;
  using Elt = float;
};

// v = hi + lo + O(2^-24 |v|): both halves rounded to nearest TF32 (10-bit mantissa), so the tensor core sees exact inputs
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
  lo = __uint_as_float(l);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// split a float4 and store the halves at `addr` (hi tile) and `addr + lo_delta` (lo tile): hi = top 19 bits (exactly a
// TF32 number), lo = v - hi (exact in fp32; the tensor core keeps its leading bits)
__device__ __forceinline__ void split_store(uint32_t addr, uint32_t lo_delta, const float4& v) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
  st_shared_v4(addr, h);
  st_shared_v4(addr + lo_delta, l);
}

// kPre (fp32 only): x is stored split-precision (value = x[..] + x[.. + x_lo_off]); a template parameter because a
// predicated-off FADD on a just-loaded register still waits for the load, which serialises the gather.
// fp32 output of one row's 16 accumulator columns [col0, col0 + 16): fused activation / backward epilogues, then plain or
// split-precision stores.  (bias already added)
// operands of the backward epilogues (act >= 2) for one row's columns [col0, col0 + 16): s = y2 (act 2, 3), y_old (3, 4)
__device__ __forceinline__ void epilogue_operands(const GGParams& p, int64_t j, int col0, float (&sa)[16], float (&so)[16]) {
  const float* ar = reinterpret_cast<const float*>(p.y2) + j * p.y2_row + col0;
  const float* yo = reinterpret_cast<const float*>(p.y) + j * p.y_row + col0;
  const bool vec = (col0 + 16 <= p.cout) && ((p.cout & 3) == 0) && ((p.y_row & 3) == 0) && ((p.y2_row & 3) == 0);
  if (vec) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), o = a;
      if (p.act != 4) a = __ldg(reinterpret_cast<const float4*>(ar) + q);
      if (p.act != 2) o = *(reinterpret_cast<const float4*>(yo) + q);
      sa[4 * q] = a.x; sa[4 * q + 1] = a.y; sa[4 * q + 2] = a.z; sa[4 * q + 3] = a.w;
      so[4 * q] = o.x; so[4 * q + 1] = o.y; so[4 * q + 2] = o.z; so[4 * q + 3] = o.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const bool in = col0 + i < p.cout;
      sa[i] = (in && p.act != 4) ? __ldg(ar + i) : 0.f;
      so[i] = (in && p.act != 2) ? yo[i] : 0.f;
    }
  }
}

// pre_a / pre_y: the act >= 2 operands already in registers (the persistent kernel fetches them one column block ahead).
// kActs = false compiles the activation epilogues out (sparse convolutions: act is always 0).
template <bool kActs = true>
__device__ __forceinline__ void fp32_row_epilogue(const GGParams& p, float (&f)[16], int64_t j, int col0,
                                                  const float* pre_a = nullptr, const float* pre_y = nullptr) {
  float g2[16];
  if (kActs && p.act == 1) {
    // softplus(beta = 100) and its derivative sigmoid(100 h) from ONE fast exponential: e = exp(-|t|) in (0, 1];
    // softplus = max(h, 0) + log(1 + e) / 100, sigmoid = (t >= 0 ? 1 : e) / (1 + e).  (ex2 / lg2 approximations: absolute
    // error of log(1 + e) <= 2e-7, i.e. 2e-9 on h-scale values; the libm expf + log1pf pair cost ~150 instructions per
    // element and made the four epilogue warps the bottleneck of the layer: 910 us for 524 k rows, profiles/r2za_launches.csv)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float t = 100.f * f[i];
      const float e = __expf(-fabsf(t));
      const float r = __fdividef(1.f, 1.f + e);
      g2[i] = (t >= 0.f) ? r : e * r;
      f[i] = fmaxf(f[i], 0.f) + __logf(1.f + e) * 0.01f;
    }
  }
  const bool vec = (col0 + 16 <= p.cout) && ((p.cout & 3) == 0);
  if (kActs && p.act >= 2) {
    // backward-pass epilogues of the SDF decoder (y2 is an INPUT here, plain fp32, row stride y2_row):
    //   2: y = v * 100 s (1 - s)       (through the softplus derivative s = sigmoid(100 h))
    //   3: y = y_old + v * s           4: y = y_old + v
    float sa[16], so[16];
    if (pre_a != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { sa[i] = pre_a[i]; so[i] = pre_y[i]; }
    } else {
      epilogue_operands(p, j, col0, sa, so);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (col0 + i < p.cout) {
        if (p.act == 2) f[i] *= 100.f * sa[i] * (1.f - sa[i]);
        else if (p.act == 3) f[i] = so[i] + f[i] * sa[i];
        else f[i] += so[i];
      }
    }
  }
  auto store = [&](float* base, int64_t row_stride, int64_t lo_off, const float (&val)[16], bool split) {
    float* yr = base + j * row_stride + col0;
    if (!split) {
      if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(yr + 4 * q) = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
      } else {
        for (int i = 0; i < 16 && col0 + i < p.cout; ++i) yr[i] = val[i];
      }
    } else {
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) split_tf32(val[i], hi[i], lo[i]);
      if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<float4*>(yr + 4 * q) = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          *reinterpret_cast<float4*>(yr + lo_off + 4 * q) = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
      } else {
        for (int i = 0; i < 16 && col0 + i < p.cout; ++i) { yr[i] = hi[i]; yr[lo_off + i] = lo[i]; }
      }
    }
  };
  store(reinterpret_cast<float*>(p.y), p.y_row, p.y_lo_off, f, p.y_split != 0);
  if (kActs && p.act == 1 && p.y2 != nullptr) store(reinterpret_cast<float*>(p.y2), p.y2_row, p.y2_lo_off, g2, p.y_split != 0);
}

// 16 zero bytes in global memory: missing neighbours / padding rows load from here instead of predicating the load and
// zero-filling registers (fewer instructions in the gather prologue)
__device__ float4 g_zero_page[2];

// kGroups (fp32): groups of four producer warps that alternate chunks (group g takes chunks g, g + kGroups, ...).
template <bool kSplit, bool kPre, int kGroups>
__global__ void __launch_bounds__(kGroups * 128 + 32) umma_gather_gemm_kernel(const GGParams p) {
  using T = ModeTraits<kSplit>;
  constexpr int kProducerWarps = kGroups * 4;
  constexpr int kProducerThreads = kProducerWarps * 32;
  constexpr int kThreads = kProducerThreads + 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16 B aligned: round up to 1024 (the launch reserves the slack)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  // Tiles are taken in reverse: in mask-sorted order the rows with the most neighbours (highest masks, most chunks) come
  // last, and the longest tiles should start first so that the last wave is short ones.
  const int64_t row0 = (int64_t)(gridDim.x - 1 - blockIdx.x) * kTileM;

  const int b_bytes = p.n_pad * 128;
  const int stage_bytes = (kABytes + b_bytes) * T::kOperands;
  uint8_t* stage_base = smem;
  int32_t* idx_s = reinterpret_cast<int32_t*>(smem + (size_t)p.stages * stage_bytes);
  int32_t* row_s = idx_s + p.kvol * kTileM;
  uint16_t* active = reinterpret_cast<uint16_t*>(row_s + kTileM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(active + p.num_chunks) + 7) & ~uintptr_t(7));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
  uint32_t* kmask_s = tmem_slot + 1;                       // [4]: one bit per kernel offset (K <= 128)
  int* wcount_s = reinterpret_cast<int*>(tmem_slot + 5);  // [kMaxProducerWarps + 1]
  uint32_t* kc_s = reinterpret_cast<uint32_t*>(wcount_s + kMaxProducerWarps + 1);   // [num_chunks][8]: (k << 16 | ci) per 16-byte piece

  // ---- setup ---------------------------------------------------------------------------------------------
  if (tid < kTileM) {
    const int64_t pos = row0 + tid;
    int32_t j = -1;
    if (pos < p.n_out) j = (p.order != nullptr) ? __ldg(&p.order[pos]) : (int32_t)pos;
    row_s[tid] = j;
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), kSplit ? 128 : kProducerThreads);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    kmask_s[0] = kmask_s[1] = kmask_s[2] = kmask_s[3] = 0;
    fence_mbar_init();
  }
  if (warp == kProducerWarps) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // neighbour indices of the tile (nbr is in tile order: 128 consecutive entries per offset, coalesced); loads are
  // issued in batches of 6 before the first store waits
  {
    const int total = p.kvol * kTileM;
    for (int i0 = tid; i0 < total; i0 += 6 * kThreads) {
      int32_t v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * kThreads;
        v[u] = -1;
        if (i < total) {
          const int64_t pos = row0 + (i & (kTileM - 1));
          if (pos < p.n_out) v[u] = (p.nbr != nullptr) ? __ldg(&p.nbr[(int64_t)(i >> 7) * p.n_out + pos]) : (int32_t)pos;
        }
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * kThreads;
        if (i < total) idx_s[i] = v[u];
      }
    }
  }
  // (offset, channel) of every 16-byte piece of every chunk, computed once (the divisions stay out of the main loop)
  {
    const int ktot_ = p.kvol * p.cin;
    for (int i = tid; i < p.num_chunks * 8; i += kThreads) {
      const int e0 = (i >> 3) * T::kEPR + (i & 7) * T::kEPP;
      uint32_t v = 0xffffffffu;
      if (e0 < ktot_) { const int k = e0 / p.cin; v = ((uint32_t)k << 16) | (uint32_t)(e0 - k * p.cin); }
      kc_s[i] = v;
    }
  }
  __syncthreads();
  // which kernel offsets have at least one neighbour in this tile
  if (warp < kProducerWarps) {
    for (int k = warp; k < p.kvol; k += kProducerWarps) {
      const int32_t* r = idx_s + k * kTileM + lane;
      const bool a = (r[0] >= 0) | (r[32] >= 0) | (r[64] >= 0) | (r[96] >= 0);
      if (__any_sync(0xffffffffu, a) && lane == 0) atomicOr(&kmask_s[k >> 5], 1u << (k & 31));
    }
  }
  __syncthreads();
  // compact the chunks that touch at least one active offset (all threads take part in the barriers)
  int n_total = 0;
  {
    const uint32_t km[4] = {kmask_s[0], kmask_s[1], kmask_s[2], kmask_s[3]};
    for (int c0 = 0; c0 < p.num_chunks; c0 += kProducerThreads) {
      const int c = c0 + tid;
      bool on = false;
      if (tid < kProducerThreads && c < p.num_chunks) {
        const int k_lo = (c * T::kEPR) / p.cin;
        int k_hi = ((c + 1) * T::kEPR - 1) / p.cin;
        if (k_hi >= p.kvol) k_hi = p.kvol - 1;
        for (int k = k_lo; k <= k_hi; ++k) on |= ((km[k >> 5] >> (k & 31)) & 1u) != 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, on);
      if (lane == 0 && warp < kProducerWarps) wcount_s[warp] = __popc(bal);
      __syncthreads();
      int before = 0, all = 0;
#pragma unroll
      for (int w2 = 0; w2 < kProducerWarps; ++w2) {
        const int cnt = wcount_s[w2];
        before += (w2 < warp) ? cnt : 0;
        all += cnt;
      }
      if (on) active[n_total + before + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)c;
      n_total += all;
      __syncthreads();
    }
  }
  // Every CTA streams the same weight chunks; starting each tile at a different chunk spreads the L2 slices the
  // co-resident CTAs hit at any instant (the order of accumulation is irrelevant to the sum).  With split-K
  // (gridDim.y > 1) each CTA takes a contiguous slice of the rotated list.
  int first = 0, n_active = n_total;
  if (p.ksplit > 1) {
    const int per = (n_total + p.ksplit - 1) / p.ksplit;
    first = (int)blockIdx.y * per;
    n_active = n_total - first;
    if (n_active > per) n_active = per;
    if (n_active < 0) n_active = 0;
  }
  const int rot = (n_total > 1) ? (int)(((unsigned)(row0 >> 7) * 11u) % (unsigned)n_total) : 0;
  auto chunk_at = [&](int it) -> int {
    int q = first + it + rot;
    if (q >= n_total) q -= n_total;
    return (int)active[q];
  };

  if (warp < kProducerWarps) {
    // ======================= producers =======================
    using E = typename T::Elt;
    const E* x = reinterpret_cast<const E*>(p.x);
    const E* w = reinterpret_cast<const E*>(p.w);
    if constexpr (kSplit) {
      const int grp = warp >> 2;       // chunks it = grp, grp + 2, ...
      const int tg = tid & 127;
      const int piece = tg & 7;
      const int rbase = tg >> 3;       // 0..15; rows rbase + 16 i keep r & 7, so the swizzled offset advances 2048 B per i
      const uint32_t tile_off = sw128_offset(rbase, piece);
      const int nb = p.n_pad >> 4;     // weight rows per thread (1..16)
      const float* wrow0 = w + (uint64_t)rbase * p.w_sco32;   // loop-invariant: weight row rbase
      const uint64_t wstep = 16ull * p.w_sco32;                // rows rbase + 16 i
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* zpage = reinterpret_cast<const float*>(g_zero_page);
      for (int it = grp; it < n_active; it += kGroups) {
        const int c = chunk_at(it);
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        const uint32_t kc = kc_s[c * 8 + piece];
        const bool kvalid = kc != 0xffffffffu;
        const uint32_t k = kvalid ? kc >> 16 : 0u;
        const uint32_t ci = kvalid ? (kc & 0xffffu) : 0u;
        const int32_t* idx_k = idx_s + k * kTileM;
        // every global load of the chunk is issued before anything waits
        float4 va[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int32_t src = kvalid ? idx_k[rbase + 16 * i] : -1;
          const float* g = (src >= 0 && !(p.ablate & 1)) ? x + ((uint64_t)(uint32_t)src * p.x_row32 + ci) : zpage;
          va[i] = __ldg(reinterpret_cast<const float4*>(g));
        }
        if constexpr (kPre) {   // split-precision input: add the lo halves (second batch of loads, then the adds)
          float4 vl[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int32_t src = kvalid ? idx_k[rbase + 16 * i] : -1;
            vl[i] = zero4;
            if (src >= 0) vl[i] = __ldg(reinterpret_cast<const float4*>(x + ((uint64_t)(uint32_t)src * p.x_row32 + ci) + p.x_lo_off));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { va[i].x += vl[i].x; va[i].y += vl[i].y; va[i].z += vl[i].z; va[i].w += vl[i].w; }
        }
        const float* wk = wrow0 + ((uint64_t)k * p.w_sk32 + ci);   // weight row rbase, this piece
        float4 vb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = rbase + 16 * i;
          const float* g = (kvalid && i < nb && n < p.cout && !(p.ablate & 1)) ? wk + (uint64_t)i * wstep : zpage;
          vb[i] = __ldg(reinterpret_cast<const float4*>(g));
        }
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t a_dst = smem_u32(stage_base + (size_t)s * stage_bytes) + tile_off;
        const uint32_t b_dst = a_dst + 2 * kABytes;
        if (!(p.ablate & 2)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) split_store(a_dst + i * 2048, kABytes, va[i]);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nb) split_store(b_dst + i * 2048, (uint32_t)b_bytes, vb[i]);
        } else if (va[0].x == 123.456f && vb[0].x == 654.321f) {
          st_shared_v4(a_dst, va[1]);   // keep the loads alive
        }
        if (nb > 8) {  // wide layers (Cout > 128): second half of the weight rows
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = rbase + 16 * (i + 8);
            const float* g = (kvalid && i + 8 < nb && n < p.cout) ? wk + (uint64_t)(i + 8) * wstep : zpage;
            vb[i] = __ldg(reinterpret_cast<const float4*>(g));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i + 8 < nb) split_store(b_dst + (i + 8) * 2048, (uint32_t)b_bytes, vb[i]);
        }
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
        mbar_arrive(smem_u32(&full_bar[s]));
      }
    } else {
      const int piece = tid & 7;
      const int rbase = tid >> 3;      // 0..31; rows rbase + 32 i -> +4096 B per i
      const uint32_t tile_off = sw128_offset(rbase, piece);
      // A stage is handed to the MMA warp `lag` iterations after its copies were issued and is refilled `stages`
    // iterations later: lag = stages / 2 leaves the copies and the MMAs (issue -> commit -> barrier) about half of the
    // ring each to complete in, instead of making every refill wait for MMAs issued one iteration earlier.
    const int lag = p.stages >= 2 ? p.stages / 2 : 1;
      for (int it = 0; it < n_active + lag; ++it) {
        if (it < n_active) {
          const int c = chunk_at(it);
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          uint8_t* a_tile = stage_base + (size_t)s * stage_bytes;
          const uint32_t kc = kc_s[c * 8 + piece];
          const bool kvalid = kc != 0xffffffffu;
          const uint32_t k = kvalid ? kc >> 16 : 0u;
          const uint32_t ci = kvalid ? (kc & 0xffffu) : 0u;
          const int32_t* idx_k = idx_s + k * kTileM;
          int32_t src[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) src[i] = kvalid ? idx_k[rbase + 32 * i] : -1;
          const uint32_t a_dst = smem_u32(a_tile) + tile_off;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const E* g = (src[i] >= 0) ? x + ((uint64_t)(uint32_t)src[i] * p.x_row32 + ci) : x;
            cp_async_16(a_dst + i * 4096, g, src[i] >= 0 ? 16u : 0u);
          }
          const uint32_t b_dst = a_dst + kABytes;
          const E* wk = w + ((uint64_t)rbase * p.w_sco32 + (uint64_t)k * p.w_sk32 + ci);
          for (int n = rbase, i = 0; n < p.n_pad; n += 32, ++i) {
            const bool ok = kvalid && n < p.cout;
            cp_async_16(b_dst + i * 4096, ok ? wk + (uint64_t)i * (32ull * p.w_sco32) : w, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (it >= lag) {
          // chunk (it - lag) has landed for this thread: make it visible to the async proxy and signal
          switch (lag) {  // wait_group needs an immediate
            case 1: cp_async_wait<1>(); break;
            case 2: cp_async_wait<2>(); break;
            case 3: cp_async_wait<3>(); break;
            case 4: cp_async_wait<4>(); break;
            default: cp_async_wait<5>(); break;
          }
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&full_bar[(it - lag) % p.stages]));
        }
      }
    }

    // ======================= epilogue =======================
    if (n_active > 0) {
      mbar_wait(smem_u32(tmem_full_bar), 0);
      tc_fence_after();
    }
    const int lane_grp = warp & 3;            // TMEM lanes 32 * (warp % 4) .. + 31 are the ones this warp may read
    const int32_t j32 = row_s[lane_grp * 32 + lane];
    const int64_t j = j32;
    const bool row_ok = j32 >= 0;
    for (int col0 = (warp >> 2) * 16; col0 < p.n_pad; col0 += 16 * kGroups) {
      uint32_t v[16];
      if (n_active > 0) {
        tmem_ld_x16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)col0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      if (!row_ok) continue;
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int co = col0 + i;
        f[i] = __uint_as_float(v[i]) + ((p.bias != nullptr && co < p.cout && blockIdx.y == 0) ? __ldg(&p.bias[co]) : 0.f);
      }
      if constexpr (kSplit) {
        if (p.ksplit > 1) {   // split-K partial sums (act == 0 on this path): vector reductions into the pre-zeroed rows
          const bool vec = (col0 + 16 <= p.cout) && ((p.cout & 3) == 0);
          float* yr = reinterpret_cast<float*>(p.y) + j * p.y_row + col0;
          if (n_active > 0) {
            if (vec) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yr + 4 * q), "f"(f[4 * q]),
                             "f"(f[4 * q + 1]), "f"(f[4 * q + 2]), "f"(f[4 * q + 3]) : "memory");
            } else {
              for (int i = 0; i < 16 && col0 + i < p.cout; ++i) atomicAdd(yr + i, f[i]);
            }
          } else if (blockIdx.y == 0 && p.bias != nullptr) {
            for (int i = 0; i < 16 && col0 + i < p.cout; ++i) atomicAdd(yr + i, f[i]);
          }
          continue;
        }
        fp32_row_epilogue(p, f, j, col0);
      } else {
        __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(p.y) + j * p.y_row + col0;
        if (col0 + 16 <= p.cout && (p.cout & 7) == 0) {
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
            pk[q] = *reinterpret_cast<uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(yr) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(yr + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
          for (int i = 0; i < 16 && col0 + i < p.cout; ++i) yr[i] = __float2bfloat16(f[i]);
        }
      }
    }
    tc_fence_before();
  } else {
    // ======================= MMA issuer (warp 8) =======================
    const uint32_t idesc = make_idesc(T::kFmt, kTileM, p.n_pad);
    for (int it = 0; it < n_active; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      tc_fence_after();
      if (lane == 0) {
        // descriptors differ between k-steps only in the start-address field (16-byte units): +2 per 32 B
        const uint64_t da = smem_desc_kmajor_sw128(smem_u32(stage_base + (size_t)s * stage_bytes));
        if constexpr (!kSplit) {
          const uint64_t db = da + (uint64_t)(kABytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base, da + 2 * ks, db + 2 * ks, idesc, (it | ks) != 0 ? 1u : 0u);
        } else {
          const uint64_t dal = da + (uint64_t)(kABytes >> 4), dbh = dal + (uint64_t)(kABytes >> 4);
          const uint64_t dbl = dbh + (uint64_t)(b_bytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (p.ablate & 4) break;
            umma_tf32(tmem_base, dal + 2 * ks, dbh + 2 * ks, idesc, (it | ks) != 0 ? 1u : 0u);
            umma_tf32(tmem_base, da + 2 * ks, dbl + 2 * ks, idesc, 1u);
            umma_tf32(tmem_base, da + 2 * ks, dbh + 2 * ks, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));               // smem stage reusable once these MMAs retire
        if (it == n_active - 1) umma_commit(smem_u32(tmem_full_bar));  // accumulator complete
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ===============================================================================================================
// Persistent variant (fp32, one CTA per SM looping over its tiles).  Measured on B200 with everything but the skeleton
// switched off (profiles/r1j_gather_gemm_ablation.txt), the one-tile-per-CTA kernel above still spends half of its time:
// per tile the index fetch, mask compaction, pipeline fill, accumulator drain and CTA turnover are serialised and nothing
// else runs on the SM meanwhile.  Here the roles are decoupled and pipelined ACROSS tiles:
//   * 4 epilogue warps prepare the metadata (rows, neighbour indices, active-chunk list) of tile i + 1 into the second
//     metadata buffer while tile i is in the main loop, then drain tile i's accumulator;
//   * producer groups and the MMA warp run the chunks of consecutive tiles back to back over one smem stage ring
//     (a global chunk counter continues the ring and the group alternation across tiles);
//   * two TMEM accumulators alternate, so the drain of tile i overlaps the MMAs of tile i + 1.
constexpr int kPersistGroups = 2;
constexpr int kPersistProducerWarps = kPersistGroups * 4;
constexpr int kPersistThreads = (kPersistProducerWarps + 1 + 4) * 32;   // producers | MMA warp | 4 epilogue warps

struct PersistLayout {
  int meta_bytes;   // one metadata buffer: idx[kvol][128] | row[128] | active[num_chunks] | n_active
  int fixed;        // everything except the stage ring
};
__host__ __device__ inline PersistLayout persist_layout(int kvol, int num_chunks, int nbuf = 2) {
  PersistLayout L;
  L.meta_bytes = (kvol * kTileM * 4 + kTileM * 4 + num_chunks * 2 + 16 + 15) / 16 * 16;
  L.fixed = nbuf * L.meta_bytes + num_chunks * 32 /*kc table*/ + 32 * 8 /*barriers*/ + 64 + 1024 /*alignment*/;
  return L;
}
constexpr int kMetaBufs = 3;   // metadata buffers of the fp32 persistent kernel (prepared two tiles ahead)

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// fp32 -> two bf16 halves (round to nearest): v = hi + lo + O(2^-18 |v|).  Packs 8 values into one 16-byte piece each.
__device__ __forceinline__ void split_store_bf16(uint32_t addr, uint32_t lo_delta, const float4& v0, const float4& v1) {
  const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t hp[4], lp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
    hp[i] = *reinterpret_cast<const uint32_t*>(&h2);
    lp[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hp[0]), "r"(hp[1]), "r"(hp[2]), "r"(hp[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + lo_delta), "r"(lp[0]), "r"(lp[1]), "r"(lp[2]), "r"(lp[3]) : "memory");
}

// kBx3 = false: 3xTF32 (operands split into TF32 halves, 32 channels per 128-byte chunk, weights through registers).
// kBx3 = true : bf16x3 -- fp32 storage, every operand split on chip into two bf16 halves (hi + lo carries 16 significand
//   bits, products hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM).  Against 3xTF32 this halves the shared-memory bytes
//   per channel (the pipe that bounded the TF32 kernel: 57.6 % L1/TEX, profiles/r1m) and doubles the tensor rate, at a
//   per-product rounding of <= 3 * 2^-18 instead of 2^-20; in both modes the observed error is set by the tensor core's
//   truncating fp32 accumulation (~ n_steps * 2^-24).  The weight operand is pre-split once per call into bf16 hi / lo
//   matrices [Cout][K * Cin] and arrives by TMA (one 2-D box of n_pad rows x 128 B per half and chunk, hardware swizzle),
//   so the producer warps only gather activations.
// kG = producer groups (128 threads each; a group has one chunk in flight).  More groups did not help (r2p): the gathers
// were not the critical path, the metadata preparation was (see the epilogue group below).
// kNB = metadata buffers: 3 (prepared two tiles ahead) or 2 (wide layers, where a pipeline stage is worth more).
// kLin = dense linear layer (identity row map, kvol = 1, fused activation epilogues of the render MLP); false = sparse
// convolution (act = 0): the two have different register budgets in the epilogue group (27 neighbour indices in flight per
// row vs. prefetched epilogue operands), so they are separate instantiations.
template <bool kBx3, int kG, int kNB, bool kLin>
__global__ void __launch_bounds__((kG * 4 + 5) * 32) umma_gather_gemm_persistent_kernel(const GGParams p, int num_tiles,
                                                                                        const __grid_constant__ CUtensorMap wmap) {
  constexpr int kPersistGroups = kG;
  constexpr int kPersistProducerWarps = kG * 4;
  constexpr int kPersistThreads = (kG * 4 + 5) * 32;   // producers | MMA warp | 4 epilogue warps
  constexpr int kEPR = kBx3 ? 64 : 32;   // contraction elements per 128-byte operand row (chunk)
  constexpr int kEPP = kBx3 ? 8 : 4;     // ... per 16-byte piece
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b_bytes = p.n_pad * 128;
  const int stage_bytes = (kABytes + b_bytes) * 2;
  constexpr int kMB = kNB;   // metadata buffers (NOT `nb`: the producers use that name for the 16-row weight blocks)
  const PersistLayout L = persist_layout(p.kvol, p.num_chunks, kMB);
  uint8_t* stage_base = smem;
  uint8_t* meta_base = smem + (size_t)p.stages * stage_bytes;
  uint32_t* kc_s = reinterpret_cast<uint32_t*>(meta_base + kMB * L.meta_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(kc_s + p.num_chunks * 8);
  uint64_t* full_bar = bars;                        // [kMaxStages]  128 arrivals (the producing group)
  uint64_t* empty_bar = bars + kMaxStages;          // [kMaxStages]  tcgen05.commit
  uint64_t* meta_full = bars + 2 * kMaxStages;      // [kMetaBufs]   1 arrival (epilogue group leader)
  uint64_t* meta_empty = meta_full + kMetaBufs;     // [kMetaBufs]   producer warps + MMA warp
  uint64_t* tmem_full = meta_empty + kMetaBufs;     // [2]           tcgen05.commit / plain arrive for empty tiles
  uint64_t* tmem_empty = tmem_full + 2;             // [2]           4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* scratch = tmem_slot + 1;                // [16] epilogue-group scratch (kmask[4], warp counts[4], total)

  auto meta_idx = [&](int b) { return reinterpret_cast<int32_t*>(meta_base + (size_t)b * L.meta_bytes); };
  auto meta_row = [&](int b) { return meta_idx(b) + p.kvol * kTileM; };
  auto meta_active = [&](int b) { return reinterpret_cast<uint16_t*>(meta_row(b) + kTileM); };
  auto meta_count = [&](int b) { return reinterpret_cast<int*>(meta_base + (size_t)b * L.meta_bytes + L.meta_bytes - 16); };

  // split-K (p.ksplit > 1, the small deep levels): `num_tiles` counts VIRTUAL tiles = (row tile, contraction slice); a
  // slice takes a contiguous share of the tile's active-chunk list and adds its partial sums into the pre-zeroed rows
  const int ks = p.ksplit > 1 ? p.ksplit : 1;
  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // longest tiles (highest masks, last in tile order) first
  auto vtile = [&](int i) { return num_tiles - 1 - ((int)blockIdx.x + i * (int)gridDim.x); };
  auto tile_row0 = [&](int i) { return (int64_t)(vtile(i) / ks) * kTileM; };

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 128); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < kMetaBufs; ++b) {
      mbar_init(smem_u32(&meta_full[b]), 1);
      mbar_init(smem_u32(&meta_empty[b]), kPersistProducerWarps + 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 4);
    }
    fence_mbar_init();
  }
  if (warp == kPersistProducerWarps) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  {  // (offset, channel) of every 16-byte piece of every chunk: tile-independent
    const int ktot = p.kvol * p.cin;
    for (int i = tid; i < p.num_chunks * 8; i += kPersistThreads) {
      const int e0 = (i >> 3) * kEPR + (i & 7) * kEPP;
      uint32_t v = 0xffffffffu;
      if (e0 < ktot) { const int k = e0 / p.cin; v = ((uint32_t)k << 16) | (uint32_t)(e0 - k * p.cin); }
      kc_s[i] = v;
    }
  }
  if (kBx3 && tid == 0) tma_prefetch_desc(&wmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_stride = p.tmem_cols >> 1;   // two accumulators

  if (warp < kPersistProducerWarps) {
    // ======================= producers =======================
    const float* x = reinterpret_cast<const float*>(p.x);
    const float* w = reinterpret_cast<const float*>(p.w);
    const int grp = warp >> 2;
    const int tg = tid & 127;
    const int piece = tg & 7, rbase = tg >> 3;
    const uint32_t tile_off = sw128_offset(rbase, piece);
    const int nb = p.n_pad >> 4;
    const float* zpage = reinterpret_cast<const float*>(g_zero_page);
    const float* wrow0 = w + (uint64_t)rbase * p.w_sco32;
    const uint64_t wstep = 16ull * p.w_sco32;
    int gbase = 0;   // global chunk counter at the start of the tile (same in every thread)
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i % kMB;
      mbar_wait(smem_u32(&meta_full[b]), (uint32_t)(i / kMB) & 1u);
      const int n_active = *meta_count(b);
      const int32_t* idx_s = meta_idx(b);
      const uint16_t* active = meta_active(b) + meta_count(b)[1];   // this slice's share of the list
      int it = (grp - gbase % kPersistGroups + kPersistGroups) % kPersistGroups;   // first chunk of this group in the tile
      for (; it < n_active; it += kPersistGroups) {
        const int g = gbase + it;
        const int s = g % p.stages;
        const uint32_t ph = (uint32_t)(g / p.stages) & 1u;
        const uint32_t kc = kc_s[(int)active[it] * 8 + piece];
        const bool kvalid = kc != 0xffffffffu;
        const uint32_t k = kvalid ? kc >> 16 : 0u;
        const uint32_t ci = kvalid ? (kc & 0xffffu) : 0u;
        const int32_t* idx_k = idx_s + k * kTileM;
        if constexpr (kBx3) {
          // 8 rows x one 8-channel piece (32 B of fp32 = two 128-bit loads) per thread; every load is issued before
          // the stage wait, the split + stores follow it
          float4 va[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int32_t src = kvalid ? idx_k[rbase + 16 * q] : -1;
            const float* gp = (src >= 0) ? x + ((uint64_t)(uint32_t)src * p.x_row32 + ci) : zpage;
            va[2 * q] = __ldg(reinterpret_cast<const float4*>(gp));
            va[2 * q + 1] = __ldg(reinterpret_cast<const float4*>(gp) + 1);
          }
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t a_dst = smem_u32(stage_base + (size_t)s * stage_bytes) + tile_off;
          if (tg == 0) {   // the weight halves of this chunk: two TMA boxes (n_pad rows x 128 B), hardware-swizzled
            const uint32_t bar = smem_u32(&full_bar[s]);
            const uint32_t b_dst0 = smem_u32(stage_base + (size_t)s * stage_bytes) + 2 * kABytes;
            mbar_expect_tx(bar, 2u * (uint32_t)b_bytes);
            const int32_t col = (int32_t)active[it] * kEPR;
            tma_load_2d(b_dst0, &wmap, bar, col, p.w2_row0);
            tma_load_2d(b_dst0 + (uint32_t)b_bytes, &wmap, bar, col, p.w2_rows + p.w2_row0);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) split_store_bf16(a_dst + q * 2048, kABytes, va[2 * q], va[2 * q + 1]);
        } else {
        float4 va[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int32_t src = kvalid ? idx_k[rbase + 16 * q] : -1;
          const float* gp = (src >= 0) ? x + ((uint64_t)(uint32_t)src * p.x_row32 + ci) : zpage;
          va[q] = __ldg(reinterpret_cast<const float4*>(gp));
        }
        const float* wk = wrow0 + ((uint64_t)k * p.w_sk32 + ci);
        float4 vb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = rbase + 16 * q;
          const float* gp = (kvalid && q < nb && n < p.cout) ? wk + (uint64_t)q * wstep : zpage;
          vb[q] = __ldg(reinterpret_cast<const float4*>(gp));
        }
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t a_dst = smem_u32(stage_base + (size_t)s * stage_bytes) + tile_off;
        const uint32_t b_dst = a_dst + 2 * kABytes;
#pragma unroll
        for (int q = 0; q < 8; ++q) split_store(a_dst + q * 2048, kABytes, va[q]);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q < nb) split_store(b_dst + q * 2048, (uint32_t)b_bytes, vb[q]);
        if (nb > 8) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int n = rbase + 16 * (q + 8);
            const float* gp = (kvalid && q + 8 < nb && n < p.cout) ? wk + (uint64_t)(q + 8) * wstep : zpage;
            vb[q] = __ldg(reinterpret_cast<const float4*>(gp));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q + 8 < nb) split_store(b_dst + (q + 8) * 2048, (uint32_t)b_bytes, vb[q]);
        }
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&full_bar[s]));
      }
      gbase += n_active;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&meta_empty[b]));   // this warp no longer reads metadata buffer b
    }
  } else if (warp == kPersistProducerWarps) {
    // ======================= MMA issuer =======================
    const uint32_t idesc = make_idesc(kBx3 ? 1 : 2, kTileM, p.n_pad);
    int gbase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1, mb = i % kMB;
      mbar_wait(smem_u32(&meta_full[mb]), (uint32_t)(i / kMB) & 1u);
      const int n_active = *meta_count(mb);
      mbar_wait(smem_u32(&tmem_empty[b]), ((uint32_t)(i >> 1) & 1u) ^ 1u);   // accumulator b drained (first two: free)
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)b * acc_stride;
      for (int it = 0; it < n_active; ++it) {
        const int g = gbase + it;
        const int s = g % p.stages;
        const uint32_t ph = (uint32_t)(g / p.stages) & 1u;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = smem_desc_kmajor_sw128(smem_u32(stage_base + (size_t)s * stage_bytes));
          const uint64_t dal = da + (uint64_t)(kABytes >> 4), dbh = dal + (uint64_t)(kABytes >> 4);
          const uint64_t dbl = dbh + (uint64_t)(b_bytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if constexpr (kBx3) {
              umma_bf16(acc, dal + 2 * ks, dbh + 2 * ks, idesc, (it | ks) != 0 ? 1u : 0u);
              umma_bf16(acc, da + 2 * ks, dbl + 2 * ks, idesc, 1u);
              umma_bf16(acc, da + 2 * ks, dbh + 2 * ks, idesc, 1u);
            } else {
              umma_tf32(acc, dal + 2 * ks, dbh + 2 * ks, idesc, (it | ks) != 0 ? 1u : 0u);
              umma_tf32(acc, da + 2 * ks, dbl + 2 * ks, idesc, 1u);
              umma_tf32(acc, da + 2 * ks, dbh + 2 * ks, idesc, 1u);
            }
          }
          umma_commit(smem_u32(&empty_bar[s]));
          if (it == n_active - 1) umma_commit(smem_u32(&tmem_full[b]));
        }
        __syncwarp();
      }
      if (n_active == 0 && lane == 0) mbar_arrive(smem_u32(&tmem_full[b]));   // nothing to accumulate: epilogue writes bias only
      gbase += n_active;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&meta_empty[mb]));
    }
    tc_fence_before();
  } else {
    // ======================= epilogue warps: metadata of the next tile, then drain of the current one ==========
    const int ew = warp - kPersistProducerWarps - 1;   // 0..3: index within the epilogue group
    const int lg = warp & 3;                           // TMEM lane group this warp may read (hardware: warp id % 4)
    const int et = tid - (kPersistProducerWarps + 1) * 32;   // 0..127
    // Metadata is prepared TWO tiles ahead in two halves: `issue(j)` puts the tile's global loads (row order + up to 27
    // neighbour indices per row) in flight into registers, the accumulator drain of the current tile runs under their
    // latency, and `finish(j)` stores them and compacts the active-chunk list.  (ncu, profiles/r2q_gg32: with the whole
    // preparation done back to back before each drain, producers and the MMA warp spent 18 % of all warp samples
    // waiting for the next tile's metadata - the epilogue group was the critical path of the kernel.)
    constexpr int kRegIdx = kLin ? 1 : 27;
    int32_t pre_row = -1;
    int32_t pre_idx[kRegIdx];
    auto issue = [&](int j) {
      const int64_t pos = tile_row0(j) + et;
      const bool in = pos < p.n_out;
      pre_row = -1;
      if (in) pre_row = (p.order != nullptr) ? __ldg(&p.order[pos]) : (int32_t)pos;
#pragma unroll
      for (int u = 0; u < kRegIdx; ++u) {
        pre_idx[u] = -1;
        if (u < p.kvol && in) pre_idx[u] = (p.nbr != nullptr) ? __ldg(&p.nbr[(int64_t)u * p.n_out + pos]) : (int32_t)pos;
      }
    };
    auto finish = [&](int j) {
      const int b = j % kMB;
      mbar_wait(smem_u32(&meta_empty[b]), ((uint32_t)(j / kMB) & 1u) ^ 1u);
      int32_t* idx_s = meta_idx(b);
      int32_t* row_s = meta_row(b);
      uint16_t* active = meta_active(b);
      const int64_t row0 = tile_row0(j);
      row_s[et] = pre_row;
      if (et < 4) scratch[et] = 0u;   // kmask
#pragma unroll
      for (int u = 0; u < kRegIdx; ++u)
        if (u < p.kvol) idx_s[u * kTileM + et] = pre_idx[u];
      // kernels with more than 27 offsets (the 5^3 stem): the rest in batches of 9 loads
      const int64_t pos = row0 + et;
      const bool in = pos < p.n_out;
      for (int k0 = kRegIdx; k0 < p.kvol; k0 += 9) {
        int32_t v[9];
#pragma unroll
        for (int u = 0; u < 9; ++u) {
          const int k = k0 + u;
          v[u] = -1;
          if (k < p.kvol && in) v[u] = (p.nbr != nullptr) ? __ldg(&p.nbr[(int64_t)k * p.n_out + pos]) : (int32_t)pos;
        }
#pragma unroll
        for (int u = 0; u < 9; ++u)
          if (k0 + u < p.kvol) idx_s[(k0 + u) * kTileM + et] = v[u];
      }
      bar_sync_named(1, 128);
      // which offsets have a neighbour in this tile: warp ew looks at offsets ew, ew + 4, ...
      for (int k = ew; k < p.kvol; k += 4) {
        const int32_t* r = idx_s + k * kTileM + lane;
        const bool a = (r[0] >= 0) | (r[32] >= 0) | (r[64] >= 0) | (r[96] >= 0);
        if (__any_sync(0xffffffffu, a) && lane == 0) atomicOr(&scratch[k >> 5], 1u << (k & 31));
      }
      bar_sync_named(1, 128);
      const uint32_t km[4] = {scratch[0], scratch[1], scratch[2], scratch[3]};
      int n_total = 0;
      for (int c0 = 0; c0 < p.num_chunks; c0 += 128) {
        const int c = c0 + et;
        bool on = false;
        if (c < p.num_chunks) {
          const int k_lo = (int)(kc_s[c * 8] >> 16);
          const uint32_t last = kc_s[c * 8 + 7];
          const int k_hi = (last == 0xffffffffu) ? p.kvol - 1 : (int)(last >> 16);
          for (int k = k_lo; k <= k_hi; ++k) on |= ((km[k >> 5] >> (k & 31)) & 1u) != 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) scratch[4 + ew] = __popc(bal);
        bar_sync_named(1, 128);
        int before = 0, all = 0;
#pragma unroll
        for (int w2 = 0; w2 < 4; ++w2) { const int cnt = (int)scratch[4 + w2]; before += (w2 < ew) ? cnt : 0; all += cnt; }
        if (on) active[n_total + before + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)c;
        n_total += all;
        bar_sync_named(1, 128);
      }
      if (et == 0) {
        int base = 0, cnt = n_total;
        if (ks > 1) {
          const int per = (n_total + ks - 1) / ks;
          base = (vtile(j) % ks) * per;
          if (base > n_total) base = n_total;
          cnt = n_total - base;
          if (cnt > per) cnt = per;
        }
        meta_count(b)[0] = cnt;
        meta_count(b)[1] = base;
      }
      bar_sync_named(1, 128);   // every epilogue thread's metadata writes are done
      if (et == 0) mbar_arrive(smem_u32(&meta_full[b]));
    };
    constexpr int dist = kMB - 1;   // tiles of metadata lead
    if (my_tiles > 0) { issue(0); finish(0); }
    if constexpr (kMB == 3) {
      if (my_tiles > 1) { issue(1); finish(1); }
    }
    for (int i = 0; i < my_tiles; ++i) {
      const bool ahead = i + dist < my_tiles;
      if (ahead) issue(i + dist);
      if constexpr (kMB == 2) {
        if (ahead) finish(i + dist);   // one buffer of lead: the next tile's metadata must not wait for this drain
      }
      const int b = i & 1, mb = i % kMB;
      mbar_wait(smem_u32(&tmem_full[b]), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const int n_active = *meta_count(mb);
      const int32_t j32 = meta_row(mb)[lg * 32 + lane];
      const int64_t j = j32;
      const uint32_t acc = tmem_base + (uint32_t)b * acc_stride + ((uint32_t)(lg * 32) << 16);
      // one block of 16 accumulator columns of this thread's row: TMEM -> registers -> epilogue -> global
      auto drain_block = [&](int col0, const float* pa, const float* py) {
        uint32_t v[16];
        if (n_active > 0) {
          tmem_ld_x16(acc + (uint32_t)col0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0u;
        }
        if (j32 < 0) return;
        const bool first_slice = ks == 1 || (vtile(i) % ks) == 0;
        float f[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int co = col0 + q;
          f[q] = __uint_as_float(v[q]) + ((p.bias != nullptr && co < p.cout && first_slice) ? __ldg(&p.bias[co]) : 0.f);
        }
        if (ks > 1) {   // partial sums of this slice (act == 0 on this path): vector reductions into the pre-zeroed rows
          if (n_active == 0 && !(first_slice && p.bias != nullptr)) return;
          float* yr = reinterpret_cast<float*>(p.y) + j * p.y_row + col0;
          if ((col0 + 16 <= p.cout) && ((p.cout & 3) == 0)) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yr + 4 * q), "f"(f[4 * q]),
                           "f"(f[4 * q + 1]), "f"(f[4 * q + 2]), "f"(f[4 * q + 3]) : "memory");
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q)   // (static indices: a run-time loop here put f[] into local memory, +7 % on L0)
              if (col0 + q < p.cout) atomicAdd(yr + q, f[q]);
          }
          return;
        }
        fp32_row_epilogue<kLin>(p, f, j, col0, pa, py);
      };
      if (kLin && p.act >= 2 && j32 >= 0) {
        // backward epilogues of the render MLP read one or two more [rows, Cout] operands: fetched ONE column block ahead
        // as 128-bit loads into two alternating register sets, so that the (dependent, scalar) loads no longer serialise
        // the drain (r2za: 0.9 - 1.35 ms per such layer at 524 k rows = 0.7 TB/s)
        float a0[16], y0[16], a1[16], y1[16];
        epilogue_operands(p, j, 0, a0, y0);
        for (int col0 = 0; col0 < p.n_pad; col0 += 32) {
          const bool second = col0 + 16 < p.n_pad;
          if (second) epilogue_operands(p, j, col0 + 16, a1, y1);
          drain_block(col0, a0, y0);
          if (col0 + 32 < p.n_pad) epilogue_operands(p, j, col0 + 32, a0, y0);
          if (second) drain_block(col0 + 16, a1, y1);
        }
      } else {
        for (int col0 = 0; col0 < p.n_pad; col0 += 16) drain_block(col0, nullptr, nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[b]));   // accumulator b may be overwritten
      if constexpr (kMB == 3) {
        if (ahead) finish(i + dist);
      }
    }
  }
  __syncthreads();
  if (warp == kPersistProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// returns PV2_EUNSUPPORTED when the shape does not fit the persistent kernel (the caller uses the one-tile kernel)
// ===============================================================================================================
// bf16 storage, TMA on both operands (persistent).  The gather itself is done by the copy engine:
//   * A (activations): `cp.async.bulk.tensor.2d ... tile::gather4` -- each lane of ONE producer warp hands four row indices
//     of the tile's neighbour list to the TMA unit, which fetches the four 128-byte channel runs [row][ci0 .. ci0 + 63]
//     straight into the 128B-swizzled K-major stage (UTMALDG in SASS).  A missing neighbour is row index -1: out of
//     bounds for the tensor map, which the hardware zero-fills.  32 lanes x 4 rows = the 128-row tile, one warp
//     instruction per chunk; no thread touches the data, no registers, no proxy fences;
//   * B (weights): one 2-D box [n_pad rows][64 channels] of the bf16 weight matrix [Cout][K * Cin] per chunk;
//   * the stage's mbarrier counts transaction bytes (16 KB + n_pad * 128), the MMA warp waits on it.
// Roles: warp 0 producer, warp 1 MMA issuer (kind::f16, M = 128, N = Cout), warps 2-5 metadata + epilogue exactly as in
// the fp32 persistent kernel (double-buffered tile metadata and TMEM accumulators).  Requires Cin % 64 == 0 (a chunk never
// straddles two kernel offsets) and contiguous [Cout][K][Cin] weights; other layers use the cp.async kernel above.
constexpr int kTmaThreads = 6 * 32;

__global__ void __launch_bounds__(kTmaThreads) umma_gather_gemm_tma_kernel(const GGParams p, int num_tiles,
                                                                           const __grid_constant__ CUtensorMap xmap,
                                                                           const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b_bytes = p.n_pad * 128;
  const int stage_bytes = kABytes + b_bytes;
  const int cpk = p.cin >> 6;                      // 64-channel chunks per kernel offset
  const PersistLayout L = persist_layout(p.kvol, p.num_chunks);
  uint8_t* stage_base = smem;
  uint8_t* meta_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta_base + 2 * L.meta_bytes);
  uint64_t* full_bar = bars;                        // [kMaxStages]  1 arrival (+ transaction bytes)
  uint64_t* empty_bar = bars + kMaxStages;          // [kMaxStages]  tcgen05.commit
  uint64_t* meta_full = bars + 2 * kMaxStages;      // [2]
  uint64_t* meta_empty = meta_full + 2;             // [2]  producer warp + MMA warp
  uint64_t* tmem_full = meta_empty + 2;             // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]  4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* scratch = tmem_slot + 1;

  auto meta_idx = [&](int b) { return reinterpret_cast<int32_t*>(meta_base + (size_t)b * L.meta_bytes); };
  auto meta_row = [&](int b) { return meta_idx(b) + p.kvol * kTileM; };
  auto meta_active = [&](int b) { return reinterpret_cast<uint16_t*>(meta_row(b) + kTileM); };
  auto meta_count = [&](int b) { return reinterpret_cast<int*>(meta_base + (size_t)b * L.meta_bytes + L.meta_bytes - 16); };

  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_row0 = [&](int i) { return (int64_t)(num_tiles - 1 - ((int)blockIdx.x + i * (int)gridDim.x)) * kTileM; };

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&meta_full[b]), 1);
      mbar_init(smem_u32(&meta_empty[b]), 2);
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&wmap);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_stride = p.tmem_cols >> 1;

  if (warp == 0) {
    // ======================= producer: one warp drives the copy engine =======================
    int gbase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      mbar_wait(smem_u32(&meta_full[b]), (uint32_t)(i >> 1) & 1u);
      const int n_active = *meta_count(b);
      const int32_t* idx_s = meta_idx(b);
      const uint16_t* active = meta_active(b);
      for (int it = 0; it < n_active; ++it) {
        const int g = gbase + it;
        const int s = g % p.stages;
        const uint32_t ph = (uint32_t)(g / p.stages) & 1u;
        const int c = (int)active[it];
        const int k = c / cpk;
        const int ci0 = (c - k * cpk) << 6;
        const int4 r4 = *reinterpret_cast<const int4*>(idx_s + k * kTileM + 4 * lane);   // rows 4 lane .. 4 lane + 3
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t bar = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_u32(stage_base + (size_t)s * stage_bytes);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar, (uint32_t)(kABytes + b_bytes));
          tma_load_2d(a_dst + kABytes, &wmap, bar, c << 6, p.w2_row0);
        }
        __syncwarp();
        tma_gather4_2d(a_dst + (uint32_t)lane * 512u, &xmap, bar, ci0, r4.x, r4.y, r4.z, r4.w);
      }
      gbase += n_active;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&meta_empty[b]));
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc = make_idesc(1 /*BF16*/, kTileM, p.n_pad);
    int gbase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      mbar_wait(smem_u32(&meta_full[b]), (uint32_t)(i >> 1) & 1u);
      const int n_active = *meta_count(b);
      mbar_wait(smem_u32(&tmem_empty[b]), ((uint32_t)(i >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)b * acc_stride;
      for (int it = 0; it < n_active; ++it) {
        const int g = gbase + it;
        const int s = g % p.stages;
        const uint32_t ph = (uint32_t)(g / p.stages) & 1u;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = smem_desc_kmajor_sw128(smem_u32(stage_base + (size_t)s * stage_bytes));
          const uint64_t db = da + (uint64_t)(kABytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(acc, da + 2 * ks, db + 2 * ks, idesc, (it | ks) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[s]));
          if (it == n_active - 1) umma_commit(smem_u32(&tmem_full[b]));
        }
        __syncwarp();
      }
      if (n_active == 0 && lane == 0) mbar_arrive(smem_u32(&tmem_full[b]));
      gbase += n_active;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&meta_empty[b]));
    }
    tc_fence_before();
  } else {
    // ======================= epilogue warps: metadata of the next tile, then drain of the current one ==========
    const int ew = warp - 2;                           // 0..3
    const int lg = warp & 3;                           // TMEM lane group this warp may read (hardware: warp id % 4)
    const int et = tid - 64;                           // 0..127
    auto prepare = [&](int j) {
      const int b = j & 1;
      mbar_wait(smem_u32(&meta_empty[b]), ((uint32_t)(j >> 1) & 1u) ^ 1u);
      int32_t* idx_s = meta_idx(b);
      int32_t* row_s = meta_row(b);
      uint16_t* active = meta_active(b);
      const int64_t row0 = tile_row0(j);
      const int64_t pos = row0 + et;
      const bool in = pos < p.n_out;
      {
        int32_t jr = -1;
        if (in) jr = (p.order != nullptr) ? __ldg(&p.order[pos]) : (int32_t)pos;
        row_s[et] = jr;
      }
      if (et < 4) scratch[et] = 0u;
      for (int k0 = 0; k0 < p.kvol; k0 += 9) {
        int32_t v[9];
#pragma unroll
        for (int u = 0; u < 9; ++u) {
          const int k = k0 + u;
          v[u] = -1;
          if (k < p.kvol && in) v[u] = (p.nbr != nullptr) ? __ldg(&p.nbr[(int64_t)k * p.n_out + pos]) : (int32_t)pos;
        }
#pragma unroll
        for (int u = 0; u < 9; ++u)
          if (k0 + u < p.kvol) idx_s[(k0 + u) * kTileM + et] = v[u];
      }
      bar_sync_named(1, 128);
      for (int k = ew; k < p.kvol; k += 4) {
        const int32_t* r = idx_s + k * kTileM + lane;
        const bool a = (r[0] >= 0) | (r[32] >= 0) | (r[64] >= 0) | (r[96] >= 0);
        if (__any_sync(0xffffffffu, a) && lane == 0) atomicOr(&scratch[k >> 5], 1u << (k & 31));
      }
      bar_sync_named(1, 128);
      const uint32_t km[4] = {scratch[0], scratch[1], scratch[2], scratch[3]};
      int n_total = 0;
      for (int c0 = 0; c0 < p.num_chunks; c0 += 128) {
        const int c = c0 + et;
        bool on = false;
        if (c < p.num_chunks) { const int k = c / cpk; on = ((km[k >> 5] >> (k & 31)) & 1u) != 0; }
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) scratch[4 + ew] = __popc(bal);
        bar_sync_named(1, 128);
        int before = 0, all = 0;
#pragma unroll
        for (int w2 = 0; w2 < 4; ++w2) { const int cnt = (int)scratch[4 + w2]; before += (w2 < ew) ? cnt : 0; all += cnt; }
        if (on) active[n_total + before + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)c;
        n_total += all;
        bar_sync_named(1, 128);
      }
      if (et == 0) *meta_count(b) = n_total;
      bar_sync_named(1, 128);
      if (et == 0) mbar_arrive(smem_u32(&meta_full[b]));
    };
    if (my_tiles > 0) prepare(0);
    for (int i = 0; i < my_tiles; ++i) {
      if (i + 1 < my_tiles) prepare(i + 1);
      const int b = i & 1;
      mbar_wait(smem_u32(&tmem_full[b]), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const int n_active = *meta_count(b);
      const int32_t j32 = meta_row(b)[lg * 32 + lane];
      const int64_t j = j32;
      const uint32_t acc = tmem_base + (uint32_t)b * acc_stride + ((uint32_t)(lg * 32) << 16);
      for (int col0 = 0; col0 < p.n_pad; col0 += 16) {
        uint32_t v[16];
        if (n_active > 0) {
          tmem_ld_x16(acc + (uint32_t)col0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0u;
        }
        if (j32 < 0) continue;
        float f[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int co = col0 + q;
          f[q] = __uint_as_float(v[q]) + ((p.bias != nullptr && co < p.cout) ? __ldg(&p.bias[co]) : 0.f);
        }
        __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(p.y) + j * p.y_row + col0;
        if (col0 + 16 <= p.cout && (p.cout & 7) == 0) {
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
            pk[q] = *reinterpret_cast<uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(yr) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(yr + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
          for (int q = 0; q < 16 && col0 + q < p.cout; ++q) yr[q] = __float2bfloat16(f[q]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[b]));
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember it per (kernel, device)
template <typename K>
cudaError_t ensure_smem_optin(K kernel, bool (&done)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

// ---- bf16x3 weights: [2][rows][cols] bf16 (hi half, then lo half), rows = Cout rounded up to 16, cols = K * Cin rounded up
// to 64, zero padded.  One small kernel per convolution call (<= 1.8 M weights) and one TMA descriptor over the result.
__global__ void presplit_weights_bf16_kernel(const float* __restrict__ w, int64_t w_sco, int64_t w_sk, int cout, int kvol,
                                             int cin, __nv_bfloat16* __restrict__ out, int rows, int cols) {
  const int64_t total = (int64_t)rows * cols;
  const int ktot = kvol * cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    float v = 0.f;
    if (r < cout && c < ktot) {
      const int k = c / cin;
      v = __ldg(&w[(int64_t)r * w_sco + (int64_t)k * w_sk + (c - k * cin)]);
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = h;
    out[total + i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}
// 2-D bf16 matrix [rows][cols] (row-major, cols % 8 == 0), boxes of box_rows x 64 columns (128 B), 128-byte swizzle
int encode_bf16_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (enc == nullptr) return PV2_EUNSUPPORTED;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : PV2_EUNSUPPORTED;
}

// returns PV2_EUNSUPPORTED when the shape does not fit the persistent kernel (the caller uses the one-tile kernel).
// w2 != nullptr selects the bf16x3 mode: the pre-split weights [2][w2_rows][w2_cols] (p.w2_row0 = this slice's first row).
int launch_persistent(const GGParams& p0, cudaStream_t stream, const void* w2 = nullptr, int w2_cols = 0) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("PV2_GG_PERSISTENT"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return PV2_EUNSUPPORTED;
  const bool bx3 = w2 != nullptr;
  GGParams p = p0;
  if (p.x_lo_off != 0) return PV2_EUNSUPPORTED;
  p.n_pad = (p.cout + 15) / 16 * 16;
  p.num_chunks = bx3 ? (p.kvol * p.cin + 63) / 64 : (p.kvol * p.cin + 31) / 32;
  const int tiles = (int)((p.n_out + kTileM - 1) / kTileM);
  if (p.n_pad > 256) return PV2_EUNSUPPORTED;
  if (!bx3) {
    if (tiles < 2 * PV2_SM_COUNT) return PV2_EUNSUPPORTED;   // few tiles: the split-K kernel
    // narrow layers fit two one-tile CTAs per SM, which overlaps tiles just as well and measured faster (63 vs 73 us at
    // 100 k voxels, 32 -> 32); the persistent kernel is for the layers that shared memory limits to one CTA per SM
    const int sb = (kABytes + p.n_pad * 128) * 2;
    const int fx = p.kvol * kTileM * 4 + kTileM * 4 + (p.num_chunks * 2 + 8) + (2 * kMaxStages + 1) * 8 + 128 + p.num_chunks * 32 + 1024 + 64;
    if ((112 * 1024 - fx) / sb >= 2) return PV2_EUNSUPPORTED;
  }
  if (p.x_row >= (int64_t)1 << 32 || p.w_sco >= (int64_t)1 << 32 || p.w_sk >= (int64_t)1 << 32 || p.cin > 65535)
    return PV2_EUNSUPPORTED;
  p.x_row32 = (uint32_t)p.x_row; p.w_sco32 = (uint32_t)p.w_sco; p.w_sk32 = (uint32_t)p.w_sk;
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < 2 * p.n_pad) p.tmem_cols <<= 1;     // two accumulators
  if (p.tmem_cols > 512) return PV2_EUNSUPPORTED;
  const int stage_bytes = (kABytes + p.n_pad * 128) * 2;
  PersistLayout L = persist_layout(p.kvol, p.num_chunks, kMetaBufs);
  int stages = (227 * 1024 - L.fixed) / stage_bytes;   // 227 KB: the opt-in maximum of dynamic shared memory per CTA
  p.meta_bufs = kMetaBufs;
  static int force_bufs = -1;
  if (force_bufs < 0) { const char* e = getenv("PV2_GG_META_BUFS"); force_bufs = e ? atoi(e) : 0; }   // development switch
  // two buffers only where three leave fewer than two stages (Cout = 256, the 125-offset stem); measured in one run
  // (profiles/r2y_micro_levels_mb{0,3}.txt): 96 -> 128 data gradient 146 us with 3 buffers / 2 stages, 158 us with 2 / 3
  if (stages < kPersistGroups || force_bufs == 2) {
    L = persist_layout(p.kvol, p.num_chunks, 2);
    stages = (227 * 1024 - L.fixed) / stage_bytes;
    p.meta_bufs = 2;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < kPersistGroups) return PV2_EUNSUPPORTED;          // groups <= stages (mbarrier parity aliasing)
  int groups = pv2_get_option("gg_groups");
  if (groups < 2 || groups > 4) groups = 2;   // measured (profiles/r2p_micro_levels_g{2,3,4}.txt): no difference
  if (!bx3 || stages < groups + 1) groups = 2;
  p.stages = stages;
  // split-K for the levels with fewer row tiles than SMs: ~one virtual tile per SM, at least 2 chunks per slice
  int ksplit = 1;
  if (bx3 && p0.ksplit == -1 && tiles * 2 <= PV2_SM_COUNT && p.act == 0 && !p.y_split && p.num_chunks >= 4) {
    ksplit = (PV2_SM_COUNT + tiles - 1) / tiles;
    if (ksplit > p.num_chunks / 2) ksplit = p.num_chunks / 2;
    const int cap = pv2_get_option("gg_ksplit_max");
    if (cap > 0 && ksplit > cap) ksplit = cap;
    if (ksplit < 1) ksplit = 1;
  }
  p.ksplit = ksplit;
  p.ablate = 0;
  const size_t smem = (size_t)stages * stage_bytes + L.fixed;
  const int vtiles = tiles * ksplit;
  const int grid = vtiles < PV2_SM_COUNT ? vtiles : PV2_SM_COUNT;
  int launches = 1;
  if (ksplit > 1) {
    cudaMemset2DAsync(p.y, (size_t)p.y_row * sizeof(float), 0, (size_t)p.cout * sizeof(float), (size_t)p.n_out, stream);
    ++launches;
  }
  CUtensorMap wmap;
  memset(&wmap, 0, sizeof(wmap));
  (void)groups;
  cudaError_t e = cudaSuccess;
  const bool linear = p.nbr == nullptr && p.kvol == 1;   // pv2_linear: identity map, activation epilogues
  if (bx3) {
    const int rc = encode_bf16_map(&wmap, w2, 2 * (int64_t)p.w2_rows, w2_cols, p.n_pad);
    if (rc != 0) return rc;
    static bool done3[64] = {}, done2[64] = {};
#define PV2_LAUNCH_PERSIST(BX, NB, LIN, FLAGS)                                                                       \
  do {                                                                                                              \
    e = ensure_smem_optin(umma_gather_gemm_persistent_kernel<BX, 2, NB, LIN>, FLAGS);                               \
    if (e != cudaSuccess) return (int)e;                                                                            \
    umma_gather_gemm_persistent_kernel<BX, 2, NB, LIN><<<grid, kPersistThreads, smem, stream>>>(p, vtiles, wmap);   \
  } while (0)
    static bool d3l[64] = {}, d2l[64] = {};
    if (linear) { if (p.meta_bufs == 3) PV2_LAUNCH_PERSIST(true, 3, true, d3l); else PV2_LAUNCH_PERSIST(true, 2, true, d2l); }
    else { if (p.meta_bufs == 3) PV2_LAUNCH_PERSIST(true, 3, false, done3); else PV2_LAUNCH_PERSIST(true, 2, false, done2); }
  } else {
    static bool done3[64] = {}, done2[64] = {};
    static bool d3l[64] = {}, d2l[64] = {};
    if (linear) { if (p.meta_bufs == 3) PV2_LAUNCH_PERSIST(false, 3, true, d3l); else PV2_LAUNCH_PERSIST(false, 2, true, d2l); }
    else { if (p.meta_bufs == 3) PV2_LAUNCH_PERSIST(false, 3, false, done3); else PV2_LAUNCH_PERSIST(false, 2, false, done2); }
#undef PV2_LAUNCH_PERSIST
  }
  PV2_DONE(launches);
}


// bf16 storage through TMA gather4.  PV2_EUNSUPPORTED -> the caller uses the cp.async kernel.
int launch_tma_bf16(const GGParams& p0, cudaStream_t stream, int64_t n_in, const void* w_full, int cout_full) {
  // Measured on B200 (profiles/r2e_micro_bf16_*.txt): the copy engine takes ~40-70 cycles per gather4 (four 128-byte
  // rows), i.e. 13-18 B/clk/SM -- faster than the cp.async producers only for long, narrow launches (1 M voxels x 64
  // channels: 349 vs 451 us), slower for 128 channels or ~100 k voxels.  "auto" follows that crossover.
  const int mode = pv2_get_option("gg_tma");
  if (mode == 0 || tensor_map_encoder() == nullptr) return PV2_EUNSUPPORTED;
  if (mode < 0 && !(p0.cin == 64 && p0.n_out >= 262144)) return PV2_EUNSUPPORTED;
  GGParams p = p0;
  if ((p.cin & 63) != 0 || p.nbr == nullptr || p.x_row != p.cin || p.w_sk != p.cin || p.w_sco != (int64_t)p.kvol * p.cin)
    return PV2_EUNSUPPORTED;
  p.n_pad = (p.cout + 15) / 16 * 16;
  if (p.n_pad > 256 || (int64_t)p.kvol * p.cin >= (1 << 30) || n_in >= (int64_t)1 << 31) return PV2_EUNSUPPORTED;
  p.num_chunks = p.kvol * (p.cin >> 6);
  const int tiles = (int)((p.n_out + kTileM - 1) / kTileM);
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < 2 * p.n_pad) p.tmem_cols <<= 1;
  if (p.tmem_cols > 512) return PV2_EUNSUPPORTED;
  const int stage_bytes = kABytes + p.n_pad * 128;
  const PersistLayout L = persist_layout(p.kvol, p.num_chunks);
  // two CTAs per SM when at least three stages fit in half of the shared memory
  int budget = 227 * 1024;
  if ((113 * 1024 - L.fixed) / stage_bytes >= 3) budget = 113 * 1024;
  int stages = (budget - L.fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return PV2_EUNSUPPORTED;
  p.stages = stages;
  p.ksplit = 1; p.ablate = 0;
  const size_t smem = (size_t)stages * stage_bytes + L.fixed;
  CUtensorMap xmap, wmap;
  EncodeTiledFn enc = tensor_map_encoder();
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.cin, (cuuint64_t)n_in};
    const cuuint64_t strides[1] = {(cuuint64_t)p.cin * 2};
    const cuuint32_t box[2] = {64u, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    if (enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p.x), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return PV2_EUNSUPPORTED;
  }
  if (encode_bf16_map(&wmap, w_full, cout_full, (int64_t)p.kvol * p.cin, p.n_pad) != 0) return PV2_EUNSUPPORTED;
  static bool done[64] = {};
  cudaError_t e = ensure_smem_optin(umma_gather_gemm_tma_kernel, done);
  if (e != cudaSuccess) return (int)e;
  const int ctas = (budget < 200 * 1024) ? 2 : 1;
  const int grid = tiles < ctas * PV2_SM_COUNT ? tiles : ctas * PV2_SM_COUNT;
  umma_gather_gemm_tma_kernel<<<grid, kTmaThreads, smem, stream>>>(p, tiles, xmap, wmap);
  PV2_DONE(1);
}

template <bool kSplit, bool kPre, int kGroups>
int launch(const GGParams& p0, cudaStream_t stream) {
  using T = ModeTraits<kSplit>;
  GGParams p = p0;
  p.n_pad = (p.cout + 15) / 16 * 16;
  p.num_chunks = (p.kvol * p.cin + T::kEPR - 1) / T::kEPR;
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < p.n_pad) p.tmem_cols <<= 1;
  const int stage_bytes = (kABytes + p.n_pad * 128) * T::kOperands;
  if (p.x_row >= (int64_t)1 << 32 || p.w_sco >= (int64_t)1 << 32 || p.w_sk >= (int64_t)1 << 32 || p.cin > 65535)
    return PV2_EUNSUPPORTED;
  p.x_row32 = (uint32_t)p.x_row; p.w_sco32 = (uint32_t)p.w_sco; p.w_sk32 = (uint32_t)p.w_sk;
  const int fixed = p.kvol * kTileM * 4 + kTileM * 4 + (p.num_chunks * 2 + 8) + (2 * kMaxStages + 1) * 8 + 128 +
                    p.num_chunks * 32 + 1024 + 64;
  int stages = (220 * 1024 - fixed) / stage_bytes;
  // two resident CTAs per SM when at least two stages fit in half of the shared memory: the second CTA's setup,
  // gathers and epilogue overlap the first one's main loop (the kernel is latency-bound, not bandwidth-bound)
  static int want_ctas = -1;
  if (want_ctas < 0) { const char* e = getenv("PV2_GG_CTAS"); want_ctas = e ? atoi(e) : 2; }
  if (want_ctas >= 2 && (112 * 1024 - fixed) / stage_bytes >= 2) stages = (112 * 1024 - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return PV2_EUNSUPPORTED;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  {
    static bool done[64] = {};
    cudaError_t e = ensure_smem_optin(umma_gather_gemm_kernel<kSplit, kPre, kGroups>, done);
    if (e != cudaSuccess) return (int)e;
  }
  const unsigned tiles = (unsigned)((p.n_out + kTileM - 1) / kTileM);
  int ksplit = 1;
  if (kSplit && !p.y_split && p.act == 0 && tiles * 2 <= PV2_SM_COUNT && p.num_chunks >= 16) {  // (act: one writer per row)
    ksplit = (PV2_SM_COUNT + (int)tiles - 1) / (int)tiles;  // ~one CTA per SM in total
    if (ksplit > p.num_chunks / 8) ksplit = p.num_chunks / 8;
    // measured (profiles/r2g_ksplit.txt): with >= 8 row tiles 8 slices beat 12 (L3: 55 vs 61 us; fewer partial tiles to
    // reduce through global atomics), with 3 tiles (L4) more slices win (31 vs 55 us)
    if (tiles >= 8 && ksplit > 8) ksplit = 8;
    const int cap = pv2_get_option("gg_ksplit_max");
    if (cap > 0 && ksplit > cap) ksplit = cap;
    if (ksplit < 1) ksplit = 1;
  }
  p.ksplit = ksplit;
  {
    static int abl = -1;
    if (abl < 0) { const char* e = getenv("PV2_GG_ABLATE"); abl = e ? atoi(e) : 0; }
    p.ablate = abl;
  }
  int launches = 1;
  if (ksplit > 1) {
    // only this launch's columns: a caller may be filling other column slices of the same rows (Cout > 256)
    cudaMemset2DAsync(p.y, (size_t)p.y_row * sizeof(float), 0, (size_t)p.cout * sizeof(float), (size_t)p.n_out, stream);
    ++launches;
  }
  dim3 grid(tiles, (unsigned)ksplit);
  umma_gather_gemm_kernel<kSplit, kPre, kGroups><<<grid, kGroups * 128 + 32, smem, stream>>>(p);
  PV2_DONE(launches);
}

}  // namespace

// Producer groups of the fp32 kernel.  Narrow layers (two stages fit in half of the shared memory) run two resident
// CTAs of 2 groups each; wide layers are limited to one CTA per SM by shared memory and get a third group instead.
// PV2_GG_GROUPS = 2 forces two groups everywhere (development switch).
static int fp32_groups(int cout, int kvol, int cin) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("PV2_GG_GROUPS"); forced = e ? atoi(e) : 0; }
  if (forced == 2) return 2;
  const int n_pad = (cout + 15) / 16 * 16;
  const int num_chunks = (kvol * cin + 31) / 32;
  const int stage_bytes = (kABytes + n_pad * 128) * 2;
  const int fixed = kvol * kTileM * 4 + kTileM * 4 + (num_chunks * 2 + 8) + (2 * kMaxStages + 1) * 8 + 128 + num_chunks * 32 + 1024 + 64;
  if ((112 * 1024 - fixed) / stage_bytes >= 2) return 2;
  // A group may only wait one phase ahead on a stage's mbarrier (parity waits alias every second phase): with more
  // groups than stages a fast group would see the completion it needs as already past.  Hence groups <= stages.
  return (220 * 1024 - fixed) / stage_bytes >= 3 ? 3 : 2;
}

#ifdef PV2_MBAR_DEBUG
extern "C" int pv2_debug_dump_waits(unsigned* out, int cap) {   // development only; not part of the ABI header
  unsigned n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, pv2::g_wait_n, sizeof(n));
  if (n > 2048) n = 2048;
  if ((int)n > cap) n = cap;
  cudaMemcpyFromSymbol(out, pv2::g_wait_log, (size_t)n * 16);
  return (int)n;
}
#endif

extern "C" {

static int bx3_enabled() { return pv2_get_option("gg_bx3") != 0; }

// fp32: room for the weights pre-split into bf16 hi / lo matrices (the bf16x3 kernel's TMA operand); bf16: none.
size_t pv2_spconv_workspace_bytes(int64_t n_in, int cin, int cout, int kvol, int dtype) {
  (void)n_in;
  if (dtype != PV2_F32 || cin <= 0 || cout <= 0 || kvol <= 0) return 0;
  const size_t rows = ((size_t)cout + 15) / 16 * 16, cols = ((size_t)kvol * cin + 63) / 64 * 64;
  return (2 * rows * cols * 2 + 255) / 256 * 256;
}

// returns PV2_EUNSUPPORTED when the shape does not fit the tensor-core kernel (caller falls back to the SIMT kernel)
int pv2_spconv_gather_gemm_umma(const void* x, const void* w, int64_t w_sco, int64_t w_sk, const float* bias,
                                const int32_t* nbr, const int32_t* order, void* y, int64_t n_in, int64_t n_out, int cin,
                                int cout, int kvol, int dtype, void* workspace, size_t workspace_bytes, void* stream_) {
  PV2_CHECK_ARG(n_in >= 0 && n_out >= 0 && cin > 0 && cout > 0 && kvol > 0);
  if (n_out == 0) return 0;
  PV2_CHECK_ARG(x && w && nbr && y);
  const int epp = (dtype == PV2_BF16) ? 8 : 4;
  if (kvol > 128 || (cin % epp) != 0 || (w_sco % epp) != 0 || (w_sk % epp) != 0) return PV2_EUNSUPPORTED;
  if (cout > 256 && (cout % 16) != 0) return PV2_EUNSUPPORTED;
  if (((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)y & 15)) return PV2_EUNSUPPORTED;
  if (dtype != PV2_BF16 && dtype != PV2_F32) return PV2_EUNSUPPORTED;
  GGParams p{};
  p.x = x; p.w = w; p.w_sco = w_sco; p.w_sk = w_sk; p.bias = bias; p.nbr = nbr; p.order = order; p.y = y;
  p.n_out = n_out; p.cin = cin; p.cout = cout; p.kvol = kvol;
  p.x_row = cin; p.x_lo_off = 0;
  p.y_row = cout; p.y_lo_off = 0; p.y_split = 0; p.act = 0; p.y2 = nullptr; p.y2_row = 0; p.y2_lo_off = 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  // tcgen05.mma takes N <= 256: wider outputs (the 384-channel data gradient of dec3, spconv_unet_v1m1_base.py:171-177)
  // run as balanced column slices, each a multiple of 16 wide
  const int nsl = (cout + 255) / 256;
  const int per = ((cout + nsl - 1) / nsl + 15) / 16 * 16;
  const int eb = (dtype == PV2_BF16) ? 2 : 4;
  // bf16x3 persistent kernel: fp32 layers with whole 8-channel pieces and enough row tiles to fill the SMs (the small
  // deep levels keep the split-K 3xTF32 kernel); needs the pre-split weight workspace
  const int64_t tiles = (n_out + kTileM - 1) / kTileM;
  const int w2_rows = (cout + 15) / 16 * 16, w2_cols = (kvol * cin + 63) / 64 * 64;
  const bool small_ok = pv2_get_option("gg_bx3_split") > 0;   // split-K persistent kernel on the deep levels (off by default)
  bool bx3 = dtype == PV2_F32 && bx3_enabled() && (cin % 8) == 0 && (tiles * 2 > PV2_SM_COUNT || small_ok) && workspace != nullptr &&
             workspace_bytes >= pv2_spconv_workspace_bytes(n_in, cin, cout, kvol, dtype) && ((uintptr_t)workspace & 127) == 0 &&
             tensor_map_encoder() != nullptr;
  if (bx3) {
    const int64_t total = (int64_t)w2_rows * w2_cols;
    presplit_weights_bf16_kernel<<<pv2_grid_for(total, 256), 256, 0, stream>>>(
        (const float*)w, w_sco, w_sk, cout, kvol, cin, (__nv_bfloat16*)workspace, w2_rows, w2_cols);
    pv2_note_launches(1);
  }
  for (int co0 = 0; co0 < cout; co0 += per) {
    GGParams q = p;
    q.cout = (cout - co0 < per) ? cout - co0 : per;
    q.w = (const char*)w + (size_t)co0 * w_sco * eb;
    q.bias = bias ? bias + co0 : nullptr;
    q.y = (char*)y + (size_t)co0 * eb;
    if (bx3) {
      q.w2_row0 = co0; q.w2_rows = w2_rows;
      q.ksplit = -1;   // auto: split the contraction when the row tiles do not fill the SMs
      const int rp = launch_persistent(q, stream, workspace, w2_cols);
      q.ksplit = 0;
      if (rp == 0) continue;
      if (rp != PV2_EUNSUPPORTED) return rp;
    }
    if (dtype == PV2_F32) {
      const int rp = launch_persistent(q, stream);
      if (rp == 0) continue;
      if (rp != PV2_EUNSUPPORTED) return rp;
    }
    if (dtype == PV2_BF16 && tiles * 2 > PV2_SM_COUNT) {
      q.w2_row0 = co0;
      const int rt = launch_tma_bf16(q, stream, n_in, w, cout);
      if (rt == 0) continue;
      if (rt != PV2_EUNSUPPORTED) return rt;
    }
    const int rc = dtype == PV2_BF16 ? launch<false, false, 2>(q, stream)
                                     : (fp32_groups(q.cout, kvol, cin) == 3 ? launch<true, false, 3>(q, stream)
                                                                            : launch<true, false, 2>(q, stream));
    if (rc != 0) return rc;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Dense per-row linear layer on the same tensor-core kernel (identity row map): the render MLP (decoders.py:6-109).
//   y[j, 0:cout] = act( x[j, 0:cin] . w[0:cout, 0:cin]^T + bias )          fp32 storage, 3xTF32 arithmetic
// x is plain fp32 (x_presplit = 0) or split-precision (hi at x[j*x_row + c], lo at +x_lo_off; the halves are summed on
// load).  y / y2 can be written plain or split.
// room for the weights pre-split into bf16 hi / lo halves (bf16x3 arithmetic, weights by TMA); 0 for shapes that stay 3xTF32
size_t pv2_linear_workspace_bytes(int64_t rows, int cin, int cout, int x_presplit) {
  (void)rows;
  if (x_presplit || cin <= 0 || cout <= 0 || (cin % 8) != 0) return 0;
  const size_t r = ((size_t)cout + 15) / 16 * 16, c = ((size_t)cin + 63) / 64 * 64;
  return (2 * r * c * 2 + 255) / 256 * 256;
}

int pv2_linear(const float* x, int64_t x_row, int64_t x_lo_off, int x_presplit, const float* w, const float* bias,
               float* y, int64_t y_row, int64_t y_lo_off, int y_split, int act, float* y2, int64_t y2_row,
               int64_t y2_lo_off, int64_t rows, int cin, int cout, void* workspace, size_t workspace_bytes,
               void* stream_) {
  PV2_CHECK_ARG(rows >= 0 && cin > 0 && cout > 0 && cout <= 256 && (cin % 4) == 0 && act >= 0 && act <= 4);
  PV2_CHECK_ARG((act != 2 && act != 3) || y2 != nullptr);
  PV2_CHECK_ARG(act < 2 || !y_split);
  if (rows == 0) return 0;
  PV2_CHECK_ARG(x && w && y);
  PV2_CHECK_ARG((x_row % 4) == 0 && (x_lo_off % 4) == 0 && (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0);
  PV2_CHECK_ARG(!x_presplit || x_lo_off > 0);
  PV2_CHECK_ARG(x_presplit || x_row >= cin);
  GGParams p{};
  p.x = x; p.x_row = x_row; p.x_lo_off = x_presplit ? x_lo_off : 0;
  p.w = w; p.w_sco = cin; p.w_sk = cin;
  p.bias = bias; p.nbr = nullptr; p.order = nullptr; p.y = y; p.n_out = rows; p.cin = cin; p.cout = cout; p.kvol = 1;
  p.y_row = y_row; p.y_lo_off = y_lo_off; p.y_split = y_split; p.act = act;
  p.y2 = y2; p.y2_row = y2_row; p.y2_lo_off = y2_lo_off;
  if (x_presplit) return launch<true, true, 2>(p, (cudaStream_t)stream_);
  {
    // bf16x3 persistent kernel (same arithmetic as the sparse convolutions) when the rows fill the SMs and the caller
    // brought the weight workspace; "linear_bx3" = 0 keeps the 3xTF32 kernel (A/B switch)
    const size_t need = pv2_linear_workspace_bytes(rows, cin, cout, 0);
    if (pv2_get_option("linear_bx3") != 0 && need > 0 && workspace != nullptr && workspace_bytes >= need &&
        ((uintptr_t)workspace & 127) == 0 && rows >= 128 * PV2_SM_COUNT && tensor_map_encoder() != nullptr) {
      cudaStream_t stream = (cudaStream_t)stream_;
      const int w2_rows = (cout + 15) / 16 * 16, w2_cols = (cin + 63) / 64 * 64;
      presplit_weights_bf16_kernel<<<pv2_grid_for((int64_t)w2_rows * w2_cols, 256), 256, 0, stream>>>(
          w, cin, cin, cout, 1, cin, (__nv_bfloat16*)workspace, w2_rows, w2_cols);
      pv2_note_launches(1);
      GGParams q = p;
      q.w2_row0 = 0; q.w2_rows = w2_rows;
      const int rb = launch_persistent(q, stream, workspace, w2_cols);
      if (rb != PV2_EUNSUPPORTED) return rb;
    }
    const int rp = launch_persistent(p, (cudaStream_t)stream_);
    if (rp != PV2_EUNSUPPORTED) return rp;
  }
  return fp32_groups(cout, 1, cin) == 3 ? launch<true, false, 3>(p, (cudaStream_t)stream_)
                                        : launch<true, false, 2>(p, (cudaStream_t)stream_);
}

}  // extern "C"
