// Weight gradient of the sparse convolutions on the Blackwell tensor cores, second generation: MN-major bf16 operands.
//
//   dW[co, k, ci] += sum_j dY[j, co] * X[nbr[k][j], ci]        (reference call sites spconv_unet_v1m1_base.py:47-66,135-177;
//                                                               the arithmetic lives in spconv's backward)
//
// The contraction runs over voxel rows j, and both operands are stored row-major [row][channel] in HBM.  tcgen05's
// kind::f16 accepts MN-major operands (the contraction index runs over the ROWS of the shared-memory tile), so a
// gathered row block goes into shared memory exactly as it is gathered -- 128-byte runs of 64 channels, 8-row swizzle
// groups, the same tile the forward kernel builds -- and is read "transposed" by the tensor core.  The first-generation
// kernel (wgrad_umma.cu) needed K-major TF32 tiles: a separate dY transpose kernel per call and 4-byte transposing
// stores for every gathered element.
//
//   fp32 storage: bf16x3 -- every value split on chip into two bf16 halves (hi + lo = 16 significand bits), products
//                 lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM (same arithmetic as the forward kernel);
//   bf16 storage: one MMA per k-step, rows copied with 16-byte cp.async.
//
// CTA = (row chunk, kernel offset k, 128-wide slice of Cout).  Stage = 32 consecutive tile-order rows (two MMA k-steps);
// with mask-sorted rows most 32-row blocks of an offset hold no pair at all and are skipped through the rulebook's
// per-block activity bytes.  Warps 0-3 gather (dY rows through `order`, X rows through nbr[k]), warp 4 issues the MMAs
// (M = 128 output channels, N = Cin <= 256), warps 0-3 finally add the partial dW_k into global memory with 128-bit
// reductions.  Algorithmic bytes per call: N (Cin + Cout) b + K Cin Cout 4 + 4 K N.
#include "pv2_common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace {

using namespace pv2;

constexpr int kRows = 32;            // contraction rows per stage
constexpr int kPanelBytes = kRows * 128;   // one 64-channel panel of a stage: 32 rows x 128 B
constexpr int kMaxStagesW = 6;
constexpr int kGroupsW = 3;          // producer groups of 4 warps; group g gathers stages g, g + 3, ... of the active list
constexpr int kProducerThreadsW = kGroupsW * 128;
constexpr int kThreadsW = kProducerThreadsW + 32;

struct WMParams {
  const void* x;       // [n_in][cin]   (fp32 or bf16)
  const void* dy;      // [n_out][cout]
  const int32_t* nbr;  // [kvol][n_out] in tile order, or nullptr (identity, kvol = 1)
  const int32_t* order;       // optional [n_out]: tile position -> row of dy
  const uint8_t* blk_active;  // optional [kvol][ceil(n_out/32)]
  float* dw;           // [cout][kvol][dw_row], already offset to this launch's first input channel
  int64_t dw_row, x_row, dy_row;
  int64_t n_out;
  int cin, cout, kvol;       // cin: this launch's slice (<= 256)
  int n_pad;                 // cin rounded up to 16
  int a_panels, b_panels;    // 64-channel panels of the dY slice (1 or 2) and of the X slice (1..4)
  int64_t rows_per_chunk;
  int max_iters;
  int stages;
  uint32_t tmem_cols;
};

__device__ float4 g_zero_page_w[2];

__device__ __forceinline__ void split8_store(uint32_t addr, uint32_t lo_delta, const float4& v0, const float4& v1) {
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

// kBf16: storage type of x / dy.  kBP: 64-channel panels of X a thread may hold (2: Cin <= 128, 4: Cin <= 256).
// Stage layout: [A_hi: a_panels x 4 KB][A_lo][B_hi: b_panels x 4 KB][B_lo] (the lo halves only exist for fp32 storage).
template <bool kBf16, int kBP>
__global__ void __launch_bounds__(kThreadsW, 1) umma_wgrad_mn_kernel(const WMParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k = blockIdx.y;
  const int co0 = blockIdx.z * 128;
  const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_chunk;
  int64_t r_end = r_begin + p.rows_per_chunk;
  if (r_end > p.n_out) r_end = p.n_out;
  const int n_iters = (int)((r_end - r_begin + kRows - 1) / kRows);

  constexpr int kHalves = kBf16 ? 1 : 2;
  const int a_bytes = p.a_panels * kPanelBytes, b_bytes = p.b_panels * kPanelBytes;
  const int stage_bytes = kHalves * (a_bytes + b_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStagesW;
  uint64_t* tmem_full_bar = bars + 2 * kMaxStagesW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStagesW + 1);
  int* n_act_s = reinterpret_cast<int*>(tmem_slot + 1);
  uint16_t* list_s = reinterpret_cast<uint16_t*>(tmem_slot + 2);
  uint8_t* flag_s = reinterpret_cast<uint8_t*>(list_s + p.max_iters);

  // active 32-row blocks of this (chunk, offset)
  int n_act = n_iters;
  if (p.nbr != nullptr) {
    if (p.blk_active != nullptr) {
      const int64_t nblk = (p.n_out + kRows - 1) / kRows;
      const uint8_t* ba = p.blk_active + (int64_t)k * nblk + r_begin / kRows;
      for (int st = tid; st < n_iters; st += kThreadsW) flag_s[st] = __ldg(&ba[st]);
    } else {
      for (int st = warp; st < n_iters; st += kThreadsW / 32) {
        const int64_t pos = r_begin + (int64_t)st * kRows + lane;
        const int32_t src = (pos < r_end) ? __ldg(&p.nbr[(int64_t)k * p.n_out + pos]) : -1;
        const bool any = __any_sync(0xffffffffu, src >= 0);
        if (lane == 0) flag_s[st] = any ? 1 : 0;
      }
    }
    __syncthreads();
    if (warp == 0) {
      int cnt = 0;
      for (int b0 = 0; b0 < n_iters; b0 += 32) {
        const int st = b0 + lane;
        const bool on = st < n_iters && flag_s[st] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (on) list_s[cnt + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)st;
        cnt += __popc(bal);
      }
      if (lane == 0) *n_act_s = cnt;
    }
    __syncthreads();
    n_act = *n_act_s;
    if (n_act == 0) return;
  }
  auto stage_at = [&](int it) -> int { return (p.nbr != nullptr) ? (int)list_s[it] : it; };

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 128); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == kGroupsW * 4) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kGroupsW * 4) {
    // ------------------------------- producers -------------------------------
    // Group g owns stages g, g + kGroupsW, ... of the active list, so up to kGroupsW stage gathers are in flight per CTA.
    // Within a group: 8 threads per row (one 16-byte bf16 piece = 8 channels of every 64-channel panel each), 16 rows
    // per pass, two passes -> a quarter warp stores one whole 128-byte row: conflict-free with the 128-byte swizzle.
    // The row indices (dY row through `order`, X row through the neighbour map) are fetched one own-stage ahead.
    const int grp = warp >> 2;
    const int tg = tid & 127;
    const int piece = tg & 7, rbase = tg >> 3;            // rows rbase, rbase + 16
    const uint32_t off0 = sw128_offset(rbase, piece);      // row rbase + 16 -> + 2048
    const uint8_t* zp = reinterpret_cast<const uint8_t*>(g_zero_page_w);
    constexpr int eb = kBf16 ? 2 : 4;
    const int64_t ch_a = co0 + piece * 8;                  // + 64 per panel
    const int ch_b = piece * 8;
    auto fetch_idx = [&](int it, int64_t (&jrow)[2], int32_t (&src)[2]) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        jrow[rr] = -1; src[rr] = -1;
        if (it < n_act) {
          const int64_t pos = r_begin + (int64_t)stage_at(it) * kRows + rbase + 16 * rr;
          if (pos < r_end) {
            jrow[rr] = (p.order != nullptr) ? (int64_t)__ldg(&p.order[pos]) : pos;
            src[rr] = (p.nbr != nullptr) ? __ldg(&p.nbr[(int64_t)k * p.n_out + pos]) : (int32_t)pos;
          }
        }
      }
    };
    int64_t jrow[2]; int32_t src[2];
    fetch_idx(grp, jrow, src);
    const int lag = kBf16 ? 1 : 0;     // own-stage iterations between issuing the copies and signalling them (cp.async)
    int it_prev = -1;
    for (int it = grp; it < n_act + lag * kGroupsW; it += kGroupsW) {
      if (it < n_act) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes), b_base = a_base + kHalves * a_bytes;
        const uint8_t* a_src[2][2];
        const uint8_t* b_src[2][kBP];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
          for (int pn = 0; pn < 2; ++pn) {
            const int64_t ch = ch_a + 64 * pn;
            a_src[rr][pn] = (jrow[rr] >= 0 && pn < p.a_panels && ch < p.cout)
                                ? reinterpret_cast<const uint8_t*>(p.dy) + (jrow[rr] * p.dy_row + ch) * eb : zp;
          }
#pragma unroll
          for (int pn = 0; pn < kBP; ++pn) {
            const int ch = ch_b + 64 * pn;
            b_src[rr][pn] = (src[rr] >= 0 && ch < p.cin)
                                ? reinterpret_cast<const uint8_t*>(p.x) + ((int64_t)src[rr] * p.x_row + ch) * eb : zp;
          }
        }
        if constexpr (kBf16) {
          fetch_idx(it + kGroupsW, jrow, src);            // next own stage's indices, in flight across the wait
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
            for (int pn = 0; pn < 2; ++pn)
              if (pn < p.a_panels)
                cp_async_16(a_base + pn * kPanelBytes + off0 + rr * 2048, a_src[rr][pn], a_src[rr][pn] != zp ? 16u : 0u);
#pragma unroll
            for (int pn = 0; pn < kBP; ++pn)
              if (pn < p.b_panels && ch_b + 64 * pn < p.n_pad)
                cp_async_16(b_base + pn * kPanelBytes + off0 + rr * 2048, b_src[rr][pn], b_src[rr][pn] != zp ? 16u : 0u);
          }
        } else {
          float4 va[2][2][2], vb[2][kBP][2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
            for (int pn = 0; pn < 2; ++pn) {
              va[rr][pn][0] = __ldg(reinterpret_cast<const float4*>(a_src[rr][pn]));
              va[rr][pn][1] = __ldg(reinterpret_cast<const float4*>(a_src[rr][pn]) + 1);
            }
#pragma unroll
            for (int pn = 0; pn < kBP; ++pn) {
              vb[rr][pn][0] = __ldg(reinterpret_cast<const float4*>(b_src[rr][pn]));
              vb[rr][pn][1] = __ldg(reinterpret_cast<const float4*>(b_src[rr][pn]) + 1);
            }
          }
          fetch_idx(it + kGroupsW, jrow, src);            // next own stage's indices, in flight across the wait
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
            for (int pn = 0; pn < 2; ++pn)
              if (pn < p.a_panels)
                split8_store(a_base + pn * kPanelBytes + off0 + rr * 2048, (uint32_t)a_bytes, va[rr][pn][0], va[rr][pn][1]);
#pragma unroll
            for (int pn = 0; pn < kBP; ++pn)
              if (pn < p.b_panels && ch_b + 64 * pn < p.n_pad)
                split8_store(b_base + pn * kPanelBytes + off0 + rr * 2048, (uint32_t)b_bytes, vb[rr][pn][0], vb[rr][pn][1]);
          }
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&full_bar[s]));
        }
      }
      if constexpr (kBf16) {
        cp_async_commit();
        if (it_prev >= 0) {
          cp_async_wait<1>();                              // the previous own stage's copies have landed
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&full_bar[it_prev % p.stages]));
        }
        it_prev = (it < n_act) ? it : -1;
      }
    }
    // ------------------------------- epilogue (warps 0-3: TMEM lanes 32 w .. 32 w + 31) -------------------------------
    if (warp < 4) {
      mbar_wait(smem_u32(tmem_full_bar), 0);
      tc_fence_after();
      const int co = co0 + warp * 32 + lane;
      for (int col0 = 0; col0 < p.n_pad; col0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col0, v);
        tmem_ld_wait();
        if (co >= p.cout) continue;
        float* dst = p.dw + ((int64_t)co * p.kvol + k) * p.dw_row + col0;
        if (col0 + 16 <= p.cin && (p.dw_row & 3) == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(__uint_as_float(v[4 * q])),
                         "f"(__uint_as_float(v[4 * q + 1])), "f"(__uint_as_float(v[4 * q + 2])),
                         "f"(__uint_as_float(v[4 * q + 3])) : "memory");
        } else {
          for (int i = 0; i < 16 && col0 + i < p.cin; ++i) atomicAdd(dst + i, __uint_as_float(v[i]));
        }
      }
    }
    tc_fence_before();
  } else {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc = make_idesc_mn(1 /*BF16*/, 128, p.n_pad);
    for (int it = 0; it < n_act; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_base = a_base + kHalves * a_bytes;
        // MN-major: panel stride (LBO) = 4 KB, 8-row group stride (SBO) = 1 KB; a k-step = 16 rows = 2 KB further
#pragma unroll
        for (int ks = 0; ks < kRows / 16; ++ks) {
          const uint64_t dah = smem_desc_mnmajor_sw128(a_base + ks * 2048, kPanelBytes, 1024);
          const uint64_t dbh = smem_desc_mnmajor_sw128(b_base + ks * 2048, kPanelBytes, 1024);
          if constexpr (kBf16) {
            umma_bf16(tmem_base, dah, dbh, idesc, (it | ks) != 0 ? 1u : 0u);
          } else {
            const uint64_t dal = smem_desc_mnmajor_sw128(a_base + a_bytes + ks * 2048, kPanelBytes, 1024);
            const uint64_t dbl = smem_desc_mnmajor_sw128(b_base + b_bytes + ks * 2048, kPanelBytes, 1024);
            umma_bf16(tmem_base, dal, dbh, idesc, (it | ks) != 0 ? 1u : 0u);
            umma_bf16(tmem_base, dah, dbl, idesc, 1u);
            umma_bf16(tmem_base, dah, dbh, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
        if (it == n_act - 1) umma_commit(smem_u32(tmem_full_bar));
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kGroupsW * 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

template <bool kBf16, int kBP>
int launch_wgrad_mn(WMParams p, cudaStream_t stream) {
  p.n_pad = (p.cin + 15) / 16 * 16;
  p.b_panels = (p.n_pad + 63) / 64;
  p.a_panels = (p.cout > 64) ? 2 : 1;   // the M = 128 MMA reads two panels; with one, lanes >= 64 accumulate the next
                                         // tile's bytes (finite or not) into accumulator rows nobody reads
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < p.n_pad) p.tmem_cols <<= 1;
  constexpr int kHalves = kBf16 ? 1 : 2;
  const int stage_bytes = kHalves * (p.a_panels + p.b_panels) * kPanelBytes;
  const int m_tiles = (p.cout + 127) / 128;
  // ~3 waves of one CTA per SM (the CTAs of different offsets differ a lot in active blocks), chunks of >= 512 rows
  int64_t chunks = (3LL * PV2_SM_COUNT + (int64_t)p.kvol * m_tiles - 1) / ((int64_t)p.kvol * m_tiles);
  const int64_t max_chunks = (p.n_out + 511) / 512;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int64_t min_chunks = (p.n_out + 65535) / 65536;
  if (chunks < min_chunks) chunks = min_chunks;
  p.rows_per_chunk = ((p.n_out + chunks - 1) / chunks + kRows - 1) / kRows * kRows;
  chunks = (p.n_out + p.rows_per_chunk - 1) / p.rows_per_chunk;
  p.max_iters = (int)(p.rows_per_chunk / kRows);
  // + one panel the M = 128 MMA may read past a one-panel A tile of the last stage
  const int fixed = (2 * kMaxStagesW + 2) * 8 + 64 + 1024 + 3 * p.max_iters + 16 + kPanelBytes;
  int stages = (220 * 1024 - fixed) / stage_bytes;
  if (stages > kMaxStagesW) stages = kMaxStagesW;
  // a group may only wait one phase ahead on a stage's mbarrier (parity waits alias every second phase): groups <= stages
  if (stages < kGroupsW) return PV2_EUNSUPPORTED;
  p.stages = stages;
  static bool done[64] = {};
  {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64 || !done[dev]) {
      e = cudaFuncSetAttribute(umma_wgrad_mn_kernel<kBf16, kBP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return (int)e;
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
  }
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  dim3 grid((unsigned)chunks, (unsigned)p.kvol, (unsigned)m_tiles);
  umma_wgrad_mn_kernel<kBf16, kBP><<<grid, kThreadsW, smem, stream>>>(p);
  PV2_DONE(1);
}

}  // namespace

extern "C" {

// x [n_in][cin], dy [n_out][cout] in `dtype` (PV2_F32: bf16x3 arithmetic; PV2_BF16), dw fp32 [cout][kvol][cin] must be
// zeroed by the caller.  Returns PV2_EUNSUPPORTED for shapes this kernel does not take.
int pv2_wgrad_mn(const void* x, const void* dy, const int32_t* nbr, const int32_t* order, const uint8_t* blk_active,
                 float* dw, int64_t n_in, int64_t n_out, int cin, int cout, int kvol, int dtype, void* stream_) {
  PV2_CHECK_ARG(n_in >= 0 && n_out >= 0 && cin > 0 && cout > 0 && kvol > 0);
  if (n_out == 0 || n_in == 0) return 0;
  PV2_CHECK_ARG(x && dy && dw);
  if (dtype != PV2_F32 && dtype != PV2_BF16) return PV2_EUNSUPPORTED;
  if ((cin % 8) || (cout % 8) || kvol > 65535) return PV2_EUNSUPPORTED;
  if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15) || ((uintptr_t)dw & 15)) return PV2_EUNSUPPORTED;
  if (nbr == nullptr && (kvol != 1 || n_in != n_out)) return PV2_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  WMParams p{};
  p.x = x; p.dy = dy; p.nbr = nbr; p.order = (nbr != nullptr) ? order : nullptr;
  p.blk_active = (nbr != nullptr) ? blk_active : nullptr;
  p.dw = dw; p.dw_row = cin; p.x_row = cin; p.dy_row = cout;
  p.n_out = n_out; p.cout = cout; p.kvol = kvol;
  const int eb = dtype == PV2_BF16 ? 2 : 4;
  // the accumulator holds N = Cin <= 256 columns: wider inputs (dec3's 384-channel concat) run as balanced slices
  const int nsl = (cin + 255) / 256;
  const int per = ((cin + nsl - 1) / nsl + 15) / 16 * 16;
  for (int ci0 = 0; ci0 < cin; ci0 += per) {
    WMParams q = p;
    q.cin = (cin - ci0 < per) ? cin - ci0 : per;
    q.x = (const char*)x + (size_t)ci0 * eb;
    q.dw = dw + ci0;
    const bool wide = q.cin > 128;
    const int rc = dtype == PV2_BF16 ? (wide ? launch_wgrad_mn<true, 4>(q, stream) : launch_wgrad_mn<true, 2>(q, stream))
                                     : (wide ? launch_wgrad_mn<false, 4>(q, stream) : launch_wgrad_mn<false, 2>(q, stream));
    if (rc != 0) return rc;
  }
  return 0;
}

}  // extern "C"
