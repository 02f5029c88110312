// Weight gradient of the sparse convolutions (and of the render MLP's dense layers) on the Blackwell tensor cores.
//
//   dW[co, k, ci] += sum_j dY[j, co] * X[nbr[k][j], ci]            (wgrad of SubMConv3d / SparseConv3d / Inverse)
//
// This is a GEMM whose contraction runs over the voxel rows j.  tcgen05's TF32 path takes K-major operands only
// (measured: setting the MN-major bits of the instruction descriptor with kind::tf32 yields an all-zero accumulator),
// so both operands are brought into K-major form, i.e. "channel x 32 consecutive rows" tiles:
//   * A = dY^T: a small transposing kernel writes dY once as [Cout][N] (rows in tile order) and the tile rows are then
//     plain 128-byte runs copied RAW with cp.async and split in place into TF32 hi / lo halves once they have landed;
//   * B = X[nbr[k][j]]^T: gathered rows cannot be pre-transposed (a different permutation per offset k), so each lane
//     owns one row j, reads its channels with 128-bit loads, splits them into TF32 hi/lo and writes them as a column of
//     the swizzled tile — lanes of a warp hit 32 different banks, so the transposing stores are conflict-free.
// One CTA owns (row chunk, kernel offset k, 128-wide slice of Cout), accumulates its partial dW_k in TMEM
// (128 lanes = output channels, Cin columns) over all its rows with 3xTF32 arithmetic (hi*hi + lo*hi + hi*lo) and
// finally adds it into dW with 128-bit vector reductions.
#include "pv2_common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace {

using namespace pv2;

constexpr int kRowsPerStage = 32;   // contraction rows per pipeline stage = one 128-byte K-major line (4 MMA k-steps)
constexpr int kMaxABytes = 128 * 128;  // 128 output channels x 128 B (the A tile holds only the channels that exist)
constexpr int kMaxStages = 4;
constexpr int kThreads = 160;

struct WGParams {
  const float* x;      // [n_in] rows of cin floats, row stride x_row, optional additive second half at +x_lo
  int64_t x_row, x_lo;
  const float* dyt;    // dy^T in tile order: [cout][np] raw fp32 (split into TF32 halves on chip)
  int64_t np;          // padded row count (multiple of 32) = row length of dyt
  const int32_t* nbr;  // [kvol][n_out] in tile order (nbr[k][pos] feeds output row order[pos]) or nullptr (identity)
  const int32_t* order;  // optional [n_out]: position -> row (mask-sorted, pv2_rulebook_row_order); dyt is in this order
  const uint8_t* blk_active;  // optional [kvol][ceil(n_out/32)]: 1 iff the 32-row block has a pair at offset k
  int max_iters;       // rows_per_chunk / 32 (capacity of the active-stage list)
  int a_bytes;         // bytes of one A tile half: round8(min(128, cout)) rows x 128 B.  The MMA (M = 128) reads past it
                       // into the neighbouring tiles; those rows only feed accumulator lanes >= cout, which nobody reads
  float* dw;           // [cout][kvol][dw_row] (already offset to this launch's first input channel)
  int64_t dw_row;      // full Cin of the weight tensor (cin below may be a <= 256-wide slice of it)
  int64_t n_out;
  int cin, cout, kvol;
  int n_pad;           // cin rounded up to 16
  int64_t rows_per_chunk;
  int stages;
  uint32_t tmem_cols;
};

__device__ __forceinline__ void split_tf32_dev(float v, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
  lo = __uint_as_float(l);
}

// kPre: x is split-precision (value = x[..] + x[.. + x_lo]); a template parameter so that the plain path carries no
// (predicated-off but still scoreboard-waiting) adds between its gather loads.
// kUnits: float4 units of a gathered row each lane may hold (8: Cin <= 128, small enough register footprint for two
// resident CTAs per SM; 16: Cin <= 256).
template <bool kPre, int kUnits>
__global__ void __launch_bounds__(kThreads, kUnits == 8 ? 2 : 1) umma_wgrad_kernel(const WGParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k = blockIdx.y;
  const int co0 = blockIdx.z * 128;
  const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_chunk;
  int64_t r_end = r_begin + p.rows_per_chunk;
  if (r_end > p.n_out) r_end = p.n_out;
  const int n_iters = (int)((r_end - r_begin + kRowsPerStage - 1) / kRowsPerStage);

  const int b_bytes = p.n_pad * 128;
  const int kABytes = p.a_bytes;
  const int a_rows = kABytes >> 7;
  const int stage_bytes = 2 * (kABytes + b_bytes);   // [A_hi][A_lo][B_hi][B_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
  int* n_act_s = reinterpret_cast<int*>(tmem_slot + 1);
  uint16_t* list_s = reinterpret_cast<uint16_t*>(tmem_slot + 2);   // [max_iters] active 32-row stages, ascending
  uint8_t* flag_s = reinterpret_cast<uint8_t*>(list_s + p.max_iters);

  // Which 32-row stages of this (row chunk, offset k) contain at least one pair?  With mask-sorted rows most stages of
  // an offset are empty (a surface voxel has ~9 of 27 neighbours); they are skipped by producers and MMA issuer alike.
  int n_act = n_iters;
  if (p.nbr != nullptr) {
    if (p.blk_active != nullptr) {   // precomputed per rulebook: one coalesced byte per stage
      const int64_t nblk = (p.n_out + kRowsPerStage - 1) / kRowsPerStage;
      const uint8_t* ba = p.blk_active + (int64_t)k * nblk + r_begin / kRowsPerStage;
      for (int st = tid; st < n_iters; st += kThreads) flag_s[st] = __ldg(&ba[st]);
    } else {
      for (int st = warp; st < n_iters; st += kThreads / 32) {
        const int64_t pos = r_begin + (int64_t)st * kRowsPerStage + lane;
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
    if (n_act == 0) return;   // uniform for the CTA; nothing allocated yet
  }
  auto stage_at = [&](int it) -> int { return (p.nbr != nullptr) ? (int)list_s[it] : it; };

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 128); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------- producers -------------------------------
    const int piece = tid & 7;   // A: 16-byte piece (4 rows j) of a channel's 128-byte line
    const int cbase = tid >> 3;  // A: channels cbase, cbase+16, ...
    const uint32_t a_off = sw128_offset(cbase, piece);
    // A stage is handed to the MMA warp `lag` iterations after its copies were issued and is refilled `stages`
    // iterations later: lag = stages / 2 leaves the copies and the MMAs (issue -> commit -> barrier) about half of the
    // ring each to complete in, instead of making every refill wait for MMAs issued one iteration earlier.
    const int lag = p.stages >= 2 ? p.stages / 2 : 1;
    const int n_units = p.n_pad / 4;
    constexpr int kMaxUnitsPerWarp = kUnits;  // n_pad <= 256 -> 64 units / 4 warps
    const int upw = n_units / 4;          // n_pad is a multiple of 16 -> n_units is a multiple of 4
    const uint32_t col = (uint32_t)((lane & 3) * 4);
    const int jc = lane >> 2;
    auto load_src = [&](int it) -> int32_t {
      if (it >= n_act) return -1;
      const int64_t pos = r_begin + (int64_t)stage_at(it) * kRowsPerStage + lane;
      if (pos >= r_end) return -1;
      return (p.nbr != nullptr) ? __ldg(&p.nbr[(int64_t)k * p.n_out + pos]) : (int32_t)pos;
    };
    // B operand rows travel global -> registers -> (split, transposed) shared memory.  The loads of stage it + 1 are
    // issued before stage it is split and stored (register double buffer) and the neighbour index is fetched two stages
    // ahead, so one gather is always in flight per warp.
    auto load_rows = [&](int32_t src, float4 (&xv)[kMaxUnitsPerWarp]) {
      const float* xr = p.x + (int64_t)(src >= 0 ? src : 0) * p.x_row;
#pragma unroll
      for (int q = 0; q < kMaxUnitsPerWarp; ++q) {
        const int u = warp * upw + q;   // each lane reads one contiguous run of its row (whole sectors, fetched once)
        xv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < upw && src >= 0 && u * 4 < p.cin) xv[q] = __ldg(reinterpret_cast<const float4*>(xr) + u);
      }
      if constexpr (kPre) {
        float4 xl[kMaxUnitsPerWarp];
#pragma unroll
        for (int q = 0; q < kMaxUnitsPerWarp; ++q) {
          const int u = warp * upw + q;
          xl[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < upw && src >= 0 && u * 4 < p.cin) xl[q] = __ldg(reinterpret_cast<const float4*>(xr + p.x_lo) + u);
        }
#pragma unroll
        for (int q = 0; q < kMaxUnitsPerWarp; ++q) { xv[q].x += xl[q].x; xv[q].y += xl[q].y; xv[q].z += xl[q].z; xv[q].w += xl[q].w; }
      }
    };
    float4 xcur[kMaxUnitsPerWarp], xnext[kMaxUnitsPerWarp];
    load_rows(load_src(0), xcur);
    int32_t src_next = load_src(1);
    for (int it = 0; it < n_act + lag; ++it) {
      if (it < n_act) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        load_rows(src_next, xnext);          // stage it + 1 (all zeros past the end)
        src_next = load_src(it + 2);
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        uint8_t* a_hi = smem + (size_t)s * stage_bytes;
        uint8_t* b_hi = a_hi + 2 * kABytes;
        const int64_t j0 = r_begin + (int64_t)stage_at(it) * kRowsPerStage;   // position (column of dyt), not row id
        // A: rows = output channels, 128 B = dy^T[co][j0 .. j0+31]
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = cbase + 16 * i;
          if (c >= a_rows) continue;              // rows the tile does not hold
          const bool ok = (co0 + c) < p.cout;
          const float* g = ok ? p.dyt + ((int64_t)(co0 + c) * p.np + j0 + piece * 4) : p.dyt;
          const uint32_t dst = smem_u32(a_hi) + a_off + i * 2048;
          cp_async_16(dst, g, ok ? 16u : 0u);   // raw fp32 into the hi tile; split in place once it has landed
        }
        // B: lane = row j; transposing stores (bank = f(lane) only -> conflict-free)
#pragma unroll
        for (int q = 0; q < kMaxUnitsPerWarp; ++q) {
          const int u = warp * upw + q;
          if (q < upw) {
            const float vv[4] = {xcur[q].x, xcur[q].y, xcur[q].z, xcur[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // hi = top 19 bits (exactly a TF32 number), lo = v - hi (exact)
              const float h = __uint_as_float(__float_as_uint(vv[e]) & 0xffffe000u);
              const float l = vv[e] - h;
              const uint32_t off = sw128_offset(u * 4 + e, jc) + col;
              *reinterpret_cast<float*>(b_hi + off) = h;
              *reinterpret_cast<float*>(b_hi + b_bytes + off) = l;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < kMaxUnitsPerWarp; ++q) xcur[q] = xnext[q];
      }
      cp_async_commit();
      if (it >= lag) {
        switch (lag) {
          case 1: cp_async_wait<1>(); break;
          case 2: cp_async_wait<2>(); break;
          default: cp_async_wait<3>(); break;
        }
        {
          // this thread's eight A pieces of stage (it - lag) have landed: split them in place
          const uint32_t a2 = smem_u32(smem + (size_t)((it - lag) % p.stages) * stage_bytes) + a_off;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (cbase + 16 * i >= a_rows) continue;
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a2 + i * 2048) : "memory");
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a2 + i * 2048), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a2 + i * 2048 + kABytes), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&full_bar[(it - lag) % p.stages]));
      }
    }
    // ------------------------------- epilogue -------------------------------
    if (n_act > 0) {
      mbar_wait(smem_u32(tmem_full_bar), 0);
      tc_fence_after();
      const int co = co0 + warp * 32 + lane;
      for (int col0 = 0; col0 < p.n_pad; col0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col0, v);
        tmem_ld_wait();
        if (co >= p.cout) continue;
        float* dst = p.dw + ((int64_t)co * p.kvol + k) * p.dw_row + col0;
        if (col0 + 16 <= p.cin) {
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
    const uint32_t idesc = make_idesc(2 /*TF32*/, 128, p.n_pad);
    for (int it = 0; it < n_act; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dah = smem_desc_kmajor_sw128(smem_u32(smem + (size_t)s * stage_bytes));
        const uint64_t dal = dah + (uint64_t)(kABytes >> 4);
        const uint64_t dbh = dal + (uint64_t)(kABytes >> 4);
        const uint64_t dbl = dbh + (uint64_t)(b_bytes >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_tf32(tmem_base, dal + 2 * ks, dbh + 2 * ks, idesc, (it | ks) != 0 ? 1u : 0u);
          umma_tf32(tmem_base, dah + 2 * ks, dbl + 2 * ks, idesc, 1u);
          umma_tf32(tmem_base, dah + 2 * ks, dbh + 2 * ks, idesc, 1u);
        }
        umma_commit(smem_u32(&empty_bar[s]));
        if (it == n_act - 1) umma_commit(smem_u32(tmem_full_bar));
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dy [rows][cols] (row stride in_row, optional additive half at +in_lo) -> out[cols][np] = dy^T (rows taken in `order`),
// zero-padded for rows in [rows, np).  32 x 32 tiles through shared memory: coalesced on both sides.
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ in, int64_t in_row, int64_t in_lo,
                                                             const int32_t* __restrict__ order, float* __restrict__ out,
                                                             int64_t rows, int cols, int64_t np) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 8 * i;
    const int c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      const int64_t rr = (order != nullptr) ? (int64_t)__ldg(&order[r]) : r;   // column r of out = row order[r] of in
      v = in[rr * in_row + c];
      if (in_lo) v += in[rr * in_row + in_lo + c];
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    const int64_t r = r0 + tx;
    if (c < cols && r < np) {
      out[(int64_t)c * np + r] = tile[tx][ty + 8 * i];
    }
  }
}

int launch_wgrad(WGParams p, cudaStream_t stream) {
  p.n_pad = (p.cin + 15) / 16 * 16;
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < p.n_pad) p.tmem_cols <<= 1;
  {
    const int rows = p.cout < 128 ? (p.cout + 7) / 8 * 8 : 128;
    p.a_bytes = rows * 128;
  }
  const int stage_bytes = 2 * (p.a_bytes + p.n_pad * 128);
  // + the active-stage list (added below) + what the M = 128 MMA may read past a short A tile of the last stage
  int fixed = (2 * kMaxStages + 2) * 8 + 64 + 1024 + (kMaxABytes - p.a_bytes);
  const int m_tiles = (p.cout + 127) / 128;
  // ~4 CTAs per SM in total, chunks of at least 256 rows
  int64_t chunks = (4LL * PV2_SM_COUNT + (int64_t)p.kvol * m_tiles - 1) / ((int64_t)p.kvol * m_tiles);
  const int64_t max_chunks = (p.n_out + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int64_t min_chunks = (p.n_out + 65535) / 65536;   // <= 2048 stages per CTA (uint16 list in shared memory)
  if (chunks < min_chunks) chunks = min_chunks;
  p.rows_per_chunk = ((p.n_out + chunks - 1) / chunks + kRowsPerStage - 1) / kRowsPerStage * kRowsPerStage;
  chunks = (p.n_out + p.rows_per_chunk - 1) / p.rows_per_chunk;
  p.max_iters = (int)(p.rows_per_chunk / kRowsPerStage);
  fixed += 3 * p.max_iters + 16;
  // two resident CTAs per SM when two stages fit in half of the shared memory (the kernel is latency-bound: a second
  // CTA's gathers overlap the first one's splits and stores); otherwise one CTA with up to four stages
  static int want_ctas = -1;
  if (want_ctas < 0) { const char* e = getenv("PV2_WGRAD_CTAS"); want_ctas = e ? atoi(e) : 2; }
  int stages = (224 * 1024 - fixed) / stage_bytes;
  if (want_ctas >= 2 && (112 * 1024 - fixed) / stage_bytes >= 2) stages = (112 * 1024 - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return PV2_EUNSUPPORTED;
  p.stages = stages;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_wgrad_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_wgrad_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_wgrad_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_wgrad_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  dim3 grid((unsigned)chunks, (unsigned)p.kvol, (unsigned)m_tiles);
  const bool small = p.n_pad <= 128;
  if (p.x_lo != 0) {
    if (small) umma_wgrad_kernel<true, 8><<<grid, kThreads, smem, stream>>>(p);
    else umma_wgrad_kernel<true, 16><<<grid, kThreads, smem, stream>>>(p);
  } else {
    if (small) umma_wgrad_kernel<false, 8><<<grid, kThreads, smem, stream>>>(p);
    else umma_wgrad_kernel<false, 16><<<grid, kThreads, smem, stream>>>(p);
  }
  PV2_DONE(1);
}

}  // namespace

extern "C" {

size_t pv2_wgrad_workspace_bytes(int64_t n_in, int64_t n_out, int cin, int cout) {
  if (n_in < 0 || n_out < 0) return 0;
  const int64_t np = (n_out + 31) / 32 * 32;
  return ((size_t)np * cout * 4 + 255) / 256 * 256;
}

// fp32 only.  x [n_in][cin] / dy [n_out][cout] may carry a row stride and an additive second half (+lo offset), so the
// render MLP can hand over its split-precision activations directly.  Returns PV2_EUNSUPPORTED for shapes the
// tensor-core kernel does not take (caller falls back to the SIMT kernel).
int pv2_wgrad_umma(const float* x, int64_t x_row, int64_t x_lo, const float* dy, int64_t dy_row, int64_t dy_lo,
                   const int32_t* nbr, const int32_t* order, const uint8_t* blk_active, float* dw, int64_t n_in,
                   int64_t n_out, int cin, int cout, int kvol,
                   void* workspace, size_t workspace_bytes, void* stream_) {
  PV2_CHECK_ARG(n_in >= 0 && n_out >= 0 && cin > 0 && cout > 0 && kvol > 0);
  if (n_out == 0 || n_in == 0) return 0;
  PV2_CHECK_ARG(x && dy && dw);
  if ((cin % 4) || (cin > 256 && (cin % 16)) || (x_row % 4) || (x_lo % 4) || ((uintptr_t)x & 15)) return PV2_EUNSUPPORTED;
  if (nbr == nullptr && (kvol != 1 || n_in != n_out)) return PV2_EINVAL;
  if (workspace == nullptr || workspace_bytes < pv2_wgrad_workspace_bytes(n_in, n_out, cin, cout) ||
      ((uintptr_t)workspace & 15))
    return PV2_EWORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t np = (n_out + 31) / 32 * 32;
  float* dyt = (float*)workspace;
  dim3 tg((unsigned)(np / 32), (unsigned)((cout + 31) / 32));
  transpose_split_kernel<<<tg, 256, 0, stream>>>(dy, dy_row, dy_lo, order, dyt, n_out, cout, np);
  pv2_note_launches(1);
  WGParams p{};
  p.x = x; p.x_row = x_row; p.x_lo = x_lo; p.dyt = dyt; p.np = np; p.nbr = nbr; p.order = (nbr != nullptr) ? order : nullptr;
  p.blk_active = (nbr != nullptr) ? blk_active : nullptr;
  p.dw = dw;
  p.n_out = n_out; p.cout = cout; p.kvol = kvol; p.dw_row = cin;
  // the accumulator holds N = Cin <= 256 columns: wider inputs (dec3's 384-channel concat) run as balanced slices
  const int nsl = (cin + 255) / 256;
  const int per = ((cin + nsl - 1) / nsl + 15) / 16 * 16;
  for (int ci0 = 0; ci0 < cin; ci0 += per) {
    WGParams q = p;
    q.cin = (cin - ci0 < per) ? cin - ci0 : per;
    q.x = x + ci0;
    q.dw = dw + ci0;
    const int rc = launch_wgrad(q, stream);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // extern "C"
