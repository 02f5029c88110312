// Thin inline-PTX layer for the Blackwell (sm_100a) tensor-core path: mbarrier, cp.async, proxy fences,
// tcgen05 alloc / mma / commit / ld, and the shared-memory / instruction descriptors.
// Bit layouts follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor, UMMA::InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pv2 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifdef PV2_MBAR_DEBUG
// Development build (PV2_MBAR_DEBUG): a wait that times out records (block, thread, barrier address, parity) and RETURNS,
// so that the kernel drains and the host can read who was waiting for what (pv2_debug_dump_waits).
static __device__ unsigned g_wait_log[4 * 2048];
static __device__ unsigned g_wait_n;
#endif
// Bounded wait (~2 s of SM clocks): a protocol bug becomes a trap (launch error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
#ifdef PV2_MBAR_DEBUG
    if (clock64() - t0 > 200000000LL) {
      if ((threadIdx.x & 31) == 0) {
        const unsigned i = atomicAdd(&g_wait_n, 1u);
        if (i < 2048) { g_wait_log[4 * i] = blockIdx.x; g_wait_log[4 * i + 1] = threadIdx.x; g_wait_log[4 * i + 2] = bar; g_wait_log[4 * i + 3] = parity; }
      }
      return;
    }
#else
    if (clock64() - t0 > 4000000000LL) __trap();
#endif
  }
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes (st.shared / cp.async) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA (cp.async.bulk.tensor) --------------------------------------------------------------------------
// the barrier additionally waits for `bytes` of asynchronous-proxy transactions in its current phase
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D tile load: box (c0 .. c0 + box0, c1 .. c1 + box1) of the tensor map -> shared memory (layout / swizzle as encoded
// in the map), completion signalled on `bar` as transaction bytes.  Out-of-bounds elements are zero-filled.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// four rows (r0..r3, any order, out-of-bounds rows zero-filled) x the box width starting at column c0 -> 4 consecutive
// box rows in shared memory
__device__ __forceinline__ void tma_gather4_2d(uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t r0,
                                               int32_t r1, int32_t r2, int32_t r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// ---- cp.async (LDGSTS) ----------------------------------------------------------------------------------
// 16-byte copy; src_bytes = 0 zero-fills the destination (used for missing neighbours / padding)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  // no "memory" clobber: ordering against the consumers is established by commit/wait_group + fence + mbarrier, and a
  // clobber here would chain every index load behind the previous copy (measured: 170 cycles per copy instead of ~10)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- tcgen05 --------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues on behalf of the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread l of warp w reads lane 32*(w%4)+l)
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------------
// K-major operand tile stored as rows of 128 bytes, 8-row groups of 1024 bytes, 128-byte swizzle
// (16-byte chunk c of row r lives at chunk c ^ (r & 7)).  The tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t smem_desc_kmajor_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);      // start address, 16-byte units
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                      // layout type: SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk16) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4));
}
// MN-major operand tile (kind::f16 only): the contraction index k runs over the ROWS of the tile.  Rows of 128 bytes
// (64 consecutive MN elements), 8-row groups of 1024 bytes with the 128-byte swizzle, i.e. physically the same tile the
// K-major descriptor above describes -- read "transposed".  `lbo_bytes` = distance between 64-element MN panels,
// `sbo_bytes` = distance between 8-row k groups (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units,
// cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t smem_desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 (BF16 x BF16 -> F32) / kind::tf32 (TF32 x TF32 -> F32), both operands K-major, M = 128
__host__ __device__ __forceinline__ uint32_t make_idesc(int fmt /*1 = BF16, 2 = TF32*/, int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                    // accumulator format F32
  d |= (uint32_t)fmt << 7;         // A format
  d |= (uint32_t)fmt << 10;        // B format
  d |= (uint32_t)(n >> 3) << 17;   // N / 8
  d |= (uint32_t)(m >> 4) << 24;   // M / 16
  return d;
}
// same with both operands MN-major (bits 15 / 16): the weight-gradient contraction runs over voxel rows
__host__ __device__ __forceinline__ uint32_t make_idesc_mn(int fmt, int m, int n) {
  return make_idesc(fmt, m, n) | (1u << 15) | (1u << 16);
}

}  // namespace pv2
