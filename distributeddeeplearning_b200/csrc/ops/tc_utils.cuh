// tcgen05 / TMEM / TMA / mbarrier PTX wrappers for sm_100a (no CUTLASS dependency).
//
// Conventions used by every GEMM-shaped kernel in this repo
//   * operand tiles live in shared memory as rows of 128 bytes (64 bf16) with the 128B swizzle:
//     16-byte chunk j of row r is stored at chunk (j ^ (r & 7)); tiles are 1024 B aligned.
//     TMA (CU_TENSOR_MAP_SWIZZLE_128B) and the cp.async gather producers both write this image.
//   * K-major operand  : row = M/N index, the 128 B row = 64 consecutive K elements.
//                        descriptor: SBO = 1024 (8 rows), K-advance of 16 elements = +32 B.
//   * MN-major operand : row = K index, the 128 B row = 64 consecutive M/N elements.
//                        descriptor: SBO = 1024 (8 K-rows), LBO = bytes between 64-wide M/N chunks,
//                        K-advance of 16 elements = +2048 B.
//   * accumulators: fp32 in TMEM, D row i -> TMEM lane i, D column j -> TMEM column j.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../common.cuh"

namespace ddl {
namespace tc {

DDL_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

DDL_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------------------------
DDL_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
DDL_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DDL_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
DDL_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// c_wait_hint_ns != 0: pass that suspend-time hint to try_wait (the thread sleeps in hardware until the phase completes
// or the time is up, instead of returning after the short system default and re-issuing the poll loop) — tuning hook
// set_conv_wait_hint, A/B-measured in BASELINE.md
__constant__ unsigned int c_wait_hint_ns;
DDL_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const unsigned int hint = c_wait_hint_ns;
  if (hint != 0u) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
  return ok != 0;
}
// Bounded wait with a post-mortem.  A protocol bug must neither hang the GPU nor kill the context without a trace:
// the first waiter that exceeds the bound records WHERE it was stuck (site id, block, thread, parity) in a device-side
// diagnostic block and raises an abort flag; every other waiter of this and later kernels then drains immediately, so
// the launch terminates (with garbage results) and the host can read the record (conv_timeout_info) and raise.
// `site` = 16 * kernel family (1 one-tile, 2 persistent, 3 deep-ring, 4 wgrad) + role (1 producer waits for a free
// stage, 2 MMA waits for operands, 3 MMA waits for a drained accumulator, 4 epilogue waits for the accumulator).
// The bound is an iteration count (each failed try_wait already suspends the thread for a hardware-defined window),
// which keeps the spin loop short: the ncu source view showed the former clock64()-based loop of the TMA / MMA warps
// taking ~25 % of all issued instructions in short-K kernels.
__device__ unsigned int g_mbar_diag[8];     // [0] abort flag, [1] site, [2] blockIdx.x, [3] threadIdx.x, [4] parity

DDL_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t site = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0) {
      if (*reinterpret_cast<volatile unsigned int*>(&g_mbar_diag[0]) != 0u) return;     // somebody timed out: drain
      if (spins > (c_wait_hint_ns != 0u ? (1u << 21) : (1u << 24))) {     // hinted iterations last up to the hint
        if (atomicCAS(&g_mbar_diag[0], 0u, 1u) == 0u) {
          g_mbar_diag[1] = site;
          g_mbar_diag[2] = blockIdx.x + gridDim.x * blockIdx.y;
          g_mbar_diag[3] = threadIdx.x;
          g_mbar_diag[4] = parity;
          __threadfence();
        }
        return;
      }
    }
  }
}

// ---- thread-block cluster helpers (CTA pairs sharing the weight tile by TMA multicast) ------------------------
DDL_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DDL_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D TMA load delivered to the same shared-memory offset of every CTA in `cta_mask`; each destination CTA's mbarrier
// at the same offset receives the complete_tx for the bytes written into that CTA.
DDL_DEVICE void tma_load_2d_multicast(uint32_t dst_smem, const CUtensorMap* m, int x, int y, uint64_t* bar,
                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset in every CTA of `cta_mask`
DDL_DEVICE void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// ---- CTA pair (cta_group::2): one MMA spans two SMs, each SM holding its 128 rows of A and HALF of B ---------------
template <int COLS>
DDL_DEVICE void tmem_alloc_pair(uint32_t* slot_in_smem) {      // one whole warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot_in_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
DDL_DEVICE void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N columns: N/2 per CTA]; issued by ONE thread of the leader CTA
DDL_DEVICE void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
DDL_DEVICE void umma_commit_pair_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// address of `local` (a shared::cta address) in CTA `rank` of this cluster
DDL_DEVICE uint32_t mapa_shared(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// TMA loads of a CTA pair: data lands in the ISSUING CTA's shared memory, the complete_tx goes to `bar_cluster_addr`,
// which may be the peer (leader) CTA's mbarrier (shared::cluster address from mapa_shared).
DDL_DEVICE void tma_load_2d_pair(uint32_t dst_smem, const CUtensorMap* m, int x, int y, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(bar_cluster_addr) : "memory");
}
DDL_DEVICE void tma_load_4d_pair(uint32_t dst_smem, const CUtensorMap* m, int c0, int c1, int c2, int c3,
                                 uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_cluster_addr)
      : "memory");
}
DDL_DEVICE void tma_load_5d_pair(uint32_t dst_smem, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4,
                                 uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
         "r"(bar_cluster_addr)
      : "memory");
}
DDL_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}

DDL_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- cp.async (16 B, zero-fill when !valid) ---------------------------------------------------
DDL_DEVICE void cp_async_16(uint32_t dst_smem, const void* src, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
DDL_DEVICE void cp_async_8(uint32_t dst_smem, const void* src, bool valid) {
  const uint32_t sz = valid ? 8u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
DDL_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
DDL_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// ---- TMA -------------------------------------------------------------------------------------
DDL_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
DDL_DEVICE void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------
template <int COLS>
DDL_DEVICE void tmem_alloc(uint32_t* slot_in_smem) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(slot_in_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
DDL_DEVICE void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}
DDL_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DDL_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread i of the warp receives lane (base_lane + i), 32 columns.
DDL_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
DDL_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ----------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128B swizzle (layout_type = 2), version 1 (Blackwell).
DDL_DEVICE uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);              // start address  [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;         // leading offset [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;         // stride offset  [32,46)
  d |= 1ull << 46;                                                      // version = 1
  d |= 2ull << 61;                                                      // SWIZZLE_128B
  return d;
}
// The start-address field (16-byte units, bits [0,14)) sits in the descriptor's low word and never carries out of it for
// addresses below 256 KB, so stepping through stages / K slices is ONE 32-bit add on a pre-built descriptor.  The MMA
// thread issues 4 MMAs per k-block on its own: rebuilding two 64-bit descriptors per MMA (~25 dependent integer
// instructions each) made that single thread, not the tensor core, the pacing resource of a one-CTA-per-SM kernel.
DDL_DEVICE uint64_t desc_join(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, M x N tile, operand major-ness (0 = K, 1 = MN).
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
DDL_DEVICE void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// fp8 operands (kind::f8f6f4, dense, fp32 accumulate): format codes 0 = e4m3, 1 = e5m2; K = 32 elements (32 bytes) per MMA
__host__ __device__ constexpr uint32_t idesc_f8(int m, int n, int a_fmt, int b_fmt, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(a_fmt) << 7) | (static_cast<uint32_t>(b_fmt) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
DDL_DEVICE void umma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
DDL_DEVICE void umma_f8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// ---- MX block-scaled fp8 (kind::mxf8f6f4.block_scale, scale vector = 32 elements, UE8M0 scales in TMEM) ------------
// instruction descriptor of the block-scaled kinds: no C format field; scale format bit 23 (1 = E8M0); a_sf_id / b_sf_id
// select WHICH of the 4 bytes of a scale column this MMA's K sub-block uses.
__host__ __device__ constexpr uint32_t idesc_mxf8(int m, int n, int a_fmt, int b_fmt, int a_sf_id, int b_sf_id) {
  return (static_cast<uint32_t>(b_sf_id) << 4) | (static_cast<uint32_t>(a_fmt) << 7) | (static_cast<uint32_t>(b_fmt) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (1u << 23) | (static_cast<uint32_t>(m >> 4) << 24) |
         (static_cast<uint32_t>(a_sf_id) << 29);
}
DDL_DEVICE void umma_mxf8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate,
                          uint32_t tmem_sfa, uint32_t tmem_sfb) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(tmem_sfa), "r"(tmem_sfb) : "memory");
}
// shared memory -> TMEM copy of one 512-byte scale atom: 32 rows x 16 bytes, broadcast to the four 32-lane quarters
DDL_DEVICE void tmem_cp_32x128b_warpx4(uint32_t dst_tmem, uint64_t src_desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" :: "r"(dst_tmem), "l"(src_desc) : "memory");
}
// un-swizzled (interleave) K-major descriptor of a dense [rows][16 B] block: 8-row core matrices 128 bytes apart
DDL_DEVICE uint64_t smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// plain (non-tensor) bulk copy global -> shared, completing on an mbarrier
DDL_DEVICE void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Arrive on `bar` when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
DDL_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}

// vectorised fp32 reduction into global memory (wgrad split-K)
DDL_DEVICE void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

}  // namespace tc
}  // namespace ddl
