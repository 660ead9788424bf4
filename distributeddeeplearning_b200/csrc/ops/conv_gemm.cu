// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 in, fp32 accumulate).
//
// Replaces the cuDNN conv fwd / bwd-data / bwd-filter and cuBLAS FC calls of the reference's
// training step (N4, N5).  All three passes are GEMMs on NHWC data:
//
//   fwd    Y[m][co]  = sum_k  im2col(X)[m][k]  * W[co][k]          k = (r, s, ci)
//   dgrad  dX[m][ci] = sum_k  im2colT(dY)[m][k] * W[co][(r,s,ci)]   k = (r, s, co)   (B is MN-major)
//   wgrad  dW[co][k] = sum_m  dY[m][co]        * im2col(X)[m][k]    (A and B are MN-major, split-K)
//
// One CTA = 6 warps:
//   warps 0-3  epilogue (TMEM -> registers -> smem -> coalesced global stores; fused BN statistics /
//              bias / ReLU / residual add).  In the gather modes (strided convs, stem) the same warps
//              first act as im2col producers: each thread owns one 128-byte row of the operand tile and
//              fills it with cp.async (zero-fill = padding / tails) in the 128B-swizzled image tcgen05 reads.
//   warp 4     TMA producer: weights always; the activation operand too whenever it is a plain matrix
//              (2-D map) or a stride-1 convolution window (ONE 4-D box per filter tap — the hardware does
//              the im2col, out-of-bounds coordinates supply the zero padding)
//   warp 5     TMEM allocator + the single thread that issues tcgen05.mma and tcgen05.commit
// Stages are handed over with mbarriers (full/empty ring + one accumulator-ready barrier).  The pipeline
// depth is chosen per launch (short-K layers take less shared memory, so more CTAs are resident per SM
// and one CTA's epilogue overlaps the others' loads).
#include "conv_gemm.cuh"
#include "tc_utils.cuh"
#include "../launch.h"

#include <dlfcn.h>

namespace ddl {

int g_pdl = 1;        // launch.h: opted-in kernels are launched with programmatic stream serialization

using namespace tc;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int kATileBytes = kBlockM * 128;        // 16 KB
constexpr int kProducerThreads = 128;              // warps 0-3: gather producers (and epilogue group 0)
constexpr int kEpiGroups = 2;                      // epilogue warp groups: group g owns columns [g*N/2, (g+1)*N/2)
constexpr int kEpiThreads = 128 * kEpiGroups;      // warps 0-3 (+ warps 6-9)
constexpr int kThreads = 192 + 128 * (kEpiGroups - 1);
constexpr int kLag = 2;                           // cp.async groups kept in flight per producer
constexpr int kMaxStages = 4;

template <int BLOCK_N>
struct FwdCfg {
  static constexpr int kBTileBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kPitch = BLOCK_N * 2 + 16;                 // epilogue staging row pitch (bytes)
  static constexpr int kEpiBytes = kBlockM * kPitch;
  static constexpr int kMaxStagesN = (BLOCK_N <= 64) ? 4 : 3;
  static constexpr int smem_bytes(int stages) {
    const int pipe = stages * kStageBytes;
    return (pipe > kEpiBytes ? pipe : kEpiBytes) + 256 /*barriers*/ + 8192 /*stats scratch*/ + 1024 /*alignment*/;
  }
};

DDL_DEVICE void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

DDL_DEVICE void tma_load_4d(uint32_t dst_smem, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

DDL_DEVICE void tma_load_5d(uint32_t dst_smem, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4,
                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      :: "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
         "r"(smem_u32(bar))
      : "memory");
}

constexpr bool mode_a_tma(int mode) {
  return mode == kConvGemm || mode == kConvTileFwd || mode == kConvTileDgrad || mode == kConvGemmDgrad ||
         mode == kConvStemTma;
}
constexpr bool mode_b_mn(int mode) { return mode == kConvDgrad || mode == kConvTileDgrad || mode == kConvGemmDgrad; }
constexpr bool mode_tile(int mode) { return mode == kConvTileFwd || mode == kConvTileDgrad || mode == kConvStemTma; }
constexpr bool mode_stem(int mode) { return mode == kConvStem || mode == kConvStemTma; }


// ---------------------------------------------------------------------------------------------
// epilogue building blocks (shared by the one-tile and the persistent kernel)
// ---------------------------------------------------------------------------------------------
// All staging traffic uses 32-bit shared-window addresses (ld/st.shared): generic LD/ST through the aligned
// dynamic-smem pointer cost 64-bit address arithmetic per access and showed up as ~15 % of the epilogue's
// instruction stream in the ncu source view (profiles/ncu_summary.md).
DDL_DEVICE void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
DDL_DEVICE uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
DDL_DEVICE void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
DDL_DEVICE uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
DDL_DEVICE float lds_f32(uint32_t addr) { return __uint_as_float(lds32(addr)); }

// Attribution experiments (tools/epi_probe.py): bits 12-14 of the variant word switch epilogue phases OFF.  The results
// are then wrong by construction; production launches never set them (native.py builds variant words from bits 0-8).
constexpr int kDbgNoStats = 1 << 12;   // no statistics pass (and no atomics)
constexpr int kDbgNoStore = 1 << 13;   // no global stores
constexpr int kDbgNoDrain = 1 << 14;   // no TMEM -> staging copy

// TMEM accumulator columns [c_begin, c_end) of this thread's row -> (bias, ReLU) -> bf16 -> staging row.
// `stg_row`: shared address of the row's staging line; `zero_row`: the row lies outside the image (tile modes).
template <bool BIAS_RELU>
DDL_DEVICE void epi_tmem_to_stage(uint32_t taddr_row, uint32_t stg_row, int c_begin, int c_end, const ConvArgs& a,
                                  int n0, bool zero_row, float deq = 1.0f) {
  if (a.variant & kDbgNoDrain) return;
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(taddr_row + c0, v);
    tmem_ld_wait();
    uint32_t packed[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x0 = __uint_as_float(v[2 * j]) * deq, x1 = __uint_as_float(v[2 * j + 1]) * deq;
      if (BIAS_RELU) {
        if (a.bias) {
          const int cb = n0 + c0 + 2 * j;
          if (cb < a.n_valid) x0 += a.bias[cb];
          if (cb + 1 < a.n_valid) x1 += a.bias[cb + 1];
        }
        if (a.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
      }
      packed[j] = zero_row ? 0u : pack_bf16x2(x0, x1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(stg_row + c0 * 2 + j * 16, packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
  }
}

// ---- packed fp32 pairs (FADD2 / FFMA2: one issue slot for two lanes of statistics math) -----------------------------
DDL_DEVICE uint64_t f2_pack(float lo, float hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
DDL_DEVICE void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DDL_DEVICE uint64_t f2_from_bf16x2(uint32_t u) {            // {bf16 lo, bf16 hi} -> {fp32, fp32}: two ALU ops
  uint64_t d;
  const uint32_t lo = u << 16, hi = u & 0xffff0000u;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
DDL_DEVICE uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
DDL_DEVICE uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// Per-thread partial BN statistics of 8 consecutive output channels (4 packed pairs each): they live in registers
// ACROSS the tiles a persistent CTA walks (as long as the tiles share their first channel n0) and are flushed — warp
// shuffles, one shared-memory fold, one atomic per channel and statistic — only when n0 changes or the CTA is done.
struct StatAcc {
  uint64_t s[4], q[4];
  int n0;                       // first channel of the tile column these sums belong to; -1 = empty
  DDL_DEVICE void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i] = 0ull; q[i] = 0ull; }
    n0 = -1;
  }
};

// Fold the staged bf16 tile (row pitch BLOCK_N*2+16) into `acc`.  Thread = (column group of 8, row slice); invalid rows
// hold zeros.  Entered by all kEpiThreads epilogue threads after the staging writes were made visible (named barrier 1).
template <int BLOCK_N, bool TILE>
DDL_DEVICE void epi_stats_accum(uint32_t stg, int etid, const ConvArgs& a, int n0, int m0, StatAcc& acc) {
  constexpr int kPitch = BLOCK_N * 2 + 16;
  constexpr int kColGroups = BLOCK_N / 8;                 // 16 (N=128) or 8 (N=64)
  constexpr int kSlices = kEpiThreads / kColGroups;       // row slices
  constexpr int kRowsPer = kBlockM / kSlices;             // rows per thread
  if (a.variant & kDbgNoStats) return;
  const int cg = etid % kColGroups;
  const int sl = etid / kColGroups;
  const uint32_t p0 = stg + (sl * kRowsPer) * kPitch + cg * 16;
  acc.n0 = n0;
  if (a.bnr_y == nullptr) {
#pragma unroll
    for (int r = 0; r < kRowsPer; ++r) {
      const uint4 u = lds128(p0 + r * kPitch);
      uint64_t f;
      f = f2_from_bf16x2(u.x); acc.s[0] = f2_add(acc.s[0], f); acc.q[0] = f2_fma(f, f, acc.q[0]);
      f = f2_from_bf16x2(u.y); acc.s[1] = f2_add(acc.s[1], f); acc.q[1] = f2_fma(f, f, acc.q[1]);
      f = f2_from_bf16x2(u.z); acc.s[2] = f2_add(acc.s[2], f); acc.q[2] = f2_fma(f, f, acc.q[2]);
      f = f2_from_bf16x2(u.w); acc.s[3] = f2_add(acc.s[3], f); acc.q[3] = f2_fma(f, f, acc.q[3]);
    }
    return;
  }
  // fused BN-backward reduction (see ConvArgs::bnr_y): S1 = sum dm, S2 = sum dm * y with the ReLU mask of the BN
  // recomputed from y.  The y tile is read with the same coalesced 16-byte accesses the stores use.
  const int c0 = n0 + cg * 8;
  if (c0 >= a.n_valid) return;
  uint64_t sc[4], sh[4];
#pragma unroll
  for (int qd = 0; qd < 2; ++qd) {
    const float4 g4 = reinterpret_cast<const float4*>(a.bnr_gamma + c0)[qd];
    const float4 b4 = reinterpret_cast<const float4*>(a.bnr_beta + c0)[qd];
    const float4 m4 = reinterpret_cast<const float4*>(a.bnr_mean + c0)[qd];
    const float4 i4 = reinterpret_cast<const float4*>(a.bnr_invstd + c0)[qd];
    const float s0 = g4.x * i4.x, s1 = g4.y * i4.y, s2 = g4.z * i4.z, s3 = g4.w * i4.w;
    sc[2 * qd] = f2_pack(s0, s1); sh[2 * qd] = f2_pack(b4.x - m4.x * s0, b4.y - m4.y * s1);
    sc[2 * qd + 1] = f2_pack(s2, s3); sh[2 * qd + 1] = f2_pack(b4.z - m4.z * s2, b4.w - m4.w * s3);
  }
#pragma unroll 2
  for (int r = 0; r < kRowsPer; ++r) {
    const int row = sl * kRowsPer + r;
    int m;
    if (TILE) m = static_cast<int>(lds32(stg + row * kPitch + BLOCK_N * 2));
    else m = (m0 + row) < a.M ? (m0 + row) : -1;
    if (m < 0) continue;
    const uint4 u = lds128(p0 + r * kPitch);
    const uint4 yv = *reinterpret_cast<const uint4*>(a.bnr_y + static_cast<size_t>(m) * a.ldc + c0);
    const uint32_t du[4] = {u.x, u.y, u.z, u.w}, yu[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t y2 = f2_from_bf16x2(yu[j]);
      float z0, z1, d0, d1;
      f2_unpack(f2_fma(y2, sc[j], sh[j]), z0, z1);
      f2_unpack(f2_from_bf16x2(du[j]), d0, d1);
      const uint64_t dm = f2_pack(z0 > 0.f ? d0 : 0.f, z1 > 0.f ? d1 : 0.f);
      acc.s[j] = f2_add(acc.s[j], dm);
      acc.q[j] = f2_fma(dm, y2, acc.q[j]);
    }
  }
}

// Publish `acc` (sums of tile column acc.n0) and clear it.  Contains one named barrier: must be reached by all
// kEpiThreads epilogue threads together; `red` = [epi warps][2][BLOCK_N] floats of scratch.  A later flush may rewrite
// the scratch only after another barrier of the same group (every caller has one per tile / chunk in between).
template <int BLOCK_N>
DDL_DEVICE void epi_stats_flush(uint32_t red, int etid, int ew, const ConvArgs& a, StatAcc& acc) {
  constexpr int kColGroups = BLOCK_N / 8;
  if (a.variant & kDbgNoStats) { acc.clear(); return; }
  const int cg = etid % kColGroups;
  const int n0 = acc.n0;
  // threads with the same column group inside a warp: lanes cg, cg+kColGroups, ... -> xor shuffles
#pragma unroll
  for (int off = kColGroups; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc.s[i] = f2_add(acc.s[i], __shfl_xor_sync(0xffffffffu, acc.s[i], off));
      acc.q[i] = f2_add(acc.q[i], __shfl_xor_sync(0xffffffffu, acc.q[i], off));
    }
  }
  // cross-warp fold in shared memory ([epi warp][2][BLOCK_N] floats), then ONE atomic per channel and statistic
  if ((threadIdx.x & 31) < kColGroups) {
    const uint32_t rb = red + ((ew * 2) * BLOCK_N + cg * 8) * 4;
    sts128(rb, static_cast<uint32_t>(acc.s[0]), static_cast<uint32_t>(acc.s[0] >> 32), static_cast<uint32_t>(acc.s[1]),
           static_cast<uint32_t>(acc.s[1] >> 32));
    sts128(rb + 16, static_cast<uint32_t>(acc.s[2]), static_cast<uint32_t>(acc.s[2] >> 32),
           static_cast<uint32_t>(acc.s[3]), static_cast<uint32_t>(acc.s[3] >> 32));
    sts128(rb + BLOCK_N * 4, static_cast<uint32_t>(acc.q[0]), static_cast<uint32_t>(acc.q[0] >> 32),
           static_cast<uint32_t>(acc.q[1]), static_cast<uint32_t>(acc.q[1] >> 32));
    sts128(rb + BLOCK_N * 4 + 16, static_cast<uint32_t>(acc.q[2]), static_cast<uint32_t>(acc.q[2] >> 32),
           static_cast<uint32_t>(acc.q[3]), static_cast<uint32_t>(acc.q[3] >> 32));
  }
  acc.clear();
  named_bar_sync(1, kEpiThreads);
  if (a.bnr_y == nullptr) {
    for (int c = etid; c < 2 * BLOCK_N; c += kEpiThreads) {
      const int which = c / BLOCK_N, col = c - which * BLOCK_N;
      float v = 0.f;
#pragma unroll
      for (int wq = 0; wq < 4 * kEpiGroups; ++wq) v += lds_f32(red + ((wq * 2 + which) * BLOCK_N + col) * 4);
      if (n0 + col < a.n_valid) atomicAdd((which ? a.sumsq : a.sum) + n0 + col, v);
    }
  } else {
    for (int col = etid; col < BLOCK_N; col += kEpiThreads) {      // dbeta += S1, dgamma += invstd * (S2 - mean * S1)
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int wq = 0; wq < 4 * kEpiGroups; ++wq) {
        t1 += lds_f32(red + ((wq * 2 + 0) * BLOCK_N + col) * 4);
        t2 += lds_f32(red + ((wq * 2 + 1) * BLOCK_N + col) * 4);
      }
      const int ch = n0 + col;
      if (ch < a.n_valid) {
        atomicAdd(a.sum + ch, t1);
        atomicAdd(a.sumsq + ch, a.bnr_invstd[ch] * (t2 - a.bnr_mean[ch] * t1));
      }
    }
  }
}

// The coalesced 16-byte global stores of the staged tile (optional `+= add`, optional zero-fill of the three stride-2
// siblings): thread = (16-byte column chunk, row phase); a tile row is BLOCK_N*2 contiguous bytes.
template <int BLOCK_N, bool TILE>
DDL_DEVICE void epi_store(uint32_t stg, int etid, const ConvArgs& a, int n0, int m0) {
  constexpr int kPitch = BLOCK_N * 2 + 16;
  constexpr int kVecPerRow = BLOCK_N / 8;
  constexpr int kRowStep = kEpiThreads / kVecPerRow;
  constexpr int kIters = kBlockM / kRowStep;
  const int ch = etid % kVecPerRow;
  const int r0 = etid / kVecPerRow;
  if (n0 + ch * 8 >= a.ldc) return;                       // channel padding of the last N tile (ldc = row width)
  if (a.variant & kDbgNoStore) return;
  uint32_t sp = stg + r0 * kPitch + ch * 16;
  const size_t coff = static_cast<size_t>(n0 + ch * 8);
  if (a.add == nullptr && !(TILE && a.zfill)) {
    // plain stores (every forward launch): nothing but the shared load, the row's address and the store
    if (!TILE) {
      const int rows = a.M - m0;                            // valid rows of this tile (>= kBlockM except in the last one)
      uint4* gp = reinterpret_cast<uint4*>(a.out + static_cast<size_t>(m0 + r0) * a.ldc + coff);
      const size_t gstep = static_cast<size_t>(kRowStep) * a.ldc / 8;      // ldc % 8 == 0 (16-byte rows)
      if (rows >= kBlockM) {
#pragma unroll
        for (int i = 0; i < kIters; ++i) gp[i * gstep] = lds128(sp + i * kRowStep * kPitch);
      } else {
#pragma unroll 1
        for (int r = r0; r < rows; r += kRowStep, sp += kRowStep * kPitch, gp += gstep) *gp = lds128(sp);
      }
    } else {
      __nv_bfloat16* gcol = a.out + coff;
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const int m = static_cast<int>(lds32(sp + i * kRowStep * kPitch - ch * 16 + BLOCK_N * 2));
        if (m >= 0) *reinterpret_cast<uint4*>(gcol + static_cast<size_t>(m) * a.ldc) = lds128(sp + i * kRowStep * kPitch);
      }
    }
    return;
  }
#pragma unroll 4
  for (int r = r0; r < kBlockM; r += kRowStep, sp += kRowStep * kPitch) {
    int m;
    if (TILE) m = static_cast<int>(lds32(sp - ch * 16 + BLOCK_N * 2));
    else m = (m0 + r) < a.M ? (m0 + r) : -1;
    if (m < 0) continue;
    uint4 val = lds128(sp);
    const size_t off = static_cast<size_t>(m) * a.ldc + coff;
    if (a.add) {
      uint4 o = *reinterpret_cast<const uint4*>(a.add + off);
      if (a.add_mask) {
        const uint32_t bits = a.add_mask[off >> 3];
        o.x &= ((bits & 1u) ? 0x0000ffffu : 0u) | ((bits & 2u) ? 0xffff0000u : 0u);
        o.y &= ((bits & 4u) ? 0x0000ffffu : 0u) | ((bits & 8u) ? 0xffff0000u : 0u);
        o.z &= ((bits & 16u) ? 0x0000ffffu : 0u) | ((bits & 32u) ? 0xffff0000u : 0u);
        o.w &= ((bits & 64u) ? 0x0000ffffu : 0u) | ((bits & 128u) ? 0xffff0000u : 0u);
      }
      float2 p, q;
      p = unpack_bf16x2(val.x); q = unpack_bf16x2(o.x); val.x = pack_bf16x2(p.x + q.x, p.y + q.y);
      p = unpack_bf16x2(val.y); q = unpack_bf16x2(o.y); val.y = pack_bf16x2(p.x + q.x, p.y + q.y);
      p = unpack_bf16x2(val.z); q = unpack_bf16x2(o.z); val.z = pack_bf16x2(p.x + q.x, p.y + q.y);
      p = unpack_bf16x2(val.w); q = unpack_bf16x2(o.w); val.w = pack_bf16x2(p.x + q.x, p.y + q.y);
    }
    *reinterpret_cast<uint4*>(a.out + off) = val;
    if (TILE && a.zfill) {      // even outH/outW guaranteed by the host: all three siblings exist
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(a.out + off + a.ldc) = z;
      *reinterpret_cast<uint4*>(a.out + off + static_cast<size_t>(a.outW) * a.ldc) = z;
      *reinterpret_cast<uint4*>(a.out + off + static_cast<size_t>(a.outW + 1) * a.ldc) = z;
    }
  }
}

// One-shot form (one tile per call: accumulate, publish, store) for the kernels that do not carry sums across tiles.
template <int BLOCK_N, bool STATS, bool TILE>
DDL_DEVICE void epi_stats_store(uint32_t stg, uint32_t red, int etid, int ew, const ConvArgs& a, int n0, int m0) {
  if (STATS) {
    StatAcc acc;
    acc.clear();
    epi_stats_accum<BLOCK_N, TILE>(stg, etid, a, n0, m0, acc);
    acc.n0 = n0;
    epi_stats_flush<BLOCK_N>(red, etid, ew, a, acc);
  }
  epi_store<BLOCK_N, TILE>(stg, etid, a, n0, m0);
}

// ---------------------------------------------------------------------------------------------
// fwd / dgrad / plain-GEMM / stem kernel
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N, int MODE, bool STATS>
__global__ void __launch_bounds__(kThreads, 3)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ TmaSet tmAs, ConvArgs a) {
  const CUtensorMap& tmA = tmAs.m[0];
  using Cfg = FwdCfg<BLOCK_N>;
  constexpr bool kATma = mode_a_tma(MODE);
  constexpr bool kBMn = mode_b_mn(MODE);
  constexpr bool kTile = mode_tile(MODE);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nstages = a.stages;
  const int pipe_bytes = nstages * Cfg::kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (pipe_bytes > Cfg::kEpiBytes ? pipe_bytes : Cfg::kEpiBytes));
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_full = empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * BLOCK_N;
  const int KB = a.KB;

  // ---- tile -> output rows ------------------------------------------------------------------------
  int m0 = blockIdx.y * kBlockM;          // linear modes: rows [m0, m0+128)
  int tq0 = 0, tp0 = 0, tn0 = 0;          // tile modes: box origin in the output's (w, h, n) space
  if (kTile) {
    int t = blockIdx.y;
    const int wb = t % a.tiles_w; t /= a.tiles_w;
    const int hb = t % a.tiles_h;
    const int nb = t / a.tiles_h;
    tq0 = wb * a.tw; tp0 = hb * a.th; tn0 = nb * a.tn;
    m0 = 0;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full[s], kATma ? 1u : (kProducerThreads + 1u));
      mbar_init(&empty[s], 1u);
    }
    mbar_init(acc_full, 1u);
    fence_mbar_init();
  }
  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tmB);
    if (kATma) tma_prefetch_desc(&tmA);
  }
  if (warp == 5) tmem_alloc<BLOCK_N>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();      // the next kernel in the stream may start its own preamble
  pdl_wait();         // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp < 4 || warp >= 6) {
    const int egrp = warp < 4 ? 0 : 1;                 // epilogue group
    const int qw = warp & 3;                           // TMEM lane quarter this warp may read (= warp % 4)
    const int erow = qw * 32 + (threadIdx.x & 31);     // accumulator row owned in the epilogue
    const int etid = egrp * 128 + erow;                // dense index over the epilogue threads
    const int ew = egrp * 4 + qw;                      // dense epilogue warp index
    // =============================== gather producers ===================================
    if (!kATma && warp < 4) {
      const int row = threadIdx.x;
      const int m = m0 + row;
      const bool row_ok = m < a.M;
      int img = 0, dh = 0, dw = 0;
      if (row_ok) {
        const int hw = a.dstH * a.dstW;
        img = m / hw;
        const int rem = m - img * hw;
        dh = rem / a.dstW;
        dw = rem - dh * a.dstW;
      }
      int hb, wb;
      if (MODE == kConvDgrad) { hb = dh + a.pad; wb = dw + a.pad_w; }
      else { hb = dh * a.stride - a.pad; wb = dw * a.stride - a.pad_w; }
      const __nv_bfloat16* img_base = a.src + static_cast<size_t>(img) * a.srcH * a.srcW * a.srcC;
      const uint32_t row_off = row * 128u;
      const uint32_t sw = row & 7u;
      int tap_r = 0, tap_s = 0, cc = 0;   // incremental decode of kb -> (r, s, channel chunk)
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u, 17);
        const uint32_t dst = smem_u32(smem + stage * Cfg::kStageBytes) + row_off;
        if (MODE == kConvStem) {
          // k-block = RPK filter rows x SP taps x 4 channels, SP = a.cchunks (padded taps per row)
          const int SP = a.cchunks;
          const int RPK = 16 / SP;           // 64 elements / (SP * 4)
          for (int rr = 0; rr < RPK; ++rr) {
            const int r = kb * RPK + rr;
            const int h = hb + r * a.dil;
            const bool h_ok = row_ok && r < a.R && h >= 0 && h < a.srcH;
            for (int s = 0; s < SP; ++s) {
              const int e = s * RPK + rr;     // 8-byte element index in the 128-byte row: (tap, row-in-block) order
              const int w = wb + s * a.dil;
              const bool ok = h_ok && s < a.S && w >= 0 && w < a.srcW;
              const __nv_bfloat16* src = ok ? img_base + (static_cast<size_t>(h) * a.srcW + w) * 4 : a.src;
              const uint32_t chunk = (e >> 1) ^ sw;
              cp_async_8(dst + (chunk << 4) + ((e & 1) << 3), src, ok);
            }
          }
        } else {
          bool ok;
          int sh, swd;
          if (MODE == kConvDgrad) {
            const int th = hb - tap_r * a.dil, tw = wb - tap_s * a.dil;
            if (a.stride == 1) {
              sh = th; swd = tw;
              ok = row_ok && th >= 0 && th < a.srcH && tw >= 0 && tw < a.srcW;
            } else {
              sh = th / a.stride; swd = tw / a.stride;
              ok = row_ok && th >= 0 && tw >= 0 && (th - sh * a.stride) == 0 && (tw - swd * a.stride) == 0 &&
                   sh < a.srcH && swd < a.srcW;
            }
          } else {
            sh = hb + tap_r * a.dil; swd = wb + tap_s * a.dil;
            ok = row_ok && sh >= 0 && sh < a.srcH && swd >= 0 && swd < a.srcW;
          }
          const __nv_bfloat16* src =
              ok ? img_base + (static_cast<size_t>(sh) * a.srcW + swd) * a.srcC + cc * 64 : a.src;
          const int crem = a.srcC - cc * 64;      // channels left in this tap (last k-block of a tap may be partial)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool okj = ok && j * 8 < crem;
            cp_async_16(dst + ((static_cast<uint32_t>(j) ^ sw) << 4), okj ? src + j * 8 : a.src, okj);
          }
          if (++cc == a.cchunks) { cc = 0; if (++tap_s == a.S) { tap_s = 0; ++tap_r; } }
        }
        cp_async_commit();
        if (kb >= kLag) {
          cp_async_wait<kLag>();
          fence_proxy_async_smem();
          mbar_arrive(&full[(kb - kLag) % nstages]);
        }
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (int kb = (KB > kLag ? KB - kLag : 0); kb < KB; ++kb) mbar_arrive(&full[kb % nstages]);
    }

    // =================================== epilogue ========================================
    // my accumulator row -> output row
    const int row = erow;
    int my_m;                     // linear output row index of this thread's accumulator row, or -1
    if (kTile) {
      const int wl = row % a.tw;
      const int t2 = row / a.tw;
      const int hl = t2 % a.th;
      const int nl = t2 / a.th;
      const bool ok = nl < a.tn && (tn0 + nl) < a.batch && (tp0 + hl) < a.dstH && (tq0 + wl) < a.dstW;
      my_m = ok ? ((tn0 + nl) * a.outH + (tp0 + hl) * a.out_stride + a.out_pa) * a.outW +
                      (tq0 + wl) * a.out_stride + a.out_pb
                : -1;
    } else {
      my_m = (m0 + row) < a.M ? (m0 + row) : -1;
    }
    mbar_wait(acc_full, 0, 20);
    tc_fence_after();
    // pipeline buffers are free (every MMA that read them has completed): the tile is staged over them
    const uint32_t stg = smem_u32(smem);
    const uint32_t stg_row = stg + row * Cfg::kPitch;
    constexpr int kColsPerGroup = BLOCK_N / kEpiGroups;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(qw * 32) << 16);
    const bool zero_row = kTile && my_m < 0;          // rows outside the image: exact zeros (keeps the BN sums clean)
    if (!STATS && (a.bias != nullptr || a.relu))
      epi_tmem_to_stage<true>(taddr_row, stg_row, egrp * kColsPerGroup, (egrp + 1) * kColsPerGroup, a, n0, zero_row);
    else
      epi_tmem_to_stage<false>(taddr_row, stg_row, egrp * kColsPerGroup, (egrp + 1) * kColsPerGroup, a, n0, zero_row);
    // the row's destination index rides in the pad bytes of its staging row (pitch = 2*BLOCK_N + 16)
    if (kTile && egrp == 0) sts32(stg_row + BLOCK_N * 2, static_cast<uint32_t>(my_m));
    tc_fence_before();
    named_bar_sync(1, kEpiThreads);
    epi_stats_store<BLOCK_N, STATS, kTile>(stg, smem_u32(full) + 256, etid, ew, a, n0, m0);
  } else if (warp == 4) {
    // ================================== TMA producer =====================================
    if (elect_one()) {
      int tap = 0, cc = 0, tap_r = 0, tap_s = 0;
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = kTile ? static_cast<uint32_t>(a.tw * a.th * a.tn) * 128u : kATileBytes;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u, 17);
        uint8_t* sA = smem + stage * Cfg::kStageBytes;
        const uint32_t sB = smem_u32(sA + kATileBytes);
        mbar_arrive_expect_tx(&full[stage], Cfg::kBTileBytes + (kATma ? a_bytes : 0u));
        // tap of this k-block: weight tap index + (tile modes) source box offset / source map
        int widx = tap, dh = 0, dw = 0, mapi = 0;
        if (kTile) {
          if (a.ntaps > 0) {
            widx = a.tap_widx[tap]; dh = a.tap_dh[tap]; dw = a.tap_dw[tap]; mapi = a.tap_map[tap];
          } else if (MODE == kConvTileFwd) {
            dh = tap_r * a.dil - a.pad; dw = tap_s * a.dil - a.pad_w;
          } else {
            dh = a.pad - tap_r * a.dil; dw = a.pad_w - tap_s * a.dil;
          }
        }
        if (kBMn) {
          // weights W[co][(r,s,ci)]: K rows = 64 output channels, N = input channels (contiguous)
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sB + j * 8192, &tmB, widx * a.ldc + n0 + j * 64, cc * 64, &full[stage]);
        } else {
          tma_load_2d(sB, &tmB, mode_stem(MODE) ? kb * kBlockK : widx * a.kstride + cc * kBlockK, n0, &full[stage]);
        }
        if (MODE == kConvGemm || MODE == kConvGemmDgrad) {
          tma_load_2d(smem_u32(sA), &tmA, kb * kBlockK, m0, &full[stage]);
        } else if (MODE == kConvStemTma) {
          tma_load_5d(smem_u32(sA), &tmA, 0, kb, tq0, tp0, tn0, &full[stage]);
        } else if (kTile) {
          tma_load_4d(smem_u32(sA), &tmAs.m[mapi], cc * 64, tq0 + dw, tp0 + dh, tn0, &full[stage]);
        }
        if (++cc == a.cchunks) { cc = 0; ++tap; if (++tap_s == a.S) { tap_s = 0; ++tap_r; } }
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // =================================== MMA issuer =======================================
    constexpr uint32_t idesc = idesc_bf16(kBlockM, BLOCK_N, 0, kBMn ? 1 : 0);
    const uint64_t odA0 = smem_desc_sw128(smem_u32(smem), 0, 1024);
    const uint64_t odB0 = kBMn ? smem_desc_sw128(smem_u32(smem) + kATileBytes, 8192, 1024)
                               : smem_desc_sw128(smem_u32(smem) + kATileBytes, 0, 1024);
    const uint32_t oa_lo = static_cast<uint32_t>(odA0), oa_hi = static_cast<uint32_t>(odA0 >> 32);
    const uint32_t ob_lo = static_cast<uint32_t>(odB0), ob_hi = static_cast<uint32_t>(odB0 >> 32);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < KB; ++kb) {
      mbar_wait(&full[stage], phase, 18);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t so = static_cast<uint32_t>(stage) * (Cfg::kStageBytes >> 4);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t da = desc_join(oa_lo + so + 2u * k, oa_hi);
          const uint64_t db = desc_join(ob_lo + so + (kBMn ? 128u : 2u) * k, ob_hi);
          umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (kb == KB - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<BLOCK_N>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// persistent variant for the TMA-fed modes (GEMM, TILE_FWD, TILE_DGRAD, GEMM_DGRAD)
// ---------------------------------------------------------------------------------------------
// One CTA per resident slot (2 per SM) walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The operand ring
// keeps streaming across tile boundaries, the accumulator is double-buffered in TMEM (2 x BLOCK_N columns) and the
// 8 epilogue warps drain tile i while the MMA thread already accumulates tile i+1: the per-tile fixed latency
// (TMEM alloc, barrier init, first TMA round trip, epilogue) that bounds the one-tile-per-CTA kernel on short-K
// layers is paid once per CTA instead of once per tile.
template <int BLOCK_N>
struct PersistCfg {
  static constexpr int kBTileBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kPitch = BLOCK_N * 2 + 16;
  static constexpr int kEpiBytes = ((kBlockM * kPitch + 1023) / 1024) * 1024;
  // BLOCK_N = 256: one CTA per SM (its two accumulators fill the 512 TMEM columns); a 128 x 256 tile re-uses every
  // activation k-block for twice the output columns, i.e. 25 % less L2 -> shared traffic per FLOP than 128 x 128
  static constexpr int kStages = (BLOCK_N <= 64) ? 3 : 2;
  static constexpr int kRedBytes = 4 * kEpiGroups * 2 * BLOCK_N * 4;      // [epi warps][2][BLOCK_N] floats
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 256 + kRedBytes + 1024;
};

// CLUSTER = launched as thread-block clusters of 2 CTAs that work on the SAME N tile and adjacent M tiles: each CTA
// fetches half of the weight tile and TMA-multicasts it into both CTAs' shared memory, so the B operand crosses
// L2 -> SM once per pair (operand traffic per 128 x 128 tile: 32 KB -> 24 KB).  A stage may be refilled only when
// BOTH CTAs' MMAs have consumed it: `empty` counts two arrivals, delivered by a multicast tcgen05.commit.
// CL: 0 = independent CTAs, 1 = the multicast pairs described above, 2 = CTA pairs sharing ONE MMA (cta_group::2,
// M = 256): each CTA loads its 128 rows of A and only HALF of the weight tile; the leader CTA issues the MMAs, which
// read both SMs' shared memory and write both SMs' TMEM.  Bytes entering each SM per 128 x 128 outputs: 32 KB -> 24 KB.
//   barriers (CL = 2): leader.full[s] = expect_tx of BOTH CTAs' operand bytes (the peer's TMA loads complete on the
//   leader's barrier: cp.async.bulk.tensor.cta_group::2); empty[s] / acc_full[b] in both CTAs = the leader's multicast
//   commit; leader.acc_empty[b] = both epilogues (the peer arrives remotely).
template <int BLOCK_N, int MODE, bool STATS, int CL>
__global__ void __launch_bounds__(kThreads, 2)
conv_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBh,
                            const __grid_constant__ TmaSet tmAs, ConvArgs a, int n_tiles, int m_tiles) {
  using Cfg = PersistCfg<BLOCK_N>;
  constexpr bool kBMn = mode_b_mn(MODE);
  constexpr bool kTile = mode_tile(MODE);
  constexpr int kStages = Cfg::kStages;
  constexpr bool CLUSTER = CL != 0;
  constexpr bool PAIR = CL == 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + kStages * Cfg::kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + Cfg::kEpiBytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_full = empty + kMaxStages;      // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);

  const int warp = threadIdx.x >> 5;
  const int KB = a.KB;
  // work items: tiles (t -> n tile fastest), or with CLUSTER pairs of M tiles: item -> (n tile, M tiles 2j and 2j+1)
  const int crank = CLUSTER ? static_cast<int>(cluster_ctarank()) : 0;
  const int m_items = CLUSTER ? (m_tiles + 1) / 2 : m_tiles;
  const int total = n_tiles * m_items;
  const int item0 = CLUSTER ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_step = CLUSTER ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1u);
      mbar_init(&empty[s], CL == 1 ? 2u : 1u);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1u); mbar_init(&acc_empty[b], PAIR ? 2u : 1u); }
    fence_mbar_init();
  }
  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmAs.m[0]);
  }
  if (warp == 5) {
    if (PAIR) tmem_alloc_pair<2 * BLOCK_N>(tmem_slot);
    else tmem_alloc<2 * BLOCK_N>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();      // the peer's mbarriers exist before any multicast can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();         // preamble done; from here on the kernel reads and writes global memory

  // returns false for the padding tile of an odd M-tile count (its CTA still runs loads and MMAs in lock-step with
  // the peer, on the clamped last tile, but writes nothing)
  auto tile_origin = [&](int t, int& n0, int& m0, int& tq0, int& tp0, int& tn0) -> bool {
    const int nt = t % n_tiles;
    int mt = t / n_tiles;
    bool valid = true;
    if (CLUSTER) {
      mt = 2 * mt + crank;
      valid = mt < m_tiles;
      if (!valid) mt = m_tiles - 1;
    }
    n0 = nt * BLOCK_N;
    m0 = mt * kBlockM;
    tq0 = tp0 = tn0 = 0;
    if (kTile) {
      const int wb = mt % a.tiles_w; mt /= a.tiles_w;
      const int hb = mt % a.tiles_h;
      const int nb = mt / a.tiles_h;
      tq0 = wb * a.tw; tp0 = hb * a.th; tn0 = nb * a.tn;
      m0 = 0;
    }
    return valid;
  };

  if (warp < 4 || warp >= 6) {
    // =================================== epilogue warps =======================================
    const int egrp = warp < 4 ? 0 : 1;
    const int qw = warp & 3;
    const int row = qw * 32 + (threadIdx.x & 31);
    const int etid = egrp * 128 + row;
    const int ew = egrp * 4 + qw;
    constexpr int kColsPerGroup = BLOCK_N / kEpiGroups;
    const uint32_t stg_u32 = smem_u32(stg);
    const uint32_t red_u32 = smem_u32(red);
    int it = 0;
    // BN statistics ride in registers across tiles: t advances by gridDim.x, so whenever the grid is a multiple of the
    // number of N tiles (2 x 148 CTAs: 1, 2, 4 or 8 tiles) a CTA stays in ONE tile column and publishes its sums once
    StatAcc acc;
    acc.clear();
    for (int t = item0; t < total; t += item_step, ++it) {
      int n0, m0, tq0, tp0, tn0;
      const bool valid = tile_origin(t, n0, m0, tq0, tp0, tn0);
      if (CLUSTER && !valid) {          // padding tile: keep the accumulator hand-shake going, write nothing
        const int pbuf = it & 1;
        mbar_wait(&acc_full[pbuf], (it >> 1) & 1, 36);
        tc_fence_after();
        tc_fence_before();
        named_bar_sync(1, kEpiThreads);
        if (etid == 0) {
          if (PAIR && crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[pbuf]), 0));
          else mbar_arrive(&acc_empty[pbuf]);
        }
        continue;
      }
      int my_m;
      if (kTile) {
        const int wl = row % a.tw;
        const int t2 = row / a.tw;
        const int hl = t2 % a.th;
        const int nl = t2 / a.th;
        const bool ok = nl < a.tn && (tn0 + nl) < a.batch && (tp0 + hl) < a.dstH && (tq0 + wl) < a.dstW;
        my_m = ok ? ((tn0 + nl) * a.outH + (tp0 + hl) * a.out_stride + a.out_pa) * a.outW +
                        (tq0 + wl) * a.out_stride + a.out_pb
                  : -1;
      } else {
        my_m = (m0 + row) < a.M ? (m0 + row) : -1;
      }
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1, 36);
      tc_fence_after();
      const uint32_t stg_row = stg_u32 + row * Cfg::kPitch;
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + buf * BLOCK_N;
      const bool zero_row = kTile && my_m < 0;
      if (!STATS && (a.bias != nullptr || a.relu))
        epi_tmem_to_stage<true>(taddr_row, stg_row, egrp * kColsPerGroup, (egrp + 1) * kColsPerGroup, a, n0, zero_row);
      else
        epi_tmem_to_stage<false>(taddr_row, stg_row, egrp * kColsPerGroup, (egrp + 1) * kColsPerGroup, a, n0, zero_row);
      if (kTile && egrp == 0) sts32(stg_row + BLOCK_N * 2, static_cast<uint32_t>(my_m));
      tc_fence_before();
      named_bar_sync(1, kEpiThreads);
      if (etid == 0) {                                  // every epilogue thread has finished reading this TMEM buffer
        if (PAIR && crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));   // the leader issues the MMAs
        else mbar_arrive(&acc_empty[buf]);
      }
      if (STATS) {
        if (acc.n0 >= 0 && acc.n0 != n0) epi_stats_flush<BLOCK_N>(red_u32, etid, ew, a, acc);   // CTA-uniform branch
        epi_stats_accum<BLOCK_N, kTile>(stg_u32, etid, a, n0, m0, acc);
      }
      epi_store<BLOCK_N, kTile>(stg_u32, etid, a, n0, m0);
      named_bar_sync(1, kEpiThreads);      // staging may be overwritten by the next tile
    }
    if (STATS && acc.n0 >= 0) epi_stats_flush<BLOCK_N>(red_u32, etid, ew, a, acc);
  } else if (warp == 4) {
    // ===================================== TMA producer =======================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = kTile ? static_cast<uint32_t>(a.tw * a.th * a.tn) * 128u : kATileBytes;
      for (int t = item0; t < total; t += item_step) {
        int n0, m0, tq0, tp0, tn0;
        tile_origin(t, n0, m0, tq0, tp0, tn0);
        if (t + item_step >= total) pdl_trigger_last_tile();
        int tap = 0, cc = 0, tap_r = 0, tap_s = 0;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, 33);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          const uint32_t sB = smem_u32(sA + kATileBytes);
          // CTA pair: every load of BOTH CTAs completes on the LEADER's barrier (it alone waits for operands), which
          // therefore expects the bytes of the whole pair; the peer CTA arrives nowhere.
          const uint32_t pair_bar = PAIR ? mapa_shared(smem_u32(&full[stage]), 0) : 0u;
          if (!PAIR) mbar_arrive_expect_tx(&full[stage], Cfg::kBTileBytes + a_bytes);
          else if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * (Cfg::kBTileBytes / 2 + a_bytes));
          int widx = tap, dh = 0, dw = 0, mapi = 0;
          if (kTile) {
            if (a.ntaps > 0) {
              widx = a.tap_widx[tap]; dh = a.tap_dh[tap]; dw = a.tap_dw[tap]; mapi = a.tap_map[tap];
            } else if (MODE == kConvTileFwd) {
              dh = tap_r * a.dil - a.pad; dw = tap_s * a.dil - a.pad_w;
            } else {
              dh = a.pad - tap_r * a.dil; dw = a.pad_w - tap_s * a.dil;
            }
          }
          if (PAIR) {
            // my half of the weight tile into MY shared memory only: the pair MMA reads the other half from the peer
            if (kBMn) tma_load_2d_pair(sB, &tmB, widx * a.ldc + n0 + crank * (BLOCK_N / 2), cc * 64, pair_bar);
            else tma_load_2d_pair(sB, &tmBh, mode_stem(MODE) ? kb * kBlockK : widx * a.kstride + cc * kBlockK,
                                  n0 + crank * (BLOCK_N / 2), pair_bar);
            if (MODE == kConvStemTma) tma_load_5d_pair(smem_u32(sA), &tmAs.m[0], 0, kb, tq0, tp0, tn0, pair_bar);
            else if (kTile) tma_load_4d_pair(smem_u32(sA), &tmAs.m[mapi], cc * 64, tq0 + dw, tp0 + dh, tn0, pair_bar);
            else tma_load_2d_pair(smem_u32(sA), &tmAs.m[0], kb * kBlockK, m0, pair_bar);
          } else {
            if (CLUSTER) {
              // my half of the weight tile, delivered to both CTAs of the pair (each CTA's `full` barrier sees both halves)
              if (kBMn) {       // BLOCK_N = 128: the two 64-column boxes are the halves
                tma_load_2d_multicast(sB + crank * 8192, &tmB, widx * a.ldc + n0 + crank * 64, cc * 64, &full[stage], 0x3);
              } else {          // K-major: rows [crank * BLOCK_N/2, +BLOCK_N/2) through the half-height box map
                tma_load_2d_multicast(sB + crank * (BLOCK_N / 2) * 128, &tmBh,
                                      mode_stem(MODE) ? kb * kBlockK : widx * a.kstride + cc * kBlockK,
                                      n0 + crank * (BLOCK_N / 2), &full[stage], 0x3);
              }
            } else if (kBMn) {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_2d(sB + j * 8192, &tmB, widx * a.ldc + n0 + j * 64, cc * 64, &full[stage]);
            } else {
              // the weight map's box holds min(BLOCK_N, 128) rows
              constexpr int kBoxRows = BLOCK_N < 128 ? BLOCK_N : 128;
#pragma unroll
              for (int h = 0; h < BLOCK_N / kBoxRows; ++h)
                tma_load_2d(sB + h * kBoxRows * 128, &tmB, mode_stem(MODE) ? kb * kBlockK : widx * a.kstride + cc * kBlockK,
                            n0 + h * kBoxRows, &full[stage]);
            }
            if (MODE == kConvStemTma) tma_load_5d(smem_u32(sA), &tmAs.m[0], 0, kb, tq0, tp0, tn0, &full[stage]);
            else if (kTile) tma_load_4d(smem_u32(sA), &tmAs.m[mapi], cc * 64, tq0 + dw, tp0 + dh, tn0, &full[stage]);
            else tma_load_2d(smem_u32(sA), &tmAs.m[0], kb * kBlockK, m0, &full[stage]);
          }
          if (++cc == a.cchunks) { cc = 0; ++tap; if (++tap_s == a.S) { tap_s = 0; ++tap_r; } }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ====================================== MMA issuer ========================================
    constexpr uint32_t idesc = idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, BLOCK_N, 0, kBMn ? 1 : 0);
    // stage-0 descriptors built once; stage / K-slice steps are 32-bit adds on the low word (see desc_join)
    const uint64_t pdA0 = smem_desc_sw128(smem_u32(smem), 0, 1024);
    const uint64_t pdB0 = kBMn ? smem_desc_sw128(smem_u32(smem) + kATileBytes, 8192, 1024)
                               : smem_desc_sw128(smem_u32(smem) + kATileBytes, 0, 1024);
    const uint32_t pa_lo = static_cast<uint32_t>(pdA0), pa_hi = static_cast<uint32_t>(pdA0 >> 32);
    const uint32_t pb_lo = static_cast<uint32_t>(pdB0), pb_hi = static_cast<uint32_t>(pdB0 >> 32);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (PAIR && crank != 0) {
      // peer CTA of a pair: the leader issues the MMAs for both (its operands' arrival is signalled to the leader directly)
    } else {
      for (int t = item0; t < total; t += item_step, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1u, 35);     // epilogue(s) have drained this TMEM buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase, 34);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t so = static_cast<uint32_t>(stage) * (Cfg::kStageBytes >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint64_t da = desc_join(pa_lo + so + 2u * k, pa_hi);
              const uint64_t db = desc_join(pb_lo + so + (kBMn ? 128u : 2u) * k, pb_hi);
              if (PAIR) umma_bf16_pair(d_tmem, da, db, idesc, (kb | k) != 0);
              else umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
            }
            if (PAIR) {
              umma_commit_pair_multicast(&empty[stage], 0x3);          // frees the stage in both CTAs
              if (kb == KB - 1) umma_commit_pair_multicast(&acc_full[buf], 0x3);
            } else {
              if (CLUSTER) umma_commit_multicast(&empty[stage], 0x3);  // frees the stage in BOTH CTAs of the pair
              else umma_commit(&empty[stage]);
              if (kb == KB - 1) umma_commit(&acc_full[buf]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();      // no CTA leaves while its peer may still multicast into it / signal its barriers
  if (warp == 5) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<2 * BLOCK_N>(tmem_base);
    else tmem_dealloc<2 * BLOCK_N>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// deep-ring persistent kernel: ONE CTA per SM (or one cta_group::2 CTA pair per SM pair)
// ---------------------------------------------------------------------------------------------
// What bounds the long-K layers is operand delivery, L2 -> shared memory: a 128 x 128 x 64 step needs 32 KB per 256
// tensor-pipe cycles = 128 B/cycle/SM, while the chip sustains ~45 B/cycle/SM out of L2 (cuBLAS's own bf16 peak with
// 256 x 256 pair tiles sits right at that limit).  This variant therefore
//   * spends the whole SM on one CTA: the ring is as deep as 227 KB allow (5 x 32 KB for 128 x 128, 3 x 48 KB for
//     128 x 256, 5 x 32 KB for a 256 x 256 pair tile), so TMA round trips (~1-1.5 k cycles under load) stay covered;
//   * supports 256-wide tiles (A is re-used for twice the columns: 96 B/cycle) and CTA pairs with ONE M = 256 MMA
//     (tcgen05.mma.cta_group::2: each SM loads its 128 rows of A and HALF of the weight tile: 64 B/cycle at N = 256);
//   * drains the accumulator in 64-column chunks through two small staging buffers (2 x 18 KB instead of one
//     34-68 KB tile image), which is what frees the shared memory for the ring.
// TMEM: 2 x BLOCK_N columns (double-buffered accumulator; N = 256 uses all 512 columns — fine with one CTA per SM).
template <int BLOCK_N, bool PAIR>
struct DeepCfg {
  static constexpr int kBLoadRows = PAIR ? BLOCK_N / 2 : BLOCK_N;   // weight rows (K-major) / columns (MN-major) per CTA
  static constexpr int kBTileBytes = kBLoadRows * 128;
  static constexpr int kSfBytes = 1024;                              // MX scale atoms of a stage: SFA 512 B + SFB 512 B
  static constexpr int kStageBytes = kATileBytes + kBTileBytes + kSfBytes;
  static constexpr int kChunk = 64;                                  // accumulator columns drained per epilogue pass
  static constexpr int kChunkPitch = kChunk * 2 + 16;
  static constexpr int kChunkBytes = kBlockM * kChunkPitch;          // 18,432
  static constexpr int kRedBytes = 4 * kEpiGroups * 2 * kChunk * 4;  // [epi warps][2][64] floats
  static constexpr int kFixed = 256 /*barriers*/ + kRedBytes + 1024 /*alignment*/;
  static constexpr int kStagesFit = (232448 - kFixed - 2 * kChunkBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kChunkBytes + kFixed;
};
constexpr int kDeepMaxStages = 8;

template <int BLOCK_N, int MODE, bool STATS, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_deep_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBh,
                      const __grid_constant__ TmaSet tmAs, ConvArgs a, int n_tiles, int m_tiles) {
  using Cfg = DeepCfg<BLOCK_N, PAIR>;
  constexpr bool kBMn = mode_b_mn(MODE);
  constexpr bool kTile = mode_tile(MODE);
  constexpr int kStages = Cfg::kStages;
  static_assert(!(PAIR && kBMn && BLOCK_N < 128), "MN-major pair tiles need 64-column halves");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + kStages * Cfg::kStageBytes;                   // two chunk staging buffers
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + 2 * Cfg::kChunkBytes);
  uint64_t* empty = full + kDeepMaxStages;
  uint64_t* acc_full = empty + kDeepMaxStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);

  const int warp = threadIdx.x >> 5;
  const int KB = a.KB;
  const int crank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int m_items = PAIR ? (m_tiles + 1) / 2 : m_tiles;
  const int total = n_tiles * m_items;
  const int item0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1u);
      mbar_init(&empty[s], 1u);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1u); mbar_init(&acc_empty[b], PAIR ? 2u : 1u); }
    fence_mbar_init();
  }
  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmBh);
    tma_prefetch_desc(&tmAs.m[0]);
  }
  // single-CTA 128-wide tiles may run MX block-scaled operands: 8 more TMEM columns for the scale factors (the CTA owns
  // the SM, so rounding the allocation up to the next power of two costs nothing)
  constexpr int kTmemCols = (!PAIR && BLOCK_N == 128) ? 512 : 2 * BLOCK_N;
  if (warp == 5) {
    if (PAIR) tmem_alloc_pair<kTmemCols>(tmem_slot);
    else tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();         // the peer's mbarriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();         // preamble done; from here on the kernel reads and writes global memory

  auto tile_origin = [&](int t, int& n0, int& m0, int& tq0, int& tp0, int& tn0) -> bool {
    const int nt = t % n_tiles;
    int mt = t / n_tiles;
    bool valid = true;
    if (PAIR) {
      mt = 2 * mt + crank;
      valid = mt < m_tiles;
      if (!valid) mt = m_tiles - 1;     // padding tile of an odd M-tile count: loads + MMAs in lock-step, no stores
    }
    n0 = nt * BLOCK_N;
    m0 = mt * kBlockM;
    tq0 = tp0 = tn0 = 0;
    if (kTile) {
      const int wb = mt % a.tiles_w; mt /= a.tiles_w;
      const int hb = mt % a.tiles_h;
      const int nb = mt / a.tiles_h;
      tq0 = wb * a.tw; tp0 = hb * a.th; tn0 = nb * a.tn;
      m0 = 0;
    }
    return valid;
  };

  if (warp < 4 || warp >= 6) {
    // =================================== epilogue warps =======================================
    const int egrp = warp < 4 ? 0 : 1;
    const int qw = warp & 3;
    const int row = qw * 32 + (threadIdx.x & 31);
    const int etid = egrp * 128 + row;
    const int ew = egrp * 4 + qw;
    const uint32_t stg_u32 = smem_u32(stg);
    const uint32_t red_u32 = smem_u32(red);
    const bool bias_relu = !STATS && (a.bias != nullptr || a.relu);
    const float deq = (a.fp8 == 1 || a.fp8 == 2) ? (*a.deq_a) * (*a.deq_b) : 1.0f;   // per-tensor fp8: undo both scales
    int it = 0;
    uint32_t chunk_ctr = 0;             // staging buffer = chunk_ctr & 1 (identical sequence in every epilogue thread)
    constexpr int kChunks = BLOCK_N / Cfg::kChunk;
    // BN statistics stay in registers across the tiles of one tile column (see the persistent kernel); one accumulator
    // set per 64-column chunk, published together when the column changes and at the end
    StatAcc acc[kChunks];
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) acc[ch].clear();
    auto flush_all = [&]() {
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        if (acc[ch].n0 >= 0) epi_stats_flush<Cfg::kChunk>(red_u32, etid, ew, a, acc[ch]);
        named_bar_sync(1, kEpiThreads);          // the next flush rewrites the scratch the fold above reads
      }
    };
    for (int t = item0; t < total; t += item_step, ++it) {
      int n0, m0, tq0, tp0, tn0;
      const bool valid = tile_origin(t, n0, m0, tq0, tp0, tn0);
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1, 52);
      tc_fence_after();
      if (PAIR && !valid) {             // padding tile: keep the accumulator hand-shake going, write nothing
        tc_fence_before();
        named_bar_sync(1, kEpiThreads);
        if (etid == 0) {
          if (crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
          else mbar_arrive(&acc_empty[buf]);
        }
        continue;
      }
      int my_m;
      if (kTile) {
        const int wl = row % a.tw;
        const int t2 = row / a.tw;
        const int hl = t2 % a.th;
        const int nl = t2 / a.th;
        const bool ok = nl < a.tn && (tn0 + nl) < a.batch && (tp0 + hl) < a.dstH && (tq0 + wl) < a.dstW;
        my_m = ok ? ((tn0 + nl) * a.outH + (tp0 + hl) * a.out_stride + a.out_pa) * a.outW +
                        (tq0 + wl) * a.out_stride + a.out_pb
                  : -1;
      } else {
        my_m = (m0 + row) < a.M ? (m0 + row) : -1;
      }
      const bool zero_row = kTile && my_m < 0;
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + buf * BLOCK_N;
      if (STATS && acc[0].n0 >= 0 && acc[0].n0 != n0) flush_all();      // CTA-uniform branch
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch, ++chunk_ctr) {
        const uint32_t sbuf = stg_u32 + (chunk_ctr & 1u) * Cfg::kChunkBytes;
        const uint32_t stg_row = sbuf + row * Cfg::kChunkPitch;
        // my 32 of the chunk's 64 columns: TMEM -> (bias, ReLU) -> bf16 -> staging row
        if (bias_relu)
          epi_tmem_to_stage<true>(taddr_row + ch * Cfg::kChunk, stg_row, egrp * 32, egrp * 32 + 32, a,
                                  n0 + ch * Cfg::kChunk, zero_row, deq);
        else
          epi_tmem_to_stage<false>(taddr_row + ch * Cfg::kChunk, stg_row, egrp * 32, egrp * 32 + 32, a,
                                   n0 + ch * Cfg::kChunk, zero_row, deq);
        if (kTile && egrp == 0) sts32(stg_row + Cfg::kChunk * 2, static_cast<uint32_t>(my_m));
        if (ch == kChunks - 1) tc_fence_before();
        // all 256 threads: the chunk is staged; everybody has also left the buffer written two chunks ago
        named_bar_sync(1, kEpiThreads);
        if (ch == kChunks - 1 && etid == 0) {           // every epilogue thread has finished reading this TMEM buffer
          if (PAIR && crank != 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
          else mbar_arrive(&acc_empty[buf]);
        }
        // (no trailing barrier: the next chunk stages into the OTHER buffer, and nobody rewrites the stats scratch or
        //  this buffer before passing the next chunk's barrier, which everyone reaches only after leaving this call)
        if (STATS) epi_stats_accum<Cfg::kChunk, kTile>(sbuf, etid, a, n0 + ch * Cfg::kChunk, m0, acc[ch]);
        epi_store<Cfg::kChunk, kTile>(sbuf, etid, a, n0 + ch * Cfg::kChunk, m0);
      }
    }
    if (STATS) flush_all();
  } else if (warp == 4) {
    // ===================================== TMA producer =======================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = kTile ? static_cast<uint32_t>(a.tw * a.th * a.tn) * 128u : kATileBytes;
      const int kElems = a.fp8 ? 128 : 64;            // elements in one 128-byte operand row = K extent of a k-block
      const int mnBoxes = Cfg::kBLoadRows / kElems;   // MN-major B: boxes of kElems columns x kElems K-rows (128 B rows)
      const uint32_t mnBoxBytes = 128u * static_cast<uint32_t>(kElems);
      for (int t = item0; t < total; t += item_step) {
        int n0, m0, tq0, tp0, tn0;
        tile_origin(t, n0, m0, tq0, tp0, tn0);
        if (t + item_step >= total) pdl_trigger_last_tile();
        int tap = 0, cc = 0, tap_r = 0, tap_s = 0;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, 49);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          const uint32_t sB = smem_u32(sA + kATileBytes);
          const uint32_t pair_bar = PAIR ? mapa_shared(smem_u32(&full[stage]), 0) : 0u;
          const bool mx = !PAIR && a.fp8 == 3;
          if (!PAIR) mbar_arrive_expect_tx(&full[stage], Cfg::kBTileBytes + a_bytes + (mx ? 1024u : 0u));
          else if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * (Cfg::kBTileBytes + a_bytes));
          if (mx) {      // this k-block's scale atoms (GEMM mode: tile = row block m0/128, weight block n0/128)
            const uint32_t sSF = sB + Cfg::kBTileBytes;
            bulk_load(sSF, a.sfa + (static_cast<size_t>(m0 / kBlockM) * KB + kb) * 512, 512, &full[stage]);
            bulk_load(sSF + 512, a.sfb + (static_cast<size_t>(n0 / 128) * KB + kb) * 512, 512, &full[stage]);
          }
          int widx = tap, dh = 0, dw = 0, mapi = 0;
          if (kTile) {
            if (a.ntaps > 0) {
              widx = a.tap_widx[tap]; dh = a.tap_dh[tap]; dw = a.tap_dw[tap]; mapi = a.tap_map[tap];
            } else if (MODE == kConvTileFwd) {
              dh = tap_r * a.dil - a.pad; dw = tap_s * a.dil - a.pad_w;
            } else {
              dh = a.pad - tap_r * a.dil; dw = a.pad_w - tap_s * a.dil;
            }
          }
          const int kcoord = widx * a.kstride + cc * kElems;
          const int nb0 = n0 + crank * Cfg::kBLoadRows;        // first weight row / column this CTA fetches
          if (PAIR) {
            if (kBMn) {
              for (int j = 0; j < mnBoxes; ++j)
                tma_load_2d_pair(sB + j * mnBoxBytes, &tmB, widx * a.ldc + nb0 + j * kElems, cc * kElems, pair_bar);
            } else if (Cfg::kBLoadRows == 128) {
              tma_load_2d_pair(sB, &tmB, kcoord, nb0, pair_bar);           // the full map's box is 128 rows
            } else {
              tma_load_2d_pair(sB, &tmBh, kcoord, nb0, pair_bar);          // half-height box map
            }
            if (kTile) tma_load_4d_pair(smem_u32(sA), &tmAs.m[mapi], cc * kElems, tq0 + dw, tp0 + dh, tn0, pair_bar);
            else tma_load_2d_pair(smem_u32(sA), &tmAs.m[0], kb * kElems, m0, pair_bar);
          } else {
            if (kBMn) {
              for (int j = 0; j < mnBoxes; ++j)
                tma_load_2d(sB + j * mnBoxBytes, &tmB, widx * a.ldc + n0 + j * kElems, cc * kElems, &full[stage]);
            } else {
              constexpr int kBoxRows = BLOCK_N < 128 ? BLOCK_N : 128;      // the weight map's box height
#pragma unroll
              for (int h = 0; h < BLOCK_N / kBoxRows; ++h)
                tma_load_2d(sB + h * kBoxRows * 128, &tmB, kcoord, n0 + h * kBoxRows, &full[stage]);
            }
            if (kTile) tma_load_4d(smem_u32(sA), &tmAs.m[mapi], cc * kElems, tq0 + dw, tp0 + dh, tn0, &full[stage]);
            else tma_load_2d(smem_u32(sA), &tmAs.m[0], kb * kElems, m0, &full[stage]);
          }
          if (++cc == a.cchunks) { cc = 0; ++tap; if (++tap_s == a.S) { tap_s = 0; ++tap_r; } }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ====================================== MMA issuer ========================================
    constexpr int kMmaM = PAIR ? 2 * kBlockM : kBlockM;
    const bool f8 = a.fp8 != 0;
    const bool mxk = !PAIR && BLOCK_N == 128 && !kBMn && a.fp8 == 3;      // MX block-scaled operands (K-major GEMM only)
    const uint32_t idesc = f8 ? idesc_f8(kMmaM, BLOCK_N, a.fp8 == 2 ? 1 : 0, 0, 0, kBMn ? 1 : 0)
                              : idesc_bf16(kMmaM, BLOCK_N, 0, kBMn ? 1 : 0);
    // MN-major B: K rows of 128 bytes; one MMA consumes 16 (bf16) / 32 (fp8) K rows, chunks of 64 / 128 columns
    const uint32_t mnKStep = f8 ? 4096u : 2048u, mnLbo = f8 ? 16384u : 8192u;
    // descriptors of stage 0 / K slice 0, built once; stages and K slices are 32-bit adds on the low word
    const uint32_t sA0 = smem_u32(smem), sB0 = sA0 + kATileBytes;
    const uint64_t dA0 = smem_desc_sw128(sA0, 0, 1024);
    const uint64_t dB0 = kBMn ? smem_desc_sw128(sB0, mnLbo, 1024) : smem_desc_sw128(sB0, 0, 1024);
    const uint32_t a_lo = static_cast<uint32_t>(dA0), a_hi = static_cast<uint32_t>(dA0 >> 32);
    const uint32_t b_lo = static_cast<uint32_t>(dB0), b_hi = static_cast<uint32_t>(dB0 >> 32);
    constexpr uint32_t kStageStep = Cfg::kStageBytes >> 4;
    const uint32_t bStep = kBMn ? (mnKStep >> 4) : 2u;         // K slice of one MMA in 16-byte units (A: always 2)
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (!(PAIR && crank != 0)) {        // the leader issues the MMAs of a pair
      for (int t = item0; t < total; t += item_step, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1u, 51);     // epilogue(s) have drained this TMEM buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase, 50);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t so = static_cast<uint32_t>(stage) * kStageStep;
            if (mxk) {
              // scale atoms of this k-block: shared memory -> TMEM (4 columns each), in order with the MMAs that use them
              const uint32_t sSF = sB0 + Cfg::kBTileBytes + static_cast<uint32_t>(stage) * Cfg::kStageBytes;
              tmem_cp_32x128b_warpx4(tmem_base + 2 * BLOCK_N, smem_desc_nosw(sSF, 0, 128));
              tmem_cp_32x128b_warpx4(tmem_base + 2 * BLOCK_N + 4, smem_desc_nosw(sSF + 512, 0, 128));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {                      // 4 x 32 bytes of K per 128-byte row
              const uint64_t da = desc_join(a_lo + so + 2u * k, a_hi);
              const uint64_t db = desc_join(b_lo + so + bStep * k, b_hi);
              const bool acc = (kb | k) != 0;
              if (mxk) {
                umma_mxf8(d_tmem, da, db, idesc_mxf8(kBlockM, BLOCK_N, 0, 0, k, k), acc, tmem_base + 2 * BLOCK_N,
                          tmem_base + 2 * BLOCK_N + 4);
              } else if (f8) {
                if (PAIR) umma_f8_pair(d_tmem, da, db, idesc, acc);
                else umma_f8(d_tmem, da, db, idesc, acc);
              } else {
                if (PAIR) umma_bf16_pair(d_tmem, da, db, idesc, acc);
                else umma_bf16(d_tmem, da, db, idesc, acc);
              }
            }
            if (PAIR) {
              umma_commit_pair_multicast(&empty[stage], 0x3);          // frees the stage in both CTAs
              if (kb == KB - 1) umma_commit_pair_multicast(&acc_full[buf], 0x3);
            } else {
              umma_commit(&empty[stage]);
              if (kb == KB - 1) umma_commit(&acc_full[buf]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();         // no CTA leaves while its peer may still signal its barriers / read its smem
  if (warp == 5) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad kernel: dW[co][k] += sum over a pixel range of dY[m][co] * im2col(X)[m][k]
// ---------------------------------------------------------------------------------------------
constexpr int kWgMaxStages = 3;
constexpr int kWgStageBytes = 32768;           // A: 2 x [64 pix][64 co], B: 2 x [64 pix][64 k]
constexpr int kWgPitch = 132;                  // fp32 staging pitch (floats)
constexpr int kWgEpiBytes = kBlockM * kWgPitch * 4;
constexpr int wg_smem_bytes(int stages) {
  return (stages * kWgStageBytes > kWgEpiBytes ? stages * kWgStageBytes : kWgEpiBytes) + 256 + 1024;
}

// SWAP = operand roles exchanged for narrow outputs (Cout tile of 64): the 128-row MMA dimension carries the k
// columns and N = 64 output channels, so no half of the tensor-core tile multiplies zero rows (ncu: the Cout = 64
// layers kept the pipe ~49 % busy doing 50 % useful work).  The accumulator comes out transposed ([k col][co]) and
// is staged transposed so the global reductions stay row-contiguous.
template <int MODE, bool SWAP>
__global__ void __launch_bounds__(kThreads, 3)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ TmaSet tmXs, WgradArgs a) {
  const CUtensorMap& tmX = tmXs.m[0];
  constexpr bool kXTma = (MODE == kConvGemm || MODE == kConvTileFwd || MODE == kConvStemTma);
  constexpr bool kTile = (MODE == kConvTileFwd || MODE == kConvStemTma);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nstages = a.stages;
  const int pipe_bytes = nstages * kWgStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (pipe_bytes > kWgEpiBytes ? pipe_bytes : kWgEpiBytes));
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* acc_full = empty + kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int col0 = blockIdx.x * 128;          // first k column of this tile
  constexpr int kCoTile = SWAP ? 64 : 128;
  const int co0 = blockIdx.y * kCoTile;       // first output channel of this tile
  const int kb_begin = blockIdx.z * a.kb_per_split;
  const int kb_end = min(kb_begin + a.kb_per_split, a.total_kb);
  const int KB = kb_end - kb_begin;
  if (KB <= 0) return;                         // uniform per CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full[s], kXTma ? 1u : (kProducerThreads + 1u));
      mbar_init(&empty[s], 1u);
    }
    mbar_init(acc_full, 1u);
    fence_mbar_init();
  }
  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tmDy);
    if (kXTma) tma_prefetch_desc(&tmX);
  }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  if (kTile) {
    // pixel boxes may hold fewer than 64 rows: the unused K rows of EVERY operand chunk must read as zero
    const int rows = a.tw * a.th * a.tn;
    const int tail16 = (64 - rows) * 8;        // 16-byte units per 8 KB chunk
    for (int st = 0; st < nstages; ++st)
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint4* base = reinterpret_cast<uint4*>(smem + st * kWgStageBytes + chunk * 8192 + rows * 128);
        for (int i = threadIdx.x; i < tail16; i += kThreads) base[i] = make_uint4(0, 0, 0, 0);
      }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4 || warp >= 6) {
    const int egrp = warp < 4 ? 0 : 1;
    const int qw = warp & 3;
    const int ew = egrp * 4 + qw;
    if (!kXTma && warp < 4) {
      const int chunk = threadIdx.x >> 6;      // which 64-column half of the B tile
      const int row = threadIdx.x & 63;        // pixel row inside the k-block
      const int colc = col0 + chunk * 64;      // first k column of my chunk
      const bool col_ok = colc < a.ncols;
      int tap_r = 0, tap_s = 0, c_off = 0;
      if (!mode_stem(MODE)) {
        const int tap = colc / a.Cpad;          // columns live in the padded (tap, Cpad) space
        c_off = colc - tap * a.Cpad;
        tap_r = tap / a.S;
        tap_s = tap - tap_r * a.S;
      }
      const int crem = a.C - c_off;             // real channels left in this tap
      const uint32_t row_off = 16384u + chunk * 8192u + row * 128u;
      const uint32_t sw = row & 7u;
      const int pq = a.P * a.Q;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < KB; ++i) {
        mbar_wait(&empty[stage], phase ^ 1u, 65);
        const uint32_t dst = smem_u32(smem + stage * kWgStageBytes) + row_off;
        const int m = (kb_begin + i) * 64 + row;
        const bool m_ok = col_ok && m < a.M;
        int img = 0, p = 0, q = 0;
        if (m_ok) {
          img = m / pq;
          const int rem = m - img * pq;
          p = rem / a.Q;
          q = rem - p * a.Q;
        }
        const int hb = p * a.stride - a.pad, wb = q * a.stride - a.pad_w;
        const __nv_bfloat16* img_base = a.x + static_cast<size_t>(img) * a.H * a.W * a.C;
        if (MODE == kConvStem) {
          const int SP = a.cchunks, RPK = 16 / SP;
          const int kbk = colc / 64;            // which k-block of the stem's packed K
          for (int rr = 0; rr < RPK; ++rr) {
            const int r = kbk * RPK + rr;
            const int h = hb + r * a.dil;
            const bool h_ok = m_ok && r < a.R && h >= 0 && h < a.H;
            for (int s = 0; s < SP; ++s) {
              const int e = s * RPK + rr;
              const int w = wb + s * a.dil;
              const bool ok = h_ok && s < a.S && w >= 0 && w < a.W;
              const __nv_bfloat16* src = ok ? img_base + (static_cast<size_t>(h) * a.W + w) * 4 : a.x;
              cp_async_8(dst + (((e >> 1) ^ sw) << 4) + ((e & 1) << 3), src, ok);
            }
          }
        } else {
          const int h = hb + tap_r * a.dil, w = wb + tap_s * a.dil;
          const bool ok = m_ok && h >= 0 && h < a.H && w >= 0 && w < a.W;
          const __nv_bfloat16* src = ok ? img_base + (static_cast<size_t>(h) * a.W + w) * a.C + c_off : a.x;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool okj = ok && j * 8 < crem;
            cp_async_16(dst + ((static_cast<uint32_t>(j) ^ sw) << 4), okj ? src + j * 8 : a.x, okj);
          }
        }
        cp_async_commit();
        if (i >= kLag) {
          cp_async_wait<kLag>();
          fence_proxy_async_smem();
          mbar_arrive(&full[(i - kLag) % nstages]);
        }
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (int i = (KB > kLag ? KB - kLag : 0); i < KB; ++i) mbar_arrive(&full[i % nstages]);
    }

    // epilogue: TMEM -> fp32 staging -> coalesced vector reductions into the gradient arena
    mbar_wait(acc_full, 0, 68);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem);
    const int row = qw * 32 + (threadIdx.x & 31);
    constexpr int kAccCols = SWAP ? 64 : 128;
#pragma unroll 1
    for (int c0 = egrp * (kAccCols / kEpiGroups); c0 < (egrp + 1) * (kAccCols / kEpiGroups); c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (SWAP) {
        // accumulator row = k column, accumulator column = output channel: stage as [co][k col]
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[(c0 + j) * kWgPitch + row] = __uint_as_float(v[j]);
      } else {
        float4* d = reinterpret_cast<float4*>(stg + row * kWgPitch + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3]));
      }
    }
    tc_fence_before();
    named_bar_sync(1, kEpiThreads);
    const int lane = threadIdx.x & 31;
    const int vc = col0 + lane * 4;                    // column in the padded (tap, Cpad) space
    const int vtap = vc / a.Cpad;
    const int vch = vc - vtap * a.Cpad;
    if (vc < a.ncols && vch < a.Cw) {                  // channel padding of a tap has no dw column
      const int c = vtap * a.Cw + vch;
      for (int r = ew; r < kCoTile; r += 4 * kEpiGroups) {
        const int co = co0 + r;
        if (co < a.Cout) {
          const float4 val = *reinterpret_cast<const float4*>(stg + r * kWgPitch + lane * 4);
          red_add_f32x4(a.dw + static_cast<size_t>(co) * a.ldw + c, val.x, val.y, val.z, val.w);
        }
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      // taps / channel offsets of the two 64-column chunks of the B tile (tile mode)
      int tdh[2] = {0, 0}, tdw[2] = {0, 0}, tmap[2] = {0, 0}, tc0[2] = {0, 0};
      if (kTile && MODE != kConvStemTma) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int colc = col0 + j * 64;
          const int tap = colc / a.Cpad;
          tc0[j] = colc - tap * a.Cpad;
          const int r = tap / a.S, sx = tap - r * a.S;
          const int oh = r * a.dil - a.pad, ow = sx * a.dil - a.pad_w;
          if (a.stride == 1) {
            tdh[j] = oh; tdw[j] = ow;
          } else {               // stride-2 source: box comes from the (pa, pb) phase sub-image
            const int st = a.stride;
            const int pa = ((oh % st) + st) % st, pb = ((ow % st) + st) % st;
            tdh[j] = (oh - pa) / st; tdw[j] = (ow - pb) / st;
            tmap[j] = pa * st + pb;
          }
        }
      }
      const uint32_t box_bytes = kTile ? static_cast<uint32_t>(a.tw * a.th * a.tn) * 128u : 8192u;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < KB; ++i) {
        mbar_wait(&empty[stage], phase ^ 1u, 65);
        const uint32_t sA = smem_u32(smem + stage * kWgStageBytes);
        if (kTile) {
          int t = kb_begin + i;
          const int wb = t % a.tiles_w; t /= a.tiles_w;
          const int hb = t % a.tiles_h;
          const int nb = t / a.tiles_h;
          const int q0 = wb * a.tw, p0 = hb * a.th, n0 = nb * a.tn;
          const bool second = (col0 + 64) < a.ncols;
          mbar_arrive_expect_tx(&full[stage], box_bytes * ((second ? 4u : 3u) - (SWAP ? 1u : 0u)));
          tma_load_4d(sA, &tmDy, co0, q0, p0, n0, &full[stage]);
          if (!SWAP) tma_load_4d(sA + 8192, &tmDy, co0 + 64, q0, p0, n0, &full[stage]);
          if (MODE == kConvStemTma) {
            // k-block (col0 / 64 + j) of the packed stem K = one box of the row-interleaved image
            tma_load_5d(sA + 16384, &tmXs.m[0], 0, col0 / 64, q0, p0, n0, &full[stage]);
            if (second) tma_load_5d(sA + 24576, &tmXs.m[0], 0, col0 / 64 + 1, q0, p0, n0, &full[stage]);
          } else {
            tma_load_4d(sA + 16384, &tmXs.m[tmap[0]], tc0[0], q0 + tdw[0], p0 + tdh[0], n0, &full[stage]);
            if (second) tma_load_4d(sA + 24576, &tmXs.m[tmap[1]], tc0[1], q0 + tdw[1], p0 + tdh[1], n0, &full[stage]);
          }
        } else {
          const int m = (kb_begin + i) * 64;
          mbar_arrive_expect_tx(&full[stage], (SWAP ? 8192u : 16384u) + (kXTma ? 16384u : 0u));
          tma_load_2d(sA, &tmDy, co0, m, &full[stage]);
          if (!SWAP) tma_load_2d(sA + 8192, &tmDy, co0 + 64, m, &full[stage]);
          if (kXTma) {
            tma_load_2d(sA + 16384, &tmX, col0, m, &full[stage]);
            tma_load_2d(sA + 24576, &tmX, col0 + 64, m, &full[stage]);
          }
        }
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    constexpr uint32_t idesc = idesc_bf16(128, SWAP ? 64 : 128, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < KB; ++i) {
      mbar_wait(&full[stage], phase, 66);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sDy = smem_u32(smem + stage * kWgStageBytes);
        const uint32_t sX = sDy + 16384;
        const uint32_t sA = SWAP ? sX : sDy;      // M operand: 128 rows = two 64-wide MN-major chunks 8 KB apart
        const uint32_t sB = SWAP ? sDy : sX;      // N operand (SWAP: one 64-channel chunk of dY)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = smem_desc_sw128(sA + k * 2048, 8192, 1024);
          const uint64_t db = smem_desc_sw128(sB + k * 2048, 8192, 1024);
          umma_bf16(tmem_base, da, db, idesc, (i | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (i == KB - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launchers
// ---------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(dlsym(lib, "cuTensorMapEncodeTiled"));
  }();
  return fn;
}

// 2-D bf16 row-major matrix [rows][cols] (cols contiguous), box = {box_cols, box_rows}, 128B swizzle.
// `esize` = element size in bytes: 2 = bf16, 1 = fp8 (maps over bytes; a 128-byte swizzle row is 128 elements then).
bool make_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                 uint32_t box_cols, uint32_t box_rows, int esize = 2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * static_cast<uint64_t>(esize)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 4-D NHWC bf16 activation [N][H][W][C]; box = {64 channels, tw, th, tn}; out-of-bounds -> zeros.
bool make_map_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int tw, int th, int tn,
                   int esize = 2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t es = static_cast<cuuint64_t>(esize);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * es, static_cast<cuuint64_t>(W) * C * es,
                           static_cast<cuuint64_t>(H) * W * C * es};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(128 / esize), static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th),
                       static_cast<cuuint32_t>(tn)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Phase (pa, pb) sub-image of a stride-`st` NHWC source: X[n][st*i + pa][st*j + pb][c] as a dense 4-D map.
bool make_map_nhwc_phase(CUtensorMap* map, const void* base, int N, int H, int W, int C, int st, int pa, int pb,
                         int tw, int th, int tn, int esize = 2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const int Hp = (H - pa + st - 1) / st, Wp = (W - pb + st - 1) / st;
  if (Hp <= 0 || Wp <= 0) return make_map_nhwc(map, base, N, H, W, C, tw, th, tn, esize);   // never referenced
  const cuuint64_t es = static_cast<cuuint64_t>(esize);
  const char* b = static_cast<const char*>(base) + (static_cast<size_t>(pa) * W + pb) * C * es;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(Wp), static_cast<cuuint64_t>(Hp),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(st) * C * es, static_cast<cuuint64_t>(st) * W * C * es,
                           static_cast<cuuint64_t>(H) * W * C * es};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(128 / esize), static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th),
                       static_cast<cuuint32_t>(tn)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(b), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Stem operand map over the zero-padded, G-row-interleaved image [N][Hp][Wp][G][4] written by pad_nhwc4
// (G = RPK filter rows per k-block): the k-block j of output pixel (n, p, q) is the 128 contiguous bytes at
// (row stride*p + G*j, col stride*q).  5-D map (64 elements | k-block | out col | out row | image) whose dimensions
// overlap in memory; box = {64, 1, tw, th, tn}.
bool make_map_stem5d(CUtensorMap* map, const void* base, int N, int Hp, int Wp, int P, int Q, int stride, int G, int KB,
                     int tw, int th, int tn) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t pitch = static_cast<cuuint64_t>(Wp) * G * 8;          // bytes per buffer row
  cuuint64_t dims[5] = {64, static_cast<cuuint64_t>(KB), static_cast<cuuint64_t>(Q), static_cast<cuuint64_t>(P),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[4] = {pitch * G, static_cast<cuuint64_t>(stride) * G * 8, pitch * stride, pitch * Hp};
  cuuint32_t box[5] = {64, 1, static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), static_cast<cuuint32_t>(tn)};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 4; ++i)
    if (strides[i] % 16 != 0) return false;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline int stem_sp(int S) { return S <= 4 ? 4 : (S <= 8 ? 8 : 16); }

inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int BLOCK_N, int MODE, bool STATS>
cudaError_t launch_fwd_t(const CUtensorMap& tmB, const TmaSet& tmA, const ConvArgs& a, int n_total, int m_tiles,
                         cudaStream_t stream) {
  using Cfg = FwdCfg<BLOCK_N>;
  auto kern = conv_gemm_kernel<BLOCK_N, MODE, STATS>;
  static int configured = 0;
  const int smem = Cfg::smem_bytes(a.stages);
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::smem_bytes(Cfg::kMaxStagesN));
    if (e != cudaSuccess) return e;
    configured = Cfg::smem_bytes(Cfg::kMaxStagesN);
  }
  dim3 grid(n_total / BLOCK_N, m_tiles);
  return launch_pdl(kern, grid, dim3(kThreads), smem, stream, 1, tmB, tmA, a);
}

int g_persistent = 1;        // TMA-fed modes use the persistent kernel (tuning hook: set_conv_persistent)
int g_num_sms = 0;

template <int BLOCK_N, int MODE, bool STATS, int CL>
cudaError_t launch_persistent_t(const CUtensorMap& tmB, const CUtensorMap& tmBh, const TmaSet& tmA, const ConvArgs& a,
                                int n_total, int m_tiles, cudaStream_t stream) {
  using Cfg = PersistCfg<BLOCK_N>;
  constexpr bool CLUSTER = CL != 0;
  auto kern = conv_gemm_persistent_kernel<BLOCK_N, MODE, STATS, CL>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const int n_tiles = n_total / BLOCK_N;
  const long long items = static_cast<long long>(n_tiles) * (CLUSTER ? (m_tiles + 1) / 2 : m_tiles);
  long long grid = (BLOCK_N > 128 ? 1LL : 2LL) * g_num_sms;      // BLOCK_N = 256 owns the SM's whole TMEM
  if (CLUSTER) {
    if (grid > 2 * items) grid = 2 * items;
    grid &= ~1LL;
    return launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(kThreads), Cfg::kSmemBytes, stream, 2, tmB, tmBh, tmA, a,
                      n_tiles, m_tiles);
  }
  if (grid > items) grid = items;
  return launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(kThreads), Cfg::kSmemBytes, stream, 1, tmB, tmBh, tmA, a,
                    n_tiles, m_tiles);
}

template <int BLOCK_N, int MODE, bool STATS, bool PAIR>
cudaError_t launch_deep_t(const CUtensorMap& tmB, const CUtensorMap& tmBh, const TmaSet& tmA, const ConvArgs& a,
                          int n_total, int m_tiles, cudaStream_t stream) {
  using Cfg = DeepCfg<BLOCK_N, PAIR>;
  auto kern = conv_gemm_deep_kernel<BLOCK_N, MODE, STATS, PAIR>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const int n_tiles = n_total / BLOCK_N;
  const long long items = static_cast<long long>(n_tiles) * (PAIR ? (m_tiles + 1) / 2 : m_tiles);
  long long grid = g_num_sms;                      // one CTA per SM
  if (PAIR) {
    if (grid > 2 * items) grid = 2 * items;
    grid &= ~1LL;
  } else if (grid > items) {
    grid = items;
  }
  return launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(kThreads), Cfg::kSmemBytes, stream, PAIR ? 2 : 1, tmB, tmBh,
                    tmA, a, n_tiles, m_tiles);
}

// Variant word (ConvArgs::variant; 0 = built-in policy).  The Python layer autotunes it per layer shape on first use
// (ops/native.py: the analogue of the reference's `cudnn.benchmark = True`, pytorch_synthetic_benchmark.py:57).
//   bits 0-3  kernel: 1 = one tile per CTA, 2 = persistent (2 CTAs/SM, 2-3 stage ring), 3 = deep ring (1 CTA/SM)
//   bits 4-7  tile width: 0 = widest that divides Cout, 1 = 64, 2 = 128, 3 = 256 (deep only)
//   bit  8    deep only: CTA pairs sharing one M = 256 MMA (cta_group::2)
constexpr int kVarOneTile = 1, kVarPersistent = 2, kVarDeep = 3;

template <int MODE>
constexpr bool mode_deep(){ return MODE == kConvGemm || MODE == kConvTileFwd || MODE == kConvTileDgrad || MODE == kConvGemmDgrad; }

template <int MODE>
cudaError_t launch_deep_mode(const CUtensorMap& tmB, const CUtensorMap& tmBh, const TmaSet& tmA, const ConvArgs& a,
                             int n_total, int m_tiles, bool stats, int bn, bool pair, cudaStream_t stream) {
  if constexpr (!mode_deep<MODE>()) {
    return cudaErrorInvalidValue;
  } else {
#define DDL_DEEP(BN, PR)                                                                                     \
  return stats ? launch_deep_t<BN, MODE, true, PR>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)              \
               : launch_deep_t<BN, MODE, false, PR>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
    if (bn == 256) { if (pair) { DDL_DEEP(256, true); } else { DDL_DEEP(256, false); } }
    if (bn == 128) { if (pair) { DDL_DEEP(128, true); } else { DDL_DEEP(128, false); } }
    if (bn == 64) {
      if constexpr (!mode_b_mn(MODE)) { if (pair) { DDL_DEEP(64, true); } }
      DDL_DEEP(64, false);
    }
#undef DDL_DEEP
    return cudaErrorInvalidValue;
  }
}

// Pipeline-depth policy.  Short-K tiles are dominated by per-CTA fixed latency (TMEM alloc, first TMA
// round trip, epilogue), so they get a shallow ring -> small shared-memory footprint -> 3-4 CTAs per SM whose
// prologues/epilogues overlap.  Long-K tiles get the deep ring.  g_force_stages (tuning hook) overrides.
int g_force_stages = 0;
int g_cluster = 0;          // tuning hook: 1 = CTA pairs with TMA-multicast weight tiles, 2 = cta_group::2 pair MMAs
int g_bn256 = 0;            // tuning hook: 1 lets long-K layers use 128 x 256 persistent tiles (measured 1.7 % SLOWER on
                            // ResNet-50: one CTA per SM leaves the epilogue half the warps; kept for A/B runs)
int g_deep = 1;             // tuning hook: 0 = never use the deep-ring kernel, 1 = policy, 2-4 = wherever it applies
int g_wgrad_swap = 1;       // tuning hook: 0 disables the operand-role swap of narrow-output wgrad tiles

template <int BLOCK_N, int MODE>
int pick_stages(int KB) {
  const int max_s = FwdCfg<BLOCK_N>::kMaxStagesN;
  int s;
  if (g_force_stages > 0) s = g_force_stages;
  else if (!mode_a_tma(MODE)) s = max_s;          // cp.async gather: needs depth > kLag whenever it wraps
  else s = (KB <= 18) ? 2 : max_s;
  if (s > max_s) s = max_s;
  if (s > KB) s = KB;
  if (!mode_a_tma(MODE) && KB > s && s <= kLag) s = kLag + 1;
  return s < 1 ? 1 : s;
}

struct WeightDesc {       // what the B tensor maps are built from (their box height depends on the tile width chosen)
  const void* w;
  int rows, cols, esize;
};

// B-operand maps for a kernel whose N tile is `bn` wide: full box (min(bn, 128) weight rows) and half box (bn / 2 rows: one
// CTA's share of a pair tile).  MN-major weights (data gradient) use square boxes of one swizzle row x as many K rows.
bool make_b_maps(int mode, const WeightDesc& wd, int bn, CUtensorMap* full, CUtensorMap* half) {
  const uint32_t kcols = 128u / wd.esize;
  if (mode_b_mn(mode)) {
    if (!make_map_2d(full, wd.w, wd.rows, wd.cols, wd.cols, kcols, kcols, wd.esize)) return false;
    *half = *full;
    return true;
  }
  if (!make_map_2d(full, wd.w, wd.rows, wd.cols, wd.cols, kcols, bn < 128 ? bn : 128, wd.esize)) return false;
  if (!make_map_2d(half, wd.w, wd.rows, wd.cols, wd.cols, kcols, bn / 2 < 128 ? bn / 2 : 128, wd.esize)) *half = *full;
  return true;
}

template <int MODE>
cudaError_t launch_fwd_mode(const WeightDesc& wd, const CUtensorMap& tmB, const CUtensorMap& tmBh, bool have_half_map,
                            const TmaSet& tmA, ConvArgs a, int n_total, int m_tiles, bool stats, cudaStream_t stream) {
  const int var_kind = a.variant & 0xf, var_bn = (a.variant >> 4) & 0xf, var_pair = (a.variant >> 8) & 1;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const long long tiles128 = static_cast<long long>(n_total / ((n_total % 128 == 0) ? 128 : 64)) * m_tiles;
  // ---- deep-ring kernel (one CTA / CTA pair per SM): long-K layers with at least a couple of waves of tiles -------
  if (mode_deep<MODE>() && g_deep &&
      (var_kind == kVarDeep || (var_kind == 0 && (g_deep >= 2 || (a.KB >= 8 && tiles128 >= 2LL * g_num_sms))))) {
    int bn = (n_total % 256 == 0) ? 256 : ((n_total % 128 == 0) ? 128 : 64);
    if (var_kind == kVarDeep && var_bn != 0) {
      const int want = var_bn == 1 ? 64 : (var_bn == 2 ? 128 : 256);
      if (n_total % want != 0) return cudaErrorInvalidValue;
      bn = want;
    }
    // test hooks: g_deep = 2 forces the deep kernel with pairs wherever possible, 3 = without pairs, 4 = without pairs
    // and without 256-wide tiles
    bool pair = var_kind == kVarDeep ? var_pair != 0 : (g_deep != 3 && g_deep != 4);
    if (var_kind != kVarDeep && g_deep == 4 && bn == 256) bn = 128;
    // pairs need two M tiles, 64-column halves of an MN-major weight tile, and a loadable half of a K-major one
    // (bn = 256: the full map's 128-row box; narrower: the half-height box map)
    if (m_tiles < 2 || (mode_b_mn(MODE) && bn < (a.fp8 ? 256 : 128))) pair = false;
    if (a.fp8 && mode_b_mn(MODE) && bn < 128) return cudaErrorInvalidValue;
    // the weight maps' box heights follow THIS kernel's tile width (not the widest tile Cout would allow)
    CUtensorMap dB, dBh;
    if (!make_b_maps(MODE, wd, bn, &dB, &dBh)) return cudaErrorUnknown;
    return launch_deep_mode<MODE>(dB, dBh, tmA, a, n_total, m_tiles, stats, bn, pair, stream);
  }
  if (var_kind == kVarDeep) return cudaErrorInvalidValue;
  bool persistent = false;
  if (mode_a_tma(MODE) && g_persistent) {
    // persistence pays when every CTA gets several tiles (measured crossover ~3.5 tiles per resident CTA);
    // below that the one-tile-per-CTA kernel with 3 CTAs/SM wins.  g_persistent == 2 forces it (tuning).
    persistent = g_persistent == 2 || tiles128 * 2 >= 7LL * 2 * g_num_sms;
  }
  if (mode_a_tma(MODE) && var_kind == kVarPersistent) persistent = true;
  if (var_kind == kVarOneTile) persistent = false;
  // 128 x 256 tiles: only where the main loop (operand traffic) matters (K >= 256) — with one CTA per SM the epilogue
  // has half the warps to hide its latency, which would cost the short-K, store-bound layers
  if (persistent && g_bn256 && n_total % 256 == 0 && a.KB >= 4 &&
      static_cast<long long>(n_total / 256) * m_tiles >= 4LL * g_num_sms) {
    return stats ? launch_persistent_t<256, MODE, true, 0>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
                 : launch_persistent_t<256, MODE, false, 0>(tmB, tmBh, tmA, a, n_total, m_tiles, stream);
  }
  if (persistent) {
    // CTA pairs sharing the weight tile through TMA multicast: where operand traffic matters (K >= 128) and the half
    // tile is loadable (K-major: half-height box map; MN-major: the two 64-column boxes of a 128-wide tile)
    const bool pair = g_cluster && m_tiles >= 2 && a.KB >= 2 &&
                      (mode_b_mn(MODE) ? (n_total % 128 == 0) : have_half_map);
    if (n_total % 128 == 0) {
      if (pair && g_cluster == 2)
        return stats ? launch_persistent_t<128, MODE, true, 2>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
                     : launch_persistent_t<128, MODE, false, 2>(tmB, tmBh, tmA, a, n_total, m_tiles, stream);
      if (pair)
        return stats ? launch_persistent_t<128, MODE, true, 1>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
                     : launch_persistent_t<128, MODE, false, 1>(tmB, tmBh, tmA, a, n_total, m_tiles, stream);
      return stats ? launch_persistent_t<128, MODE, true, 0>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
                   : launch_persistent_t<128, MODE, false, 0>(tmB, tmBh, tmA, a, n_total, m_tiles, stream);
    }
    if (n_total % 64 == 0) {
      if (pair && !mode_b_mn(MODE) && g_cluster == 2)
        return stats ? launch_persistent_t<64, MODE, true, 2>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
                     : launch_persistent_t<64, MODE, false, 2>(tmB, tmBh, tmA, a, n_total, m_tiles, stream);
      if (pair && !mode_b_mn(MODE))
        return stats ? launch_persistent_t<64, MODE, true, 1>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
                     : launch_persistent_t<64, MODE, false, 1>(tmB, tmBh, tmA, a, n_total, m_tiles, stream);
      return stats ? launch_persistent_t<64, MODE, true, 0>(tmB, tmBh, tmA, a, n_total, m_tiles, stream)
                   : launch_persistent_t<64, MODE, false, 0>(tmB, tmBh, tmA, a, n_total, m_tiles, stream);
    }
    return cudaErrorInvalidValue;
  }
  if (n_total % 128 == 0) {
    a.stages = pick_stages<128, MODE>(a.KB);
    return stats ? launch_fwd_t<128, MODE, true>(tmB, tmA, a, n_total, m_tiles, stream)
                 : launch_fwd_t<128, MODE, false>(tmB, tmA, a, n_total, m_tiles, stream);
  }
  if (n_total % 64 == 0) {
    a.stages = pick_stages<64, MODE>(a.KB);
    return stats ? launch_fwd_t<64, MODE, true>(tmB, tmA, a, n_total, m_tiles, stream)
                 : launch_fwd_t<64, MODE, false>(tmB, tmA, a, n_total, m_tiles, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

// {abort flag, site, block, thread, parity} of the first mbarrier wait that timed out since the last call; clears it.
cudaError_t conv_timeout_info(unsigned int out[8]) {
  cudaError_t e = cudaMemcpyFromSymbol(out, tc::g_mbar_diag, sizeof(unsigned int) * 8);
  if (e != cudaSuccess) return e;
  if (out[0] != 0u) {
    unsigned int zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyToSymbol(tc::g_mbar_diag, zero, sizeof(zero));
  }
  return e;
}

void set_conv_force_stages(int s) { g_force_stages = s; }
void set_conv_persistent(int on) { g_persistent = on; }
void set_wgrad_swap(int on) { g_wgrad_swap = on; }
void set_conv_bn256(int on) { g_bn256 = on; }
void set_conv_cluster(int on) { g_cluster = on; }
void set_conv_deep(int on) { g_deep = on; }
void set_conv_wait_hint(int ns) {
  const unsigned int v = ns > 0 ? static_cast<unsigned int>(ns) : 0u;
  cudaMemcpyToSymbol(tc::c_wait_hint_ns, &v, sizeof(v));
}
void pdl_early_bn(int early);       // bn_act.cu's copy of c_pdl_early
void set_pdl(int on) {
  g_pdl = on;
  const int trig = on == 2 ? 1 : (on >= 3 ? 2 : 0);
  cudaMemcpyToSymbol(c_pdl_trigger, &trig, sizeof(int));
  pdl_early_bn(trig);
}

// `w` is the bf16 weight matrix [n_total][KB*64] (fwd / gemm / stem) or [Cout][R*S*Cin] (dgrad modes).
// `a_matrix`: the A operand for the TMA-A modes (2-D matrix [M][a_cols], or the NHWC tensor in tile modes).
cudaError_t launch_conv_gemm(int mode, const ConvArgs& a_in, const void* w, int w_rows, int w_cols, int n_total,
                             const void* a_matrix, int a_cols, cudaStream_t stream) {
  ConvArgs a = a_in;
  CUtensorMap tmB;
  TmaSet tmA;
  const int bn = (n_total % 128 == 0) ? 128 : 64;
  if (n_total % 64 != 0) return cudaErrorInvalidValue;
  const int es = a.fp8 ? 1 : 2;                   // operand element size (bytes)
  const uint32_t kcols = 128u / es;               // elements in a 128-byte swizzle row
  if (a.fp8) {
    // fp8 operands run on the deep-ring kernel only (GEMM / stride-1-or-2 tile modes), with 128-element k-blocks
    if (!(mode == kConvGemm || mode == kConvTileFwd || mode == kConvTileDgrad || mode == kConvGemmDgrad))
      return cudaErrorInvalidValue;
    if (a.fp8 == 3) {
      // MX block-scaled operands: K-major GEMM, one 128-wide single-CTA tile per row block (scale atoms are per 128 rows)
      if (mode != kConvGemm || n_total % 128 != 0 || a.sfa == nullptr || a.sfb == nullptr) return cudaErrorInvalidValue;
      a.variant = kVarDeep | (2 << 4);
    } else if (a.deq_a == nullptr || a.deq_b == nullptr) {
      return cudaErrorInvalidValue;
    }
    if (mode_b_mn(mode) && n_total % 128 != 0) return cudaErrorInvalidValue;    // MN-major boxes are 128 columns wide
    if ((a.variant & 0xf) == 0) a.variant = kVarDeep;
    if ((a.variant & 0xf) != kVarDeep) return cudaErrorInvalidValue;
  }
  CUtensorMap tmBh;                 // K-major weights: half-height box (one CTA's share of a multicast pair)
  bool have_half = false;
  if (mode_b_mn(mode)) {
    if (!make_map_2d(&tmB, w, w_rows, w_cols, w_cols, kcols, kcols, es)) return cudaErrorUnknown;
    tmBh = tmB;
  } else {
    if (!make_map_2d(&tmB, w, w_rows, w_cols, w_cols, kcols, bn, es)) return cudaErrorUnknown;
    have_half = make_map_2d(&tmBh, w, w_rows, w_cols, w_cols, kcols, bn / 2, es);     // pair tiles / multicast halves
    if (!have_half) tmBh = tmB;
  }
  a.ntaps = 0;
  a.outH = a.dstH; a.outW = a.dstW; a.out_stride = 1; a.out_pa = 0; a.out_pb = 0;
  const int want_zfill = a.zfill;
  a.zfill = 0;
  int m_tiles = (a.M + kBlockM - 1) / kBlockM;
  const bool stats = a.sum != nullptr;
  const WeightDesc wd{w, w_rows, w_cols, es};
  auto dispatch = [&](const ConvArgs& args, int tiles) -> cudaError_t {
    switch (mode) {
      case kConvFwd: return launch_fwd_mode<kConvFwd>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      case kConvDgrad: return launch_fwd_mode<kConvDgrad>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      case kConvGemm: return launch_fwd_mode<kConvGemm>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      case kConvStem: return launch_fwd_mode<kConvStem>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      case kConvTileFwd: return launch_fwd_mode<kConvTileFwd>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      case kConvTileDgrad: return launch_fwd_mode<kConvTileDgrad>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      case kConvGemmDgrad: return launch_fwd_mode<kConvGemmDgrad>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      case kConvStemTma: return launch_fwd_mode<kConvStemTma>(wd, tmB, tmBh, have_half, tmA, args, n_total, tiles, stats, stream);
      default: return cudaErrorInvalidValue;
    }
  };
  if (mode == kConvGemm || mode == kConvGemmDgrad) {
    if (!make_map_2d(&tmA.m[0], a_matrix, a.M, a_cols, a_cols, kcols, 128, es)) return cudaErrorUnknown;
    for (int i = 1; i < 4; ++i) tmA.m[i] = tmA.m[0];
    return dispatch(a, m_tiles);
  }
  if (!mode_tile(mode)) {
    for (int i = 0; i < 4; ++i) tmA.m[i] = tmB;
    return dispatch(a, m_tiles);
  }
  // ---------------- tile modes ----------------
  if (a.tw * a.th * a.tn > 128 || a.tw < 1 || a.th < 1 || a.tn < 1) return cudaErrorInvalidValue;
  const int st = a.stride;
  auto set_tiles = [&](ConvArgs& x) {
    x.tiles_w = (x.dstW + x.tw - 1) / x.tw;
    x.tiles_h = (x.dstH + x.th - 1) / x.th;
    return x.tiles_w * x.tiles_h * ((x.batch + x.tn - 1) / x.tn);
  };
  if (mode == kConvStemTma) {
    // a.srcH x a.srcW is the PADDED image; the k-block geometry follows from the filter width (see stem_geometry)
    if (a.S > 16) return cudaErrorInvalidValue;
    const int SP = stem_sp(a.S), RPK = 16 / SP;
    if (!make_map_stem5d(&tmA.m[0], a_matrix, a.batch, a.srcH, a.srcW, a.dstH, a.dstW, st, RPK, a.KB, a.tw, a.th, a.tn))
      return cudaErrorInvalidValue;
    for (int i = 1; i < 4; ++i) tmA.m[i] = tmA.m[0];
    a.cchunks = RPK;
    return dispatch(a, set_tiles(a));
  }
  if (st == 1) {
    if (!make_map_nhwc(&tmA.m[0], a_matrix, a.batch, a.srcH, a.srcW, a.srcC, a.tw, a.th, a.tn, es)) return cudaErrorUnknown;
    for (int i = 1; i < 4; ++i) tmA.m[i] = tmA.m[0];
    return dispatch(a, set_tiles(a));
  }
  if (st != 2 || a.R * a.S > 16) return cudaErrorInvalidValue;
  if (mode == kConvTileFwd) {
    // stride-2 forward: every filter tap reads a dense box of one of the four 2x2 phase sub-images of x
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb)
        if (!make_map_nhwc_phase(&tmA.m[pa * 2 + pb], a_matrix, a.batch, a.srcH, a.srcW, a.srcC, 2, pa, pb, a.tw, a.th,
                                 a.tn, es))
          return cudaErrorUnknown;
    a.ntaps = a.R * a.S;
    for (int r = 0; r < a.R; ++r)
      for (int sx = 0; sx < a.S; ++sx) {
        const int t = r * a.S + sx;
        const int oh = r * a.dil - a.pad, ow = sx * a.dil - a.pad_w;
        const int pa = ((oh % 2) + 2) % 2, pb = ((ow % 2) + 2) % 2;
        a.tap_dh[t] = static_cast<signed char>(floor_div(oh, 2));
        a.tap_dw[t] = static_cast<signed char>(floor_div(ow, 2));
        a.tap_map[t] = static_cast<signed char>(pa * 2 + pb);
        a.tap_widx[t] = static_cast<signed char>(t);
      }
    return dispatch(a, set_tiles(a));
  }
  // stride-2 data gradient: four stride-1 problems, one per output phase (pa, pb) of dx; each is a conv of dy
  // with the subset of taps of matching parity, scattered to dx[2i+pa][2j+pb].  Phases with no tap stay as the
  // caller initialised them (the Python wrapper zero-fills dx when such phases exist).
  if (!make_map_nhwc(&tmA.m[0], a_matrix, a.batch, a.srcH, a.srcW, a.srcC, a.tw, a.th, a.tn, es)) return cudaErrorUnknown;
  for (int i = 1; i < 4; ++i) tmA.m[i] = tmA.m[0];
  const int fullH = a.dstH, fullW = a.dstW;
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb) {
      ConvArgs p = a;
      p.dstH = (fullH - pa + 1) / 2;
      p.dstW = (fullW - pb + 1) / 2;
      if (p.dstH <= 0 || p.dstW <= 0) continue;
      p.outH = fullH; p.outW = fullW; p.out_stride = 2; p.out_pa = pa; p.out_pb = pb;
      int nt = 0;
      for (int r = 0; r < a.R; ++r) {
        const int th_ = pa + a.pad - r * a.dil;
        if (((th_ % 2) + 2) % 2 != 0) continue;
        for (int sx = 0; sx < a.S; ++sx) {
          const int tw_ = pb + a.pad_w - sx * a.dil;
          if (((tw_ % 2) + 2) % 2 != 0) continue;
          p.tap_dh[nt] = static_cast<signed char>(floor_div(th_, 2));
          p.tap_dw[nt] = static_cast<signed char>(floor_div(tw_, 2));
          p.tap_map[nt] = 0;
          p.tap_widx[nt] = static_cast<signed char>(r * a.S + sx);
          ++nt;
        }
      }
      if (nt == 0) continue;
      p.zfill = (want_zfill && pa == 0 && pb == 0 && fullH % 2 == 0 && fullW % 2 == 0) ? 1 : 0;
      p.ntaps = nt;
      p.KB = nt * a.cchunks;
      cudaError_t e = dispatch(p, set_tiles(p));
      if (e != cudaSuccess) return e;
    }
  return cudaSuccess;
}

cudaError_t launch_conv_wgrad(const WgradArgs& a_in, const void* dy, const void* x_matrix, int splits,
                              cudaStream_t stream) {
  WgradArgs a = a_in;
  CUtensorMap tmDy;
  TmaSet tmX;
  if (a.mode == kConvTileFwd || a.mode == kConvStemTma) {
    if (a.tw * a.th * a.tn > 64 || a.tw < 1 || a.th < 1 || a.tn < 1) return cudaErrorInvalidValue;
    if (!make_map_nhwc(&tmDy, dy, a.batch, a.P, a.Q, a.dy_ld, a.tw, a.th, a.tn)) return cudaErrorUnknown;
    if (a.mode == kConvStemTma) {
      // a.H x a.W is the PADDED image, a.cchunks = SP, a.ncols = KB * 64
      if (a.cchunks < 4) return cudaErrorInvalidValue;
      const int RPK = 16 / a.cchunks;
      if (!make_map_stem5d(&tmX.m[0], x_matrix, a.batch, a.H, a.W, a.P, a.Q, a.stride, RPK, a.ncols / 64, a.tw, a.th,
                           a.tn))
        return cudaErrorInvalidValue;
      for (int i = 1; i < 4; ++i) tmX.m[i] = tmX.m[0];
    } else if (a.stride == 1) {
      if (!make_map_nhwc(&tmX.m[0], x_matrix, a.batch, a.H, a.W, a.C, a.tw, a.th, a.tn)) return cudaErrorUnknown;
      for (int i = 1; i < 4; ++i) tmX.m[i] = tmX.m[0];
    } else if (a.stride == 2) {
      for (int pa = 0; pa < 2; ++pa)
        for (int pb = 0; pb < 2; ++pb)
          if (!make_map_nhwc_phase(&tmX.m[pa * 2 + pb], x_matrix, a.batch, a.H, a.W, a.C, 2, pa, pb, a.tw, a.th, a.tn))
            return cudaErrorUnknown;
    } else {
      return cudaErrorInvalidValue;
    }
    a.tiles_w = (a.Q + a.tw - 1) / a.tw;
    a.tiles_h = (a.P + a.th - 1) / a.th;
    a.total_kb = a.tiles_w * a.tiles_h * ((a.batch + a.tn - 1) / a.tn);
  } else {
    if (!make_map_2d(&tmDy, dy, a.M, a.dy_ld, a.dy_ld, 64, 64)) return cudaErrorUnknown;
    if (a.mode == kConvGemm) {
      if (!make_map_2d(&tmX.m[0], x_matrix, a.M, a.ncols, a.ncols, 64, 64)) return cudaErrorUnknown;
    } else {
      tmX.m[0] = tmDy;
    }
    for (int i = 1; i < 4; ++i) tmX.m[i] = tmX.m[0];
    a.total_kb = (a.M + 63) / 64;
  }
  if (splits < 1) splits = 1;
  if (splits > a.total_kb) splits = a.total_kb;
  a.kb_per_split = (a.total_kb + splits - 1) / splits;
  splits = (a.total_kb + a.kb_per_split - 1) / a.kb_per_split;
  // TMA-fed wgrad tiles run best with a 2-deep ring (measured: 3 CTAs/SM beat a deeper pipeline); the cp.async
  // gather modes need depth > kLag
  const bool tma_fed = (a.mode == kConvGemm || a.mode == kConvTileFwd || a.mode == kConvStemTma);
  a.stages = tma_fed ? 2 : kWgMaxStages;
  if (g_force_stages > 0 && g_force_stages <= kWgMaxStages && tma_fed) a.stages = g_force_stages;
  if (a.stages > a.kb_per_split) a.stages = a.kb_per_split;
  if (a.stages < 1) a.stages = 1;
  const int smem = wg_smem_bytes(a.stages);
  // narrow outputs (the last 128-channel tile would be at most half full): swap operand roles, 64-channel tiles
  const int co_rem = a.Cout % 128;
  const bool swap = g_wgrad_swap && co_rem > 0 && co_rem <= 64;
  dim3 grid((a.ncols + 127) / 128, swap ? (a.Cout + 63) / 64 : (a.Cout + 127) / 128, splits);
  static bool configured[16] = {};
#define DDL_WG(MODE)                                                                                          \
  do {                                                                                                        \
    auto kern = swap ? conv_wgrad_kernel<MODE, true> : conv_wgrad_kernel<MODE, false>;                        \
    if (!configured[MODE * 2 + (swap ? 1 : 0)]) {                                                             \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                 \
                                           wg_smem_bytes(kWgMaxStages));                                      \
      if (e != cudaSuccess) return e;                                                                         \
      configured[MODE * 2 + (swap ? 1 : 0)] = true;                                                           \
    }                                                                                                         \
    kern<<<grid, kThreads, smem, stream>>>(tmDy, tmX, a);                                                     \
  } while (0)
  switch (a.mode) {
    case kConvFwd: DDL_WG(kConvFwd); break;
    case kConvGemm: DDL_WG(kConvGemm); break;
    case kConvStem: DDL_WG(kConvStem); break;
    case kConvTileFwd: DDL_WG(kConvTileFwd); break;
    case kConvStemTma: DDL_WG(kConvStemTma); break;
    default: return cudaErrorInvalidValue;
  }
#undef DDL_WG
  return cudaGetLastError();
}

}  // namespace ddl
