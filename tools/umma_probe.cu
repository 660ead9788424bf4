// Micro-experiment: which shared-memory addresses does tcgen05.mma read for a 128B-swizzled K-major operand whose
// descriptor start is NOT 1024-byte aligned and / or whose 8-row groups are NOT 1024 bytes apart?
//
// Motivation (3x3 convolutions): if an A operand can be a shifted / strided WINDOW of a haloed activation tile that
// TMA wrote once ([th+2][tw+2] pixels x 64 channels, 128 B per pixel), all nine filter taps can run their MMAs off
// ONE shared-memory tile: pixel row-groups of 8 (tw = 8) are (tw+2)*128 = 1280 B apart (SBO = 1280) and tap (r, s)
// starts r*(tw+2)+s rows into the tile.  That only works if the tensor core applies the 128B swizzle as a function
// of the ABSOLUTE shared-memory address bits (as TMA does when it writes), or if the descriptor's base_offset field
// can express the phase.  This program loads X[R][64] bf16 with one TMA box, runs D = A(window) * B^T for several
// (row shift, SBO, base_offset) settings and compares every result with the host reference.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kRows = 256;      // rows of X in shared memory (128 B each)
constexpr int kN = 64;          // B rows (output columns)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(base_off & 7u) << 49;
  d |= 2ull << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
                                                    float* out, int shift_rows, int sbo_bytes, int base_off) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                       // kRows * 128
  uint8_t* sB = smem + kRows * 128;         // kN * 128
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kN * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(done)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"((kRows + kN) * 128));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(sX)), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(sB)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // instruction descriptor: bf16 x bf16 -> f32, M = 128, N = 64, both K-major
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = make_desc(smem_u32(sX) + shift_rows * 128 + k * 32, 0, sbo_bytes, base_off);
      const uint64_t db = make_desc(smem_u32(sB) + k * 32, 0, 1024, 0);
      const uint32_t acc = k != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(done)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(done)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = threadIdx.x;      // warp w reads TMEM lanes 32w .. 32w+31
  for (int c0 = 0; c0 < kN; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[row * kN + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem));
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  CK(cudaSetDevice(0));
  std::vector<__nv_bfloat16> hX(kRows * 64), hB(kN * 64);
  std::vector<float> fX(kRows * 64), fB(kN * 64);
  srand(1);
  for (size_t i = 0; i < hX.size(); ++i) { float v = bf((rand() % 2001 - 1000) / 500.f); hX[i] = __float2bfloat16(v); fX[i] = v; }
  for (size_t i = 0; i < hB.size(); ++i) { float v = bf((rand() % 2001 - 1000) / 500.f); hB[i] = __float2bfloat16(v); fB[i] = v; }
  __nv_bfloat16 *dX, *dB;
  float* dOut;
  CK(cudaMalloc(&dX, hX.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dOut, 128 * kN * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmX, tmB;
  auto mk = [&](CUtensorMap* m, void* base, int rows) {
    cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(rows)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  mk(&tmX, dX, kRows);
  mk(&tmB, dB, kN);
  const int smem = (kRows + kN) * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Case { int shift, sbo, bo; const char* what; };
  const Case cases[] = {
      {0, 1024, 0, "aligned baseline"},
      {1, 1024, 0, "start +1 row, base_offset 0"},
      {1, 1024, 1, "start +1 row, base_offset 1"},
      {3, 1024, 0, "start +3 rows, base_offset 0"},
      {3, 1024, 3, "start +3 rows, base_offset 3"},
      {8, 1024, 0, "start +8 rows (aligned again)"},
      {0, 1280, 0, "groups 10 rows apart (SBO 1280), base_offset 0"},
      {1, 1280, 0, "SBO 1280, start +1 row, base_offset 0"},
      {1, 1280, 1, "SBO 1280, start +1 row, base_offset 1"},
      {11, 1280, 0, "SBO 1280, start +11 rows (tap r=1,s=1), base_offset 0"},
      {11, 1280, 3, "SBO 1280, start +11 rows, base_offset 3"},
      {0, 1152, 0, "groups 9 rows apart (SBO 1152)"},
  };
  std::vector<float> got(128 * kN);
  for (const Case& c : cases) {
    CK(cudaMemset(dOut, 0, 128 * kN * 4));
    probe_kernel<<<1, 128, smem>>>(tmX, tmB, dOut, c.shift, c.sbo, c.bo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-60s LAUNCH ERROR %s\n", c.what, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(got.data(), dOut, got.size() * 4, cudaMemcpyDeviceToHost));
    // reference A: MMA row i = group g = i / 8, j = i % 8  ->  X row  shift + g * (sbo / 128) + j
    double maxerr = 0;
    int bad_rows = 0;
    for (int i = 0; i < 128; ++i) {
      const int xr = c.shift + (i / 8) * (c.sbo / 128) + (i % 8);
      double rowerr = 0;
      for (int n = 0; n < kN; ++n) {
        double acc = 0;
        for (int k = 0; k < 64; ++k) acc += static_cast<double>(fX[xr * 64 + k]) * fB[n * 64 + k];
        rowerr = fmax(rowerr, fabs(acc - got[i * kN + n]));
      }
      if (rowerr > 1e-2) ++bad_rows;
      maxerr = fmax(maxerr, rowerr);
    }
    printf("%-60s shift=%2d sbo=%4d base_off=%d : %s (max err %.3g, %d/128 rows wrong)\n", c.what, c.shift, c.sbo, c.bo,
           bad_rows == 0 ? "MATCHES absolute-address window" : "differs", maxerr, bad_rows);
  }
  return 0;
}
