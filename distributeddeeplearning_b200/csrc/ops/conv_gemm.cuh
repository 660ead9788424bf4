// Host/device shared argument structs for the tcgen05 implicit-GEMM convolution kernels.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddl {

enum ConvMode : int {
  kConvFwd = 0,     // A = im2col(x) gathered, rows = output pixels
  kConvDgrad = 1,   // A = transposed-conv gather of dy, rows = input pixels, B is MN-major
  kConvGemm = 2,    // A is a plain [M][K] matrix fetched by TMA (1x1 stride-1 convs, FC)
  kConvStem = 3,    // 7x7 s2 conv on a 4-channel (3+pad) image: k-block = 2 filter rows x 8 taps x 4 ch
};

struct ConvArgs {
  const __nv_bfloat16* src;   // gather source (x for fwd / stem, dy for dgrad); unused for kConvGemm
  __nv_bfloat16* out;         // [M][ldc] (NHWC activations / gradients)
  const __nv_bfloat16* add;   // optional: out = acc + add   (same layout as out)
  const float* bias;          // optional: per output channel
  float* sum;                 // optional BN statistics: sum[c]   += sum_m out[m][c]   (of the bf16-rounded value)
  float* sumsq;               //                         sumsq[c] += sum_m out[m][c]^2
  int M;                      // GEMM rows
  int KB;                     // number of 64-element k-blocks
  int ldc;                    // channels of `out` (row stride in elements)
  int srcH, srcW, srcC;       // gather-source geometry
  int dstH, dstW;             // GEMM-row geometry (M = batch * dstH * dstW)
  int R, S, stride, pad, dil;
  int cchunks;                // srcC / 64
  int relu;                   // apply ReLU in the epilogue (after bias)
  int n_valid;                // output channels that really exist (bias is read only below this)
};

struct WgradArgs {
  const __nv_bfloat16* x;     // gather source for the B operand (im2col(x)), NHWC
  float* dw;                  // fp32 accumulator [Cout][ldw]  (split-K: red.add)
  int M;                      // pixels = batch * P * Q  (GEMM reduction dim)
  int Cout;
  int dy_ld;                  // row stride (elements) of dy; >= Cout (padded FC logits)
  int ldw;                    // row stride of dw = R*S*Cin (or 256 for the stem scratch)
  int ncols;                  // valid columns of dw
  int H, W, C;                // x geometry
  int P, Q;                   // dy geometry
  int R, S, stride, pad, dil;
  int cchunks;                // C / 64
  int kb_per_split;           // 64-pixel blocks handled by one CTA
  int total_kb;               // ceil(M / 64)
  int mode;                   // kConvFwd (gather), kConvGemm (x via TMA), kConvStem
};

}  // namespace ddl
