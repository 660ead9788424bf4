// Host/device shared definitions for the NVLink communication kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ddl {

constexpr int kMaxCommBlocks = 64;   // upper bound on the grid of any comm kernel
constexpr int kCommChannels = 8;     // independent barrier channels in the signal pad
constexpr int kCommMaxWorld = 8;

// Byte size of the signal pad every rank reserves at `flag_off` in its arena.
constexpr uint64_t kSignalPadBytes = uint64_t(kCommChannels) * kMaxCommBlocks * kCommMaxWorld * 4u;

struct CommCtx {
  uint64_t peer_base[kCommMaxWorld];  // VA of every rank's arena in THIS process
  uint64_t mc_base;                   // multicast VA aliasing all arenas (0 = unavailable)
  uint64_t flag_off;                  // signal pad
  uint64_t grad_off;                  // fp32 gradient accumulators   [total_elems]
  uint64_t weight_off;                // fp32 master weights          [total_elems]
  uint64_t wbf16_off;                 // bf16 compute copy of weights [total_elems]
  uint64_t stage_off;                 // bf16 wire staging            [total_elems] (optional)
  uint32_t* epoch_ctr;                // local: [kCommChannels][kMaxCommBlocks]
  uint32_t* error_flag;               // local: set to 1+peer on barrier timeout
  uint64_t timeout_ns;
  uint32_t debug_skew_ns;             // test hook: block b idles (b % 4) * skew ns before its data phase / prologue
  int rank;
  int world;
};

struct SgdHyper {   // lives in device memory so captured graphs see new values each replay
  float lr;
  float momentum;
  float dampening;
  float weight_decay;
  float grad_scale;   // 1/world (gradient averaging), times any loss-scale inverse
  int nesterov;
  int first_step;     // momentum buffers not initialised yet (torch: buf = grad)
  int pad;
};

struct BucketArgs {
  int64_t start;          // element offset of the bucket in the arena
  int64_t numel;          // padded bucket length (multiple of world*8)
  float* momentum;        // local momentum buffer base (indexed by arena element offset)
  const SgdHyper* hyper;  // device pointer
  int channel;
  // scalar piggy-back: kScalarSlots fp32 values at byte offset `scalar_off` of every replica are
  // averaged across ranks into `scalar_out` (local) by block 0 of this launch; scalar_off == 0 disables it.
  uint64_t scalar_off;
  float* scalar_out;
  int oneshot;            // small bucket: every rank reduces and updates the whole bucket (no broadcast phase)
  int closing;            // end the kernel with a cross-rank barrier (the LAST bucket kernel of a step must: it is what
                          // makes every peer's weight stores / accumulator clears of the whole step visible locally)
};
constexpr int kScalarSlots = 64;

cudaError_t launch_fused_sgd_local(float* w, float* g, float* m, void* wb, const SgdHyper* hp, int64_t numel,
                                   int blocks, cudaStream_t stream);
cudaError_t launch_fused_allreduce_sgd(const CommCtx& c, const BucketArgs& b, bool use_mc, bool wire_bf16,
                                       int blocks, cudaStream_t stream);
cudaError_t launch_allreduce(const CommCtx& c, int channel, uint64_t off, int64_t numel, bool bf16, float scale,
                             bool use_mc, bool oneshot, int blocks, cudaStream_t stream);
cudaError_t launch_broadcast(const CommCtx& c, int channel, uint64_t off, int64_t bytes, int root, bool use_mc,
                             int blocks, cudaStream_t stream);
cudaError_t launch_barrier(const CommCtx& c, int channel, cudaStream_t stream);
cudaError_t launch_allgather_slices(const CommCtx& c, int channel, const float* src, uint64_t dst_off,
                                    int64_t start, int64_t numel, bool use_mc, int blocks, cudaStream_t stream);

}  // namespace ddl
