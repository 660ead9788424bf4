// see step_launcher.h
#include "step_launcher.h"

#include <algorithm>
#include <mutex>
#include <stdexcept>
#include <string>

namespace ddl {

namespace {

void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("StepLauncher: ") + what + ": " + cudaGetErrorString(e));
}

struct WgradState {
  std::mutex mu;
  cudaStream_t side = nullptr;
  bool pending = false;
  cudaEvent_t ev = nullptr;
} g_wgrad;

}  // namespace

void wgrad_note(cudaStream_t side) {
  std::lock_guard<std::mutex> lk(g_wgrad.mu);
  g_wgrad.side = side;
  g_wgrad.pending = true;
}

bool wgrad_join(cudaStream_t consumer) {
  std::lock_guard<std::mutex> lk(g_wgrad.mu);
  if (!g_wgrad.pending) return false;      // also keeps graph capture legal: never wait on a stream outside the capture
  if (g_wgrad.ev == nullptr) ck(cudaEventCreateWithFlags(&g_wgrad.ev, cudaEventDisableTiming), "event create");
  ck(cudaEventRecord(g_wgrad.ev, g_wgrad.side), "record side stream");
  ck(cudaStreamWaitEvent(consumer, g_wgrad.ev, 0), "join side stream");
  g_wgrad.pending = false;
  return true;
}

StepLauncher::StepLauncher(StepLauncherConfig cfg, StreamFn stream_fn, HyperFn hyper_fn)
    : c_(std::move(cfg)), stream_fn_(std::move(stream_fn)), hyper_fn_(std::move(hyper_fn)) {
  const size_t nb = c_.bucket_start.size();
  if (nb == 0 || c_.bucket_numel.size() != nb || c_.bucket_param_count.size() != nb)
    throw std::invalid_argument("StepLauncher: inconsistent bucket plan");
  for (int32_t b : c_.param_bucket)
    if (b < 0 || static_cast<size_t>(b) >= nb) throw std::invalid_argument("StepLauncher: parameter bucket out of range");
  pending_ = c_.bucket_param_count;
  seen_.assign(c_.param_bucket.size(), 0);
}

StepLauncher::~StepLauncher() {
  for (cudaEvent_t e : events_) cudaEventDestroy(e);
}

cudaEvent_t StepLauncher::next_event() {
  // a small ring: an event may be re-recorded once the wait that consumed its previous record has been ENQUEUED
  constexpr size_t kRing = 32;
  if (events_.size() < kRing) {
    cudaEvent_t e;
    ck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event create");
    events_.push_back(e);
    return e;
  }
  cudaEvent_t e = events_[ev_next_];
  ev_next_ = (ev_next_ + 1) % kRing;
  return e;
}

void StepLauncher::launch(int b) {
  if (!have_stream_) {
    cur_ = stream_fn_();
    have_stream_ = true;
  }
  cudaStream_t stream = c_.comm_stream ? c_.comm_stream : cur_;
  if (c_.comm_stream) {
    cudaEvent_t e = next_event();
    ck(cudaEventRecord(e, cur_), "record compute stream");
    ck(cudaStreamWaitEvent(stream, e, 0), "join compute stream");
  }
  wgrad_join(stream);
  if (!hyper_uploaded_) {
    hyper_fn_(stream);
    hyper_uploaded_ = true;
  }
  const int64_t start = c_.bucket_start[b], numel = c_.bucket_numel[b];
  const int nb = static_cast<int>(c_.bucket_start.size());
  int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(c_.comm_blocks, (numel / c_.world + 2047) / 2048));
  if (c_.world == 1) {
    const int64_t wide = std::max<int64_t>(blocks, std::min<int64_t>(4LL * c_.sms, numel / 2048 + 1));
    ck(launch_fused_sgd_local(c_.W + start, c_.G + start, c_.M + start, static_cast<char*>(c_.Wb) + start * 2, c_.hyper_dev,
                              numel, static_cast<int>(wide), stream), "fused_sgd_local");
  } else {
    const bool tail = scalars_pending_ && b == nb - 1;
    const bool oneshot = numel * 4 <= c_.oneshot_bytes;
    if (oneshot) blocks = std::max<int64_t>(1, std::min<int64_t>(c_.comm_blocks, (numel + 2047) / 2048));
    BucketArgs a{};
    a.start = start;
    a.numel = numel;
    a.momentum = c_.M;
    a.hyper = c_.hyper_dev;
    a.channel = 0;
    a.scalar_off = tail ? c_.scalars_off : 0;
    a.scalar_out = tail ? c_.scalars_out : nullptr;
    a.oneshot = oneshot ? 1 : 0;
    a.closing = b == nb - 1 ? 1 : 0;
    ck(launch_fused_allreduce_sgd(c_.ctx, a, c_.use_mc, c_.wire_bf16, static_cast<int>(blocks), stream), "fused_allreduce_sgd");
  }
  ++launches_;
}

int StepLauncher::on_ready(int idx) {
  if (idx < 0 || static_cast<size_t>(idx) >= seen_.size()) throw std::out_of_range("StepLauncher: parameter index");
  if (seen_[idx]) return 0;
  seen_[idx] = 1;
  --pending_[c_.param_bucket[idx]];
  if (hold_) return 0;
  const int nb = static_cast<int>(c_.bucket_start.size());
  int n = 0;
  while (next_ < nb && pending_[next_] == 0) {     // plan order: every rank issues the same kernel sequence
    launch(next_);
    ++next_;
    ++n;
  }
  return n;
}

int StepLauncher::finish() {
  const int nb = static_cast<int>(c_.bucket_start.size());
  int n = 0;
  while (next_ < nb) {
    launch(next_);
    ++next_;
    ++n;
  }
  if (c_.comm_stream) {
    if (!have_stream_) {
      cur_ = stream_fn_();
      have_stream_ = true;
    }
    cudaEvent_t e = next_event();
    ck(cudaEventRecord(e, c_.comm_stream), "record communication stream");
    ck(cudaStreamWaitEvent(cur_, e, 0), "join communication stream");
  }
  return n;
}

void StepLauncher::reset() {
  pending_ = c_.bucket_param_count;
  std::fill(seen_.begin(), seen_.end(), 0);
  next_ = 0;
  hyper_uploaded_ = false;
  have_stream_ = false;
  scalars_pending_ = false;
}

}  // namespace ddl
