// Static gradient-bucket planner.
//
// Replaces Horovod's runtime negotiation + tensor-fusion buffer (background
// thread, rank-0 coordinator, 64 MB fusion threshold, 5 ms cycle).  The plan is a pure
// function of the parameter sizes in gradient-ready order, so every rank computes the
// identical layout with no communication; `hash` lets ranks cross-check that once.
//
// Layout rules
//   * parameters are laid out in ready order (reverse execution order) in ONE flat arena;
//   * every parameter starts on an `align_elems` boundary (TMA base alignment, 16 B vector
//     access, red.v4 epilogues);
//   * the first bucket is small (`first_cap_elems`) so the first fused allreduce can start
//     early in backward; later buckets use `cap_elems`;
//   * every bucket's length is padded to a multiple of `slice_elems` so it splits into
//     world_size equal, 16 B-aligned slices for the two-shot (reduce-scatter/all-gather) kernel.
#include "host_runtime.h"

#include <cstdint>
#include <stdexcept>

namespace ddl {

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static inline uint64_t fnv1a(uint64_t h, uint64_t v) {
  for (int i = 0; i < 8; ++i) {
    h ^= (v >> (8 * i)) & 0xffu;
    h *= 1099511628211ull;
  }
  return h;
}

BucketPlan plan_buckets(const std::vector<int64_t>& numels, int64_t first_cap_elems,
                        int64_t cap_elems, int64_t align_elems, int64_t slice_elems) {
  if (align_elems <= 0 || slice_elems <= 0 || cap_elems <= 0 || first_cap_elems <= 0)
    throw std::invalid_argument("plan_buckets: caps and alignments must be positive");
  if (slice_elems % align_elems != 0)
    throw std::invalid_argument("plan_buckets: slice_elems must be a multiple of align_elems");
  BucketPlan p;
  p.param_bucket.resize(numels.size());
  p.param_offset.resize(numels.size());
  int64_t cursor = 0;        // arena cursor (elements)
  int64_t bucket_begin = 0;  // start of the open bucket
  int64_t used = 0;          // elements used in the open bucket
  auto close_bucket = [&]() {
    if (used == 0) return;
    int64_t len = round_up(cursor - bucket_begin, slice_elems);
    p.bucket_start.push_back(bucket_begin);
    p.bucket_numel.push_back(len);
    cursor = bucket_begin + len;
    bucket_begin = cursor;
    used = 0;
  };
  for (size_t i = 0; i < numels.size(); ++i) {
    if (numels[i] < 0) throw std::invalid_argument("plan_buckets: negative numel");
    int64_t cap = p.bucket_start.empty() ? first_cap_elems : cap_elems;
    int64_t need = round_up(numels[i], align_elems);
    if (used > 0 && used + need > cap) {
      close_bucket();
    }
    p.param_bucket[i] = static_cast<int32_t>(p.bucket_start.size());
    p.param_offset[i] = cursor;
    cursor += need;
    used += need;
  }
  close_bucket();
  p.total_elems = cursor;
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < numels.size(); ++i) {
    h = fnv1a(h, static_cast<uint64_t>(numels[i]));
    h = fnv1a(h, static_cast<uint64_t>(p.param_offset[i]));
    h = fnv1a(h, static_cast<uint64_t>(p.param_bucket[i]));
  }
  for (size_t b = 0; b < p.bucket_start.size(); ++b) {
    h = fnv1a(h, static_cast<uint64_t>(p.bucket_start[b]));
    h = fnv1a(h, static_cast<uint64_t>(p.bucket_numel[b]));
  }
  p.hash = h;
  // last parameter of each bucket in ready order == the one whose gradient closes the bucket
  p.bucket_last_param.assign(p.bucket_start.size(), -1);
  p.bucket_param_count.assign(p.bucket_start.size(), 0);
  for (size_t i = 0; i < numels.size(); ++i) {
    p.bucket_last_param[p.param_bucket[i]] = static_cast<int32_t>(i);
    p.bucket_param_count[p.param_bucket[i]] += 1;
  }
  return p;
}

}  // namespace ddl
