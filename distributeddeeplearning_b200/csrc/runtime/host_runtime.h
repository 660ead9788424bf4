// Host-side runtime of b200-ddl (pure C++; no torch headers, importable without a GPU).
//
//   bucket_plan.cpp  static gradient bucket planner           (replaces Horovod N1 negotiation)
//   fd_channel.cpp   SCM_RIGHTS all-to-all over abstract unix sockets (VMM handle exchange)
//   symm_arena.cpp   CUDA VMM symmetric arena + NVLS multicast binding (replaces NCCL N3 setup)
//   watchdog.cpp     rank supervisor helpers: bounded waits, fault injection
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ddl {

struct BucketPlan {
  std::vector<int32_t> param_bucket;       // bucket id of each parameter (ready order)
  std::vector<int64_t> param_offset;       // element offset of each parameter in the arena
  std::vector<int64_t> bucket_start;       // element offset of each bucket
  std::vector<int64_t> bucket_numel;       // padded length of each bucket
  std::vector<int32_t> bucket_last_param;  // the parameter whose gradient closes the bucket
  std::vector<int32_t> bucket_param_count;
  int64_t total_elems = 0;
  uint64_t hash = 0;
};

BucketPlan plan_buckets(const std::vector<int64_t>& numels, int64_t first_cap_elems,
                        int64_t cap_elems, int64_t align_elems, int64_t slice_elems);

// ---- fd exchange ------------------------------------------------------------------------
// Every rank contributes one file descriptor and receives one from every peer
// (result[rank] == -1 for self).  `session` must be unique per job and per exchange round.
std::vector<int> exchange_fds(int rank, int world, int my_fd, const std::string& session,
                              int timeout_ms);
// Root sends one fd to every peer; returns the received fd on peers, `fd` on root.
int broadcast_fd(int rank, int world, int root, int fd, const std::string& session, int timeout_ms);

// ---- symmetric arena ----------------------------------------------------------------------
class SymmArena {
 public:
  SymmArena(int rank, int world, int device, size_t bytes);
  ~SymmArena();
  SymmArena(const SymmArena&) = delete;
  SymmArena& operator=(const SymmArena&) = delete;

  void alloc();                                          // cuMemCreate + map + zero own memory
  void exchange(const std::string& session, int timeout_ms);  // import + map all peers
  bool multicast_supported() const;
  bool mc_create(const std::string& session, int timeout_ms);  // rank0 creates, all import + add
  bool mc_bind();                                        // after a barrier: bind + map
  void release();

  size_t bytes() const { return bytes_; }
  uint64_t local_ptr() const { return ptrs_.empty() ? 0 : ptrs_[rank_]; }
  const std::vector<uint64_t>& peer_ptrs() const { return ptrs_; }
  uint64_t mc_ptr() const { return mc_ptr_; }
  int rank() const { return rank_; }
  int world() const { return world_; }
  std::string last_error() const { return err_; }

 private:
  int rank_, world_, device_;
  size_t req_bytes_, bytes_ = 0, gran_ = 0;
  unsigned long long handle_ = 0;                 // CUmemGenericAllocationHandle (own memory)
  std::vector<unsigned long long> peer_handles_;  // imported handles
  std::vector<uint64_t> ptrs_;                    // mapped VA per rank
  unsigned long long mc_handle_ = 0;
  uint64_t mc_ptr_ = 0;
  bool mc_added_ = false;
  std::string err_;
};

// true when libcuda can be loaded (i.e. a driver is present); never throws
bool driver_available();

}  // namespace ddl
