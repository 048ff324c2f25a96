// Symmetric heap over NVLink: every rank owns `bytes` of HBM created with the
// CUDA virtual-memory API; all ranks map all heaps into one contiguous VA
// window (peer r at window + r*bytes) and, when the platform supports NVLS,
// bind them to one multicast object whose VA aliases offset X of *every*
// rank's heap (multimem.ld_reduce / multimem.st / multimem.red target).
//
// Role in the reference: the device memory banks + the network between
// CCLOs.  "Transport bring-up" (configure_vnx/configure_tcp/configure_cyt_rdma,
// driver/utils/accl_network_utils/accl_network_utils.cpp:130-391) collapses to
// mapping peers' memory; `xclbin_scan` bank discovery collapses to the
// capability probe in topology().
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "accl/bootstrap.hpp"

namespace accl {
namespace cuda {

struct Topology {
  int device = -1;
  int sm_count = 0;
  int device_count = 0;
  bool multicast_supported = false;  // CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
  bool vmm_posix_fd = false;         // handle type POSIX fd supported
  bool stream_memops = true;
  size_t total_mem = 0;
  std::string name;
  int cc_major = 0, cc_minor = 0;
  std::string describe() const;
};

Topology probe_topology(int device);

class SymHeap {
public:
  // Collective over `oob`.  bytes is rounded up to the mapping granularity.
  SymHeap(Oob &oob, int device, size_t bytes, bool want_multicast);
  ~SymHeap();
  SymHeap(const SymHeap &) = delete;
  SymHeap &operator=(const SymHeap &) = delete;

  int rank() const { return rank_; }
  int world() const { return world_; }
  int device() const { return device_; }
  size_t bytes() const { return bytes_; }
  char *base(int r) const { return window_ + static_cast<size_t>(r) * bytes_; }
  char *local() const { return base(rank_); }
  char *window() const { return window_; }
  // nullptr when NVLS is unavailable (single GPU shared by ranks, no switch…)
  char *mc_base() const { return mc_va_; }
  bool has_multicast() const { return mc_va_ != nullptr; }
  const std::string &multicast_note() const { return mc_note_; }
  bool contains(const void *p) const {
    const char *c = static_cast<const char *>(p);
    return c >= local() && c < local() + bytes_;
  }
  size_t offset_of(const void *p) const {
    return static_cast<size_t>(static_cast<const char *>(p) - local());
  }

private:
  int rank_, world_, device_;
  size_t bytes_ = 0;
  char *window_ = nullptr;
  char *mc_va_ = nullptr;
  std::string mc_note_;
  std::vector<unsigned long long> handles_; // CUmemGenericAllocationHandle per rank
  unsigned long long mc_handle_ = 0;
  bool mc_bound_ = false;
  bool share_by_value_ = false;
};

} // namespace cuda
} // namespace accl
