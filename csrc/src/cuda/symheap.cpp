#include "accl/cuda/symheap.hpp"

#include <unistd.h>

#include <cstring>
#include <sstream>

#include "accl/common.hpp"
#include "accl/cuda/driver_api.hpp"

namespace accl {
namespace cuda {

std::string Topology::describe() const {
  std::ostringstream o;
  o << "device " << device << " '" << name << "' sm_" << cc_major << cc_minor
    << " SMs=" << sm_count << " mem=" << (total_mem >> 20) << "MiB"
    << " visible_devices=" << device_count
    << " multicast=" << (multicast_supported ? "yes" : "no")
    << " vmm_posix_fd=" << (vmm_posix_fd ? "yes" : "no");
  return o.str();
}

Topology probe_topology(int device) {
  Topology t;
  t.device = device;
  ACCL_CUDART(cudaGetDeviceCount(&t.device_count));
  cudaDeviceProp p{};
  ACCL_CUDART(cudaGetDeviceProperties(&p, device));
  t.name = p.name;
  t.sm_count = p.multiProcessorCount;
  t.total_mem = p.totalGlobalMem;
  t.cc_major = p.major;
  t.cc_minor = p.minor;
  auto &d = DriverApi::get();
  CUdevice dev;
  ACCL_CU(d.cuDeviceGet(&dev, device));
  int v = 0;
  if (d.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS)
    t.multicast_supported = v != 0;
  v = 0;
  if (d.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) == CUDA_SUCCESS)
    t.vmm_posix_fd = v != 0;
  return t;
}

namespace {
struct PeerInfo {
  unsigned char uuid[16];
  int mc_supported;
  int device;
  unsigned long long bytes;
};
size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }
bool all_ok(Oob &oob, bool mine) {
  auto v = oob.allgather_value<int>(mine ? 1 : 0);
  for (int x : v)
    if (!x) return false;
  return true;
}
} // namespace

SymHeap::SymHeap(Oob &oob, int device, size_t bytes, bool want_multicast)
    : rank_(oob.rank()), world_(oob.size()), device_(device) {
  auto &d = DriverApi::get();
  ACCL_CUDART(cudaSetDevice(device));
  ACCL_CUDART(cudaFree(nullptr)); // force primary context
  CUdevice cudev;
  ACCL_CU(d.cuDeviceGet(&cudev, device));
  Topology topo = probe_topology(device);
  // Handles always travel as POSIX fds (dup'ed when ranks share a process), so
  // every rank owns its own reference to every allocation and tear-down order
  // does not matter.
  const bool share_by_value = false;
  share_by_value_ = false;

  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes =
      share_by_value ? CU_MEM_HANDLE_TYPE_NONE : CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;

  size_t gran = 0;
  ACCL_CU(d.cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));

  // ---- agree on geometry and on whether NVLS is usable
  PeerInfo mine{};
  cudaDeviceProp dp{};
  ACCL_CUDART(cudaGetDeviceProperties(&dp, device));
  std::memcpy(mine.uuid, &dp.uuid, 16);
  mine.mc_supported = topo.multicast_supported ? 1 : 0;
  mine.device = device;
  mine.bytes = bytes;
  auto infos = oob.allgather_value(mine);
  bool mc = want_multicast && world_ > 1;
  size_t want = 0;
  for (int r = 0; r < world_; ++r) {
    const auto &pi = infos[static_cast<size_t>(r)];
    want = std::max<size_t>(want, pi.bytes);
    if (!pi.mc_supported) {
      if (mc) mc_note_ = "multicast not supported by rank " + std::to_string(r);
      mc = false;
    }
    for (int q = 0; q < r; ++q)
      if (std::memcmp(pi.uuid, infos[static_cast<size_t>(q)].uuid, 16) == 0) {
        if (mc) mc_note_ = "ranks share a physical GPU";
        mc = false;
      }
  }
  if (!want_multicast) mc_note_ = "disabled by configuration";
  if (world_ == 1) mc_note_ = "single rank";

  CUmulticastObjectProp mprop{};
  size_t mc_gran = 0;
  if (mc) {
    mprop.numDevices = static_cast<unsigned>(world_);
    mprop.size = round_up(want, gran);
    mprop.handleTypes = share_by_value ? 0 : CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t rec = 0;
    if (d.cuMulticastGetGranularity(&mc_gran, &mprop, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS)
      mc_gran = 0;
    if (d.cuMulticastGetGranularity(&rec, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
        rec <= want)
      mc_gran = std::max(mc_gran, rec);
    if (mc_gran == 0) {
      mc = false;
      mc_note_ = "cuMulticastGetGranularity failed";
    }
  }
  if (!all_ok(oob, mc) && mc) {
    mc = false;
    mc_note_ = "a peer cannot use multicast";
  }
  // every rank must compute the same size
  size_t g = gran;
  if (mc) g = std::max(g, mc_gran);
  g = oob.bcast_value<unsigned long long>(g, 0);
  bytes_ = round_up(want, g);

  // ---- allocate mine, share handles
  CUmemGenericAllocationHandle my_h = 0;
  ACCL_CU(d.cuMemCreate(&my_h, bytes_, &prop, 0));
  handles_.assign(static_cast<size_t>(world_), 0);
  if (share_by_value) {
    auto hs = oob.allgather_value<unsigned long long>(my_h);
    for (int r = 0; r < world_; ++r) handles_[static_cast<size_t>(r)] = hs[static_cast<size_t>(r)];
  } else {
    int fd = -1;
    ACCL_CU(d.cuMemExportToShareableHandle(&fd, my_h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    auto fds = exchange_fds(oob, fd, "heap");
    for (int r = 0; r < world_; ++r) {
      if (r == rank_) {
        handles_[static_cast<size_t>(r)] = my_h;
      } else {
        CUmemGenericAllocationHandle h = 0;
        ACCL_CU(d.cuMemImportFromShareableHandle(
            &h, reinterpret_cast<void *>(static_cast<uintptr_t>(fds[static_cast<size_t>(r)])),
            CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        handles_[static_cast<size_t>(r)] = h;
      }
      if (fds[static_cast<size_t>(r)] >= 0) ::close(fds[static_cast<size_t>(r)]);
    }
    ::close(fd);
  }

  // ---- one VA window holding every rank's heap
  CUdeviceptr va = 0;
  ACCL_CU(d.cuMemAddressReserve(&va, bytes_ * static_cast<size_t>(world_), g, 0, 0));
  window_ = reinterpret_cast<char *>(va);
  for (int r = 0; r < world_; ++r)
    ACCL_CU(d.cuMemMap(va + bytes_ * static_cast<size_t>(r), bytes_, 0,
                       handles_[static_cast<size_t>(r)], 0));
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ACCL_CU(d.cuMemSetAccess(va, bytes_ * static_cast<size_t>(world_), &acc, 1));
  ACCL_CUDART(cudaMemset(local(), 0, bytes_));
  ACCL_CUDART(cudaDeviceSynchronize());

  // ---- NVLS multicast object (rank 0 creates, everybody joins and binds)
  if (mc) {
    mprop.size = bytes_;
    CUmemGenericAllocationHandle mch = 0;
    bool ok = true;
    if (rank_ == 0) {
      CUresult r = d.cuMulticastCreate(&mch, &mprop);
      if (r != CUDA_SUCCESS) {
        ok = false;
        mc_note_ = "cuMulticastCreate: " + cu_error_string(r);
      }
    }
    ok = oob.bcast_value<int>(ok ? 1 : 0, 0) != 0;
    if (ok) {
      if (share_by_value) {
        mch = oob.bcast_value<unsigned long long>(mch, 0);
      } else {
        int fd = -1;
        if (rank_ == 0)
          ACCL_CU(d.cuMemExportToShareableHandle(&fd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        auto fds = exchange_fds(oob, fd, "mc");
        if (rank_ != 0) {
          CUresult r = d.cuMemImportFromShareableHandle(
              &mch, reinterpret_cast<void *>(static_cast<uintptr_t>(fds[0])),
              CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
          if (r != CUDA_SUCCESS) {
            ok = false;
            mc_note_ = "import multicast handle: " + cu_error_string(r);
          }
        }
        for (int f : fds)
          if (f >= 0) ::close(f);
        if (fd >= 0) ::close(fd);
      }
      if (ok) {
        CUresult r = d.cuMulticastAddDevice(mch, cudev);
        if (r != CUDA_SUCCESS) {
          ok = false;
          mc_note_ = "cuMulticastAddDevice: " + cu_error_string(r);
        }
      }
      ok = all_ok(oob, ok);
      if (ok) {
        CUresult r = d.cuMulticastBindMem(mch, 0, my_h, 0, bytes_, 0);
        if (r != CUDA_SUCCESS) {
          ok = false;
          mc_note_ = "cuMulticastBindMem: " + cu_error_string(r);
        } else {
          mc_bound_ = true;
        }
      }
      ok = all_ok(oob, ok);
      if (ok) {
        CUdeviceptr mva = 0;
        CUresult r = d.cuMemAddressReserve(&mva, bytes_, g, 0, 0);
        if (r == CUDA_SUCCESS) r = d.cuMemMap(mva, bytes_, 0, mch, 0);
        if (r == CUDA_SUCCESS) r = d.cuMemSetAccess(mva, bytes_, &acc, 1);
        if (r != CUDA_SUCCESS) {
          ok = false;
          mc_note_ = "map multicast VA: " + cu_error_string(r);
        } else {
          mc_va_ = reinterpret_cast<char *>(mva);
        }
      }
      ok = all_ok(oob, ok);
      if (!ok) mc_va_ = nullptr;
      mc_handle_ = mch;
    }
    if (mc_va_ == nullptr && mc_note_.empty()) mc_note_ = "a peer failed to join the multicast group";
  }
  oob.barrier();
  ACCL_DEBUG_LOG("SymHeap rank " << rank_ << "/" << world_ << " bytes=" << bytes_
                                 << " window=" << static_cast<void *>(window_)
                                 << " mc=" << static_cast<void *>(mc_va_) << " " << mc_note_);
}

SymHeap::~SymHeap() {
  if (!DriverApi::available()) return;
  auto &d = DriverApi::get();
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  if (mc_va_) {
    d.cuMemUnmap(reinterpret_cast<CUdeviceptr>(mc_va_), bytes_);
    d.cuMemAddressFree(reinterpret_cast<CUdeviceptr>(mc_va_), bytes_);
  }
  if (mc_handle_) {
    if (mc_bound_) {
      CUdevice dev;
      if (d.cuDeviceGet(&dev, device_) == CUDA_SUCCESS) d.cuMulticastUnbind(mc_handle_, dev, 0, bytes_);
    }
    if (!share_by_value_ || rank_ == 0) d.cuMemRelease(mc_handle_);
  }
  if (window_) {
    for (int r = 0; r < world_; ++r)
      d.cuMemUnmap(reinterpret_cast<CUdeviceptr>(window_) + bytes_ * static_cast<size_t>(r), bytes_);
    d.cuMemAddressFree(reinterpret_cast<CUdeviceptr>(window_), bytes_ * static_cast<size_t>(world_));
  }
  for (int r = 0; r < world_; ++r) {
    // imported handles are owned per process; by-value shared handles are
    // released only by their creator
    if (handles_[static_cast<size_t>(r)] == 0) continue;
    if (share_by_value_ && r != rank_) continue;
    d.cuMemRelease(handles_[static_cast<size_t>(r)]);
  }
}

} // namespace cuda
} // namespace accl
