// torch.cuda.MemPool over the symmetric heap: tensors allocated inside `with torch.cuda.use_mem_pool(pool)` live in
// the NVLink-visible heap, so collectives take them as zero-copy operands (what NCCL calls user-buffer registration;
// the reference's counterpart is allocating buffers in CCLO-visible device memory, driver/xrt/include/accl.hpp
// create_buffer).  torch's pluggable allocator wants two C symbols; they allocate from the heap of the CudaDevice
// that was attached for the CUDA device ordinal.  Every rank must allocate in the same order (the heap is symmetric):
// true for module construction / DDP bucket allocation, which is what this is for.
#include <cuda_runtime.h>

#include <map>
#include <set>
#include <mutex>
#include <unordered_map>

#include "accl/cuda/cudadevice.hpp"

namespace accl {
namespace cuda {
namespace {
std::mutex g_m;
std::map<int, CudaDevice *> g_dev;                       // CUDA ordinal -> backend that serves pool allocations
std::unordered_map<void *, std::pair<CudaDevice *, uint64_t>> g_live; // pointer -> (backend, heap offset)
// ranks as threads sharing one GPU (tests): the backend attached by the allocating thread wins over the per-device one
thread_local CudaDevice *t_dev = nullptr;
std::set<CudaDevice *> g_alive;
} // namespace

void heap_pool_attach(CudaDevice *d) {
  std::lock_guard<std::mutex> g(g_m);
  g_dev[d->device()] = d;
  g_alive.insert(d);
  t_dev = d;
}

void heap_pool_detach(CudaDevice *d) {
  std::lock_guard<std::mutex> g(g_m);
  for (auto it = g_dev.begin(); it != g_dev.end();)
    it = it->second == d ? g_dev.erase(it) : std::next(it);
  for (auto it = g_live.begin(); it != g_live.end();)
    it = it->second.first == d ? g_live.erase(it) : std::next(it);
  g_alive.erase(d);
  if (t_dev == d) t_dev = nullptr;
}

} // namespace cuda
} // namespace accl

extern "C" {

__attribute__((visibility("default"))) void *accl_heap_pool_alloc(ssize_t size, int device, cudaStream_t) {
  using namespace accl::cuda;
  std::lock_guard<std::mutex> g(g_m);
  auto it = g_dev.find(device);
  if (it == g_dev.end() || size < 0) return nullptr;
  CudaDevice *d = (t_dev && g_alive.count(t_dev) && t_dev->device() == device) ? t_dev : it->second;
  try {
    const uint64_t off = d->allocator().alloc(static_cast<size_t>(size) < 512 ? 512 : static_cast<size_t>(size), 512);
    void *p = d->heap().local() + off;
    g_live[p] = {d, off};
    return p;
  } catch (...) {
    return nullptr; // torch reports the out-of-memory
  }
}

__attribute__((visibility("default"))) void accl_heap_pool_free(void *ptr, ssize_t, int, cudaStream_t) {
  using namespace accl::cuda;
  std::lock_guard<std::mutex> g(g_m);
  auto it = g_live.find(ptr);
  if (it == g_live.end()) return;
  try {
    it->second.first->allocator().free(it->second.second);
  } catch (...) {
  }
  g_live.erase(it);
}

} // extern "C"
