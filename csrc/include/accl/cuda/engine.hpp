// Host side of the persistent engine kernel: launch / park / relaunch, the
// host command ring and stream-ordered doorbells.  See src/cuda/engine.cu.
#pragma once
#include <cuda_runtime.h>

#include "accl/cuda/devtypes.hpp"
#include "accl/cuda/launch.hpp"

namespace accl {
namespace cuda {

class CudaDevice;

class Engine {
public:
  explicit Engine(CudaDevice &dev);
  ~Engine();
  // enqueue one planned work item; completion is published to `hc` and the
  // user stream `s` is made to wait for it (stream-ordered like a kernel)
  void submit(const WorkItem &w, HostCompletion *hc, cudaStream_t s);
  // park the engine kernel now (blocks until it has left the GPU)
  void stop();
  // device-side clients (plugin kernels) are about to issue commands: keep the engine resident
  void pin();
  void unpin();
  struct Impl;

private:
  void launch_locked();
  void ensure_running_locked();
  CudaDevice &dev_;
  Impl *impl_ = nullptr;
};

} // namespace cuda
} // namespace accl
