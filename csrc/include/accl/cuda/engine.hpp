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
  // enqueue one planned work item through a proxy kernel on stream `s` (ordered after the work already
  // queued there); completion is published to `hc`; with `stream_waits` the proxy also holds the stream until
  // the engine has retired the call (stream-ordered like a direct launch)
  void submit(const WorkItem &w, HostCompletion *hc, cudaStream_t s, bool stream_waits);
  int workers() const;
  // park the engine kernel now (blocks until it has left the GPU)
  void stop();
  // device-side clients (plugin kernels) are about to issue commands: keep the engine resident
  void pin();
  // a kernel that issues commands itself is about to be launched: the engine stays resident until that kernel calls
  // device::Command::client_done() (no host callback on the stream)
  void client_begin();
  void clients_reset(); // the control block was zeroed (soft reset)
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
