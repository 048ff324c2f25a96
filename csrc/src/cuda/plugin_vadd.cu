// vector-add plugin (BASELINE config #4): a compute kernel that feeds its
// result into an all-reduce which it issues ITSELF through the device API —
// the host only launches the kernel.  Counterpart of the reference's example
// user kernel kernels/plugins/vadd_put/vadd_put.cpp:25-86 (compute, then
// accl.stream_put from inside the kernel).
#include "accl/cuda/cudadevice.hpp"
#include "accl/cuda/driver_api.hpp"
#include "accl/cuda/engine.hpp"
#include "accl/cuda/plugins.hpp"
#include "accl/device/api.cuh"

namespace accl {
namespace cuda {

__global__ void __launch_bounds__(512) k_plugin_vadd_allreduce(DevWorld w, uint64_t x_off, uint64_t y_off, uint64_t tmp_off,
                                                               uint64_t out_off, uint32_t count, uint32_t comm_adr,
                                                               uint32_t dpcfg_adr, unsigned int *counter, uint32_t *status) {
  char *heap = w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes;
  const float4 *x = reinterpret_cast<const float4 *>(heap + x_off);
  const float4 *y = reinterpret_cast<const float4 *>(heap + y_off);
  float4 *z = reinterpret_cast<float4 *>(heap + tmp_off);
  const size_t nvec = count / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = x[i], b = y[i];
    z[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
  if (blockIdx.x == 0)
    for (size_t i = nvec * 4 + threadIdx.x; i < count; i += blockDim.x)
      reinterpret_cast<float *>(heap + tmp_off)[i] = reinterpret_cast<const float *>(heap + x_off)[i] + reinterpret_cast<const float *>(heap + y_off)[i];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(counter, 1u) == gridDim.x - 1) {
    // last CTA out: the sum is complete in memory, hand it to the collective engine
    __threadfence();
    accl::device::Command accl(heap, comm_adr, dpcfg_adr);
    *status = accl.all_reduce(count, reduceFunction::SUM, tmp_off, out_off);
    *counter = 0;
  }
}

// The reference's `loopback` user kernel (kernels/plugins/loopback/loopback.cpp): consume words from the
// engine->kernel stream and hand them back on the kernel->engine stream, here with an optional +1 so a
// test can tell the data really went through the kernel.  One CTA; `Data` is the device API's stream port.
__global__ void __launch_bounds__(256) k_plugin_loopback(DevWorld w, uint64_t scratch_off, uint32_t count, int add_one,
                                                         uint32_t *status) {
  char *heap = w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes;
  float *tmp = reinterpret_cast<float *>(heap + scratch_off);
  device::Data port(w);
  uint32_t e = port.pull(tmp, static_cast<uint64_t>(count) * 4);
  if (!e) {
    if (add_one)
      for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) tmp[i] += 1.0f;
    __syncthreads();
    e = port.push(tmp, static_cast<uint64_t>(count) * 4);
  }
  if (threadIdx.x == 0 && status) *status = e;
}

void preload_vadd_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_plugin_vadd_allreduce);
  cudaFuncGetAttributes(&a, k_plugin_loopback);
}

cudaError_t launch_loopback(CudaDevice &dev, uint64_t scratch_off, uint32_t count, bool add_one, uint32_t *status_dev,
                            cudaStream_t stream) {
  ACCL_CUDART(cudaSetDevice(dev.device()));
  k_plugin_loopback<<<1, 256, 0, stream>>>(dev.world(), scratch_off, count, add_one ? 1 : 0, status_dev);
  return cudaGetLastError();
}

static void CUDART_CB unpin_cb(void *user) { static_cast<Engine *>(user)->unpin(); }

cudaError_t launch_vadd_allreduce(CudaDevice &dev, uint64_t x_off, uint64_t y_off, uint64_t tmp_off, uint64_t out_off,
                                  uint32_t count, uint32_t comm_adr, uint32_t dpcfg_adr, uint32_t *status_dev,
                                  cudaStream_t stream) {
  Engine *eng = dev.engine();
  if (!eng) throw std::runtime_error("vadd_allreduce plugin needs the persistent engine (engine=True)");
  ACCL_CUDART(cudaSetDevice(dev.device()));
  unsigned int *ctr = dev.plugin_counter(); // per device: plugin launches are stream ordered
  eng->pin(); // keep the engine resident while a device-side client may issue commands
  const uint32_t grid = static_cast<uint32_t>(std::min<size_t>(148, (count / 4 + 511) / 512 + 1));
  k_plugin_vadd_allreduce<<<grid, 512, 0, stream>>>(dev.world(), x_off, y_off, tmp_off, out_off, count, comm_adr, dpcfg_adr, ctr,
                                                    status_dev);
  cudaError_t e = cudaGetLastError();
  cudaLaunchHostFunc(stream, unpin_cb, eng);
  return e;
}

} // namespace cuda
} // namespace accl
