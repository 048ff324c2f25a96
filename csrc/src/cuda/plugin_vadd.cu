// Vector-add plugins (BASELINE config #4): compute kernels that feed the collective engine THEMSELVES through the
// device API — the host only launches them.  Counterparts of the reference's example user kernel
// kernels/plugins/vadd_put/vadd_put.cpp:25-86 (compute 16 floats, data.push the word, accl.stream_put).
//
//   k_plugin_vadd_allreduce   z = x + y is produced chunk by chunk; the moment the last CTA finishes chunk k the
//                             kernel issues all_reduce_async(chunk k) to the resident engine and goes on computing
//                             chunk k + 1: the engine's workers reduce chunk k over NVLink while the SMs of this
//                             kernel are still adding.  Chunks are issued in index order (every rank must present
//                             the same sequence of collectives); the tickets are finalized at the end.
//   k_plugin_vadd_put         the reference example itself: x + 1 is pushed tile by tile straight into the stream
//                             FIFO of the destination rank (device::Data::push, stream id 9) while computing.
//   k_plugin_loopback         the reference's `loopback` user kernel on the stream ports.
#include "accl/cuda/cudadevice.hpp"
#include "accl/cuda/driver_api.hpp"
#include "accl/cuda/engine.hpp"
#include "accl/cuda/plugins.hpp"
#include "accl/device/api.cuh"

namespace accl {
namespace cuda {

constexpr uint32_t VADD_MAX_CHUNKS = 240;
struct VaddState {                        // device scratch of the plugin, zero between launches
  unsigned int arrived[VADD_MAX_CHUNKS];  // CTAs that finished chunk k
  unsigned long long ticket[VADD_MAX_CHUNKS]; // engine ticket of chunk k's all-reduce (0: not issued yet)
  unsigned int issued;                    // chunks handed to the engine so far (in-order issue)
  unsigned int finished;                  // CTAs that left the compute loop
};

__global__ void __launch_bounds__(512) k_plugin_vadd_allreduce(DevWorld w, uint64_t x_off, uint64_t y_off, uint64_t tmp_off,
                                                               uint64_t out_off, uint32_t count, uint32_t chunk, uint32_t comm_adr,
                                                               uint32_t dpcfg_adr, VaddState *st, uint32_t *status) {
  char *heap = w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes;
  const float *x = reinterpret_cast<const float *>(heap + x_off);
  const float *y = reinterpret_cast<const float *>(heap + y_off);
  float *z = reinterpret_cast<float *>(heap + tmp_off);
  const uint32_t nchunks = (count + chunk - 1) / chunk;
  accl::device::Command accl(heap, comm_adr, dpcfg_adr);
  for (uint32_t k = 0; k < nchunks; ++k) {
    const size_t e0 = static_cast<size_t>(k) * chunk;
    const size_t n = count - e0 < chunk ? count - e0 : chunk;
    const size_t nvec = n / 4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x + e0);
    const float4 *y4 = reinterpret_cast<const float4 *>(y + e0);
    float4 *z4 = reinterpret_cast<float4 *>(z + e0);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const float4 a = x4[i], b = y4[i];
      z4[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    if (blockIdx.x == 0)
      for (size_t i = nvec * 4 + threadIdx.x; i < n; i += blockDim.x) z[e0 + i] = x[e0 + i] + y[e0 + i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&st->arrived[k], 1u) == gridDim.x - 1) {
      // last CTA out of chunk k: the chunk is complete in memory — hand it to the collective engine and keep
      // computing.  In-order issue: every rank must present the same sequence of collectives.
      __threadfence();
      while (*reinterpret_cast<volatile unsigned int *>(&st->issued) != k) dev::nanosleep(50);
      const accl::device::Ticket t = accl.all_reduce_async(static_cast<uint32_t>(n), reduceFunction::SUM, tmp_off + e0 * 4, out_off + e0 * 4);
      st->ticket[k] = t;
      __threadfence();
      atomicExch(&st->issued, k + 1);
    }
  }
  // the last CTA to leave the loop collects the engine's verdicts and re-arms the scratch
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(&st->finished, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    uint32_t rc = 0;
    for (uint32_t k = 0; k < nchunks; ++k) {
      unsigned long long t;
      while ((t = *reinterpret_cast<volatile unsigned long long *>(&st->ticket[k])) == 0) dev::nanosleep(50);
      rc |= accl.finalize_call(t);
      st->ticket[k] = 0;
      st->arrived[k] = 0;
    }
    st->issued = 0;
    st->finished = 0;
    __threadfence();
    *status = rc;
    accl.client_done();
  }
}

// The reference example: out word = in word + 1, pushed into the destination rank's stream (id `stream_id`) while
// computing.  One CTA: a stream is ordered, the tiles of several producers would interleave.
constexpr uint32_t PUT_TILE = 2048; // floats per push (8 KB)
__global__ void __launch_bounds__(256) k_plugin_vadd_put(DevWorld w, uint64_t src_off, uint32_t count, uint32_t dst_rank, uint32_t stream_id,
                                                         uint32_t timeout_us, uint32_t *status) {
  __shared__ __align__(16) float s_tile[PUT_TILE];
  char *heap = w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes;
  const float *src = reinterpret_cast<const float *>(heap + src_off);
  device::Data port(w, stream_id);
  uint32_t e = 0;
  for (uint32_t e0 = 0; e0 < count && !e; e0 += PUT_TILE) {
    const uint32_t n = count - e0 < PUT_TILE ? count - e0 : PUT_TILE;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_tile[i] = src[e0 + i] + 1.0f;
    __syncthreads();
    e = port.push(s_tile, static_cast<uint64_t>(n) * 4, static_cast<int>(dst_rank), static_cast<uint64_t>(timeout_us) * 1000ull);
    __syncthreads();
  }
  if (threadIdx.x == 0 && status) *status = e;
}

// consumer side of vadd_put: drain `count` floats of stream `stream_id` into memory
__global__ void __launch_bounds__(256) k_plugin_stream_pull(DevWorld w, uint64_t dst_off, uint32_t count, uint32_t stream_id,
                                                            uint32_t timeout_us, uint32_t *status) {
  char *heap = w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes;
  device::Data port(w, stream_id);
  uint32_t e = 0;
  for (uint32_t e0 = 0; e0 < count && !e; e0 += PUT_TILE) {
    const uint32_t n = count - e0 < PUT_TILE ? count - e0 : PUT_TILE;
    e = port.pull(heap + dst_off + static_cast<uint64_t>(e0) * 4, static_cast<uint64_t>(n) * 4, static_cast<uint64_t>(timeout_us) * 1000ull);
  }
  if (threadIdx.x == 0 && status) *status = e;
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
  cudaFuncGetAttributes(&a, k_plugin_vadd_put);
  cudaFuncGetAttributes(&a, k_plugin_stream_pull);
  cudaFuncGetAttributes(&a, k_plugin_loopback);
}

cudaError_t launch_loopback(CudaDevice &dev, uint64_t scratch_off, uint32_t count, bool add_one, uint32_t *status_dev,
                            cudaStream_t stream) {
  ACCL_CUDART(cudaSetDevice(dev.device()));
  k_plugin_loopback<<<1, 256, 0, stream>>>(dev.world(), scratch_off, count, add_one ? 1 : 0, status_dev);
  return cudaGetLastError();
}

cudaError_t launch_vadd_put(CudaDevice &dev, uint64_t src_off, uint32_t count, uint32_t dst_rank, uint32_t stream_id,
                            uint32_t *status_dev, cudaStream_t stream) {
  ACCL_CUDART(cudaSetDevice(dev.device()));
  k_plugin_vadd_put<<<1, 256, 0, stream>>>(dev.world(), src_off, count, dst_rank, dev.stream_port_id(stream_id), dev.timeout_us(), status_dev);
  return cudaGetLastError();
}

cudaError_t launch_stream_pull(CudaDevice &dev, uint64_t dst_off, uint32_t count, uint32_t stream_id, uint32_t *status_dev,
                               cudaStream_t stream) {
  ACCL_CUDART(cudaSetDevice(dev.device()));
  k_plugin_stream_pull<<<1, 256, 0, stream>>>(dev.world(), dst_off, count, dev.stream_port_id(stream_id), dev.timeout_us(), status_dev);
  return cudaGetLastError();
}

cudaError_t launch_vadd_allreduce(CudaDevice &dev, uint64_t x_off, uint64_t y_off, uint64_t tmp_off, uint64_t out_off,
                                  uint32_t count, uint32_t chunk_elems, uint32_t comm_adr, uint32_t dpcfg_adr, uint32_t *status_dev,
                                  cudaStream_t stream) {
  Engine *eng = dev.engine();
  if (!eng) throw std::runtime_error("vadd_allreduce plugin needs the persistent engine (engine=True)");
  ACCL_CUDART(cudaSetDevice(dev.device()));
  if (chunk_elems == 0) {
    // every chunk costs one trip through the engine (~10 us more than its transfer time) and only the first chunk's
    // compute is exposed: few, large chunks
    const uint64_t bytes = static_cast<uint64_t>(count) * 4;
    const uint32_t nchunks = bytes <= (8u << 20) ? 1 : bytes <= (64u << 20) ? 2 : 4;
    chunk_elems = (count + nchunks - 1) / nchunks;
    chunk_elems = (chunk_elems + 1023) & ~1023u; // shards of every chunk stay 16-byte aligned for up to 64 ranks
  }
  chunk_elems = (chunk_elems + 3) & ~3u;
  while ((count + chunk_elems - 1) / chunk_elems > VADD_MAX_CHUNKS) chunk_elems *= 2;
  VaddState *st = reinterpret_cast<VaddState *>(dev.plugin_scratch(sizeof(VaddState)));
  eng->client_begin(); // the engine stays resident until the kernel's client_done()
  // the engine's CTAs stay resident next to this kernel: leave them (and a few spare) their SMs
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev.device());
  const int avail = std::max(4, sms - eng->workers() - 4);
  const uint32_t want = static_cast<uint32_t>((std::min<size_t>(count, chunk_elems) / 4 + 511) / 512 + 1);
  const uint32_t grid = std::min<uint32_t>(static_cast<uint32_t>(avail), want);
  k_plugin_vadd_allreduce<<<grid, 512, 0, stream>>>(dev.world(), x_off, y_off, tmp_off, out_off, count, chunk_elems, comm_adr,
                                                    dpcfg_adr, st, status_dev);
  return cudaGetLastError();
}

} // namespace cuda
} // namespace accl
