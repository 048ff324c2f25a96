// The persistent collective engine: one long-running sm_100a kernel per GPU
// that owns command queues in device memory and executes calls itself.
//
//   control CTA (last block)   the "firmware": arbitrates the host ring and
//                              the device ring (client arbiter), decodes and
//                              plans device-issued descriptors from the
//                              device-resident exchange memory, publishes work
//                              items, retires them in order
//   worker CTAs                run the same `run_work` bodies as the direct
//                              launch path; the last one out writes the
//                              completion (retcode + %globaltimer duration)
//
// Reference counterpart: the continuously running CCLO — MicroBlaze main loop
// (ccl_offload_control.c:2264-2483: wait_for_call / dispatch / finalize_call),
// hostctrl (kernels/plugins/hostctrl/hostctrl.cpp:22-63) and client_arbiter
// (kernels/plugins/client_arbiter/client_arbiter.cpp:21-51).
//
// Host commands arrive through a pinned host ring; their doorbell is either a
// stream-ordered cuStreamWriteValue64 into device memory (call is ordered
// after prior work on the user's stream, and the stream then waits on the
// engine's done counter with cuStreamWaitValue64) or, for stream-less use, a
// plain store into pinned memory that the control CTA polls.  Device commands
// (plugins using accl/device/api.cuh) take a ticket in the device ring.
//
// The kernel parks itself after `engine_idle_us` without work so that
// device-wide synchronisation (cudaDeviceSynchronize, cudaFree…) cannot hang;
// the host relaunches it on the next submit (Dekker-style handshake on
// `submitted` / `state` in pinned memory).
#include "accl/cuda/engine.hpp"

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "accl/common.hpp"
#include "accl/cuda/cudadevice.hpp"
#include "accl/cuda/driver_api.hpp"
#include "accl/cuda/plan.hpp"
#include "accl/device/api.cuh"
#include "run_work.cuh"

namespace accl {
namespace cuda {

__device__ __forceinline__ void publish_completion(HostCompletion *hc, uint32_t seq, uint32_t rc, unsigned long long dur) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(hc), "r"(seq), "r"(rc), "r"(static_cast<uint32_t>(dur)),
               "r"(static_cast<uint32_t>(dur >> 32))
               : "memory");
}

enum EngineState : uint32_t { ENG_STOPPED = 0, ENG_RUNNING = 1, ENG_EXITING = 2 };

struct HostRingSlot {
  WorkItem item;
};

// lives in pinned, device-mapped host memory
struct HostRing {
  volatile unsigned long long tail_direct; // stream-less doorbell
  volatile unsigned long long submitted;   // entries the host has placed in the ring (engine must not park before fetching them)
  volatile uint32_t state;                 // EngineState, written by the control CTA
  volatile uint32_t pins;                  // device-side clients active: do not park
  volatile uint32_t stop;                  // host asks the engine to leave now
  volatile uint32_t idle_us;
  HostRingSlot slots[RING_SLOTS];
};

struct EngineArea { // at heap offset CTRL_BYTES / 2
  WorkItem issue_ring[ISSUE_SLOTS];
};

__device__ __forceinline__ EngineArea *engine_area(const DevWorld &w) {
  return reinterpret_cast<EngineArea *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes + CTRL_BYTES / 2);
}

// ------------------------------------------------------------------ workers
__device__ void engine_worker(const DevWorld &w, int nworkers) {
  __shared__ uint32_t s_err;
  __shared__ WorkItem s_item;
  __shared__ unsigned long long s_seq;
  __shared__ uint32_t s_exit;
  Ctrl *me = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
  EngineArea *ea = engine_area(w);
  unsigned long long next = 0;
  if (threadIdx.x == 0) s_seq = dev::ld_acquire_gpu(&me->done_count); // resume where the previous instance stopped
  __syncthreads();
  next = s_seq;
  for (;;) {
    if (threadIdx.x == 0) {
      uint32_t spins = 0, ex = 0;
      while (dev::ld_acquire_gpu(&me->issue_tail) <= next) {
        if (dev::ld_relaxed_sys(&me->engine_exit)) {
          ex = 1;
          break;
        }
        if (++spins > 16) dev::nanosleep(spins > 2048 ? 400 : 40);
      }
      s_exit = ex;
      s_err = 0;
    }
    __syncthreads();
    if (s_exit) return;
    // stage the item in shared memory (one 8-byte word per lane)
    {
      const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&ea->issue_ring[next % ISSUE_SLOTS]);
      unsigned long long *dst = reinterpret_cast<unsigned long long *>(&s_item);
      for (uint32_t i = threadIdx.x; i < sizeof(WorkItem) / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int nctas = static_cast<int>(s_item.n_ctas);
    if (static_cast<int>(blockIdx.x) < nctas) {
      const unsigned long long t0 = dev::globaltimer_ns();
      k::run_work(w, s_item, blockIdx.x, nctas, &s_err);
      __syncthreads();
      if (threadIdx.x == 0) {
        Completion *cp = &me->comp[s_item.req_slot];
        if (s_err) atomicOr(&cp->retcode, s_err);
        atomicMin(&cp->t_start, t0);
        atomicMax(&cp->t_end, dev::globaltimer_ns());
        __threadfence();
        if (atomicAdd(&cp->done_ctas, 1u) == static_cast<uint32_t>(nctas) - 1) {
          __threadfence();
          const uint32_t rc = cp->retcode;
          const unsigned long long dur = cp->t_end - cp->t_start;
          me->exch[exchmem::RETCODE / 4] = rc;
          me->exch[exchmem::PERFCNT_LO / 4] = static_cast<uint32_t>(dur);
          cp->retcode = 0;
          cp->done_ctas = 0;
          cp->t_start = ~0ull;
          cp->t_end = 0;
          __threadfence();
          if (s_item.dev_ticket)
            dev::st_release_sys(&me->dev_status[(s_item.dev_ticket - 1) % RING_SLOTS],
                                s_item.dev_ticket | (static_cast<unsigned long long>(rc) << 32));
          if (s_item.host_seq) dev::st_release_sys(&me->host_done, s_item.host_seq);
          if (s_item.hc_ptr) publish_completion(reinterpret_cast<HostCompletion *>(s_item.hc_ptr), s_item.req_seq, rc, dur);
          // in-order retirement: the control CTA and stream waits key off this counter
          dev::st_release_sys(&me->done_count, next + 1);
        }
      }
    }
    ++next;
    __syncthreads();
  }
  (void)nworkers;
}

// ------------------------------------------------------------------ control
__device__ void engine_control(const DevWorld &w, HostRing *hr) {
  if (threadIdx.x >= 32) return;
  const unsigned lane = threadIdx.x;
  Ctrl *me = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
  EngineArea *ea = engine_area(w);
  PlanCfg pcfg;
  memcpy(&pcfg, me->plan_cfg_words, sizeof(PlanCfg));
  unsigned long long host_fetched = me->host_fetched, dev_fetched = me->dev_fetched;
  unsigned long long issued = dev::ld_acquire_gpu(&me->done_count);
  unsigned long long last_work_ns = dev::globaltimer_ns();
  uint32_t poll = 0;
  unsigned long long direct_tail = 0;
  if (lane == 0) hr->state = ENG_RUNNING;
  for (;;) {
    // ---- what is pending?
    unsigned long long htail = 0, ready = 0;
    uint32_t stop = 0;
    if (lane == 0) {
      htail = dev::ld_acquire_sys(&me->host_tail);
      if ((poll++ & 7) == 0 || htail <= host_fetched) direct_tail = hr->tail_direct; // PCIe read: not every turn
      if (direct_tail > htail) htail = direct_tail;
      ready = dev::ld_acquire_sys(&me->dev_ready[dev_fetched % RING_SLOTS]);
      stop = hr->stop;
    }
    htail = __shfl_sync(0xffffffffu, htail, 0);
    ready = __shfl_sync(0xffffffffu, ready, 0);
    stop = __shfl_sync(0xffffffffu, stop, 0);
    WorkItem *slot = &ea->issue_ring[issued % ISSUE_SLOTS];
    bool have = false;
    if (host_fetched < htail) {
      // host command: already planned, copy it over PCIe with the whole warp
      const unsigned long long *src =
          reinterpret_cast<const unsigned long long *>(const_cast<const WorkItem *>(&hr->slots[host_fetched % RING_SLOTS].item));
      unsigned long long *dst = reinterpret_cast<unsigned long long *>(slot);
      for (uint32_t i = lane; i < sizeof(WorkItem) / 8; i += 32) dst[i] = src[i];
      ++host_fetched;
      have = true;
    } else if (ready == dev_fetched + 1) {
      // device command: decode + plan here, from the device copy of exchange memory
      if (lane == 0) {
        WorkItem wi;
        const CallDesc d = me->dev_ring[dev_fetched % RING_SLOTS];
        const uint32_t e = build_work_item_hd(me->exch, pcfg, w.world, d, me->engine_timeout_us, wi);
        wi.req_slot = N_REQ_SLOTS - 1; // device calls share the last completion record (they retire in order)
        wi.req_seq = 0;
        wi.hc_ptr = 0;
        wi.dev_ticket = dev_fetched + 1;
        wi.host_seq = 0;
        if (e) { // undecodable: retire immediately with the error
          dev::st_release_sys(&me->dev_status[dev_fetched % RING_SLOTS], (dev_fetched + 1) | (static_cast<unsigned long long>(e) << 32));
          wi.desc.scenario = static_cast<uint32_t>(operation::nop);
          wi.algo = ALGO_LOCAL;
          wi.n_ctas = 1;
          wi.dev_ticket = 0;
        }
        *slot = wi;
      }
      ++dev_fetched;
      have = true;
    }
    __syncwarp();
    if (have) {
      if (lane == 0) {
        me->host_fetched = host_fetched;
        me->dev_fetched = dev_fetched;
        __threadfence();
        dev::st_release_gpu(&me->issue_tail, issued + 1);
        // one call at a time (like the reference engine): wait for retirement
        uint32_t spins = 0;
        while (dev::ld_acquire_gpu(&me->done_count) <= issued)
          if (++spins > 8) dev::nanosleep(40);
      }
      __syncwarp();
      ++issued;
      last_work_ns = dev::globaltimer_ns();
      continue;
    }
    // ---- idle: park after idle_us unless pinned; leave at once when told to stop
    bool leave = false;
    if (lane == 0) {
      const uint32_t idle_us = hr->idle_us;
      const bool idle_long = idle_us != 0 && dev::globaltimer_ns() - last_work_ns > static_cast<unsigned long long>(idle_us) * 1000ull;
      if (stop || (idle_long && hr->pins == 0)) {
        hr->state = ENG_EXITING;
        dev::fence_sc_sys();
        const bool pending = hr->submitted != host_fetched || dev::ld_acquire_sys(&me->dev_ready[dev_fetched % RING_SLOTS]) == dev_fetched + 1;
        if (pending && !stop) {
          hr->state = ENG_RUNNING; // a submit raced with parking: keep going
        } else {
          leave = true;
        }
      }
      if (!leave) dev::nanosleep(100);
    }
    leave = __shfl_sync(0xffffffffu, leave ? 1 : 0, 0) != 0;
    if (leave) {
      if (lane == 0) {
        dev::st_release_sys(&me->engine_exit, 1u);
        __threadfence_system();
        hr->state = ENG_STOPPED;
      }
      return;
    }
  }
}

__global__ void __launch_bounds__(k::BLOCK) k_engine(DevWorld w, HostRing *hr, int nworkers) {
  if (static_cast<int>(blockIdx.x) == nworkers) engine_control(w, hr);
  else engine_worker(w, nworkers);
}

__global__ void k_engine_prepare(DevWorld w, PlanCfg cfg, uint32_t timeout_us) {
  Ctrl *me = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
  memcpy(me->plan_cfg_words, &cfg, sizeof(PlanCfg));
  me->engine_timeout_us = timeout_us;
  me->engine_exit = 0;
}

// ------------------------------------------------------ direct-launch path
// One kernel per call, stream-ordered like any other CUDA work; shares every
// device function with the engine above (one translation unit, one copy).
__global__ void __launch_bounds__(k::BLOCK) k_call(DevWorld w, WorkItem it, HostCompletion *hc) {
  __shared__ uint32_t s_err;
  const unsigned long long t0 = dev::globaltimer_ns();
  if (threadIdx.x == 0) s_err = 0;
  __syncthreads();
  k::run_work(w, it, blockIdx.x, gridDim.x, &s_err);
  __syncthreads();
#ifdef ACCL_PHASE_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    Ctrl *dc = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
    dc->dbg_calls += 1;
    dc->dbg_kernel_ns += dev::globaltimer_ns() - t0;
  }
#endif
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) { // stream helper kernels that ran before this call report through it
      Ctrl *mc = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
      const uint32_t se = mc->strm_err;
      if (se) {
        s_err |= se;
        mc->strm_err = 0;
      }
    }
    const unsigned long long t1 = dev::globaltimer_ns();
    if (it.flags & WF_CHAIN) {
      // a later kernel of the same lowered call reports for all of them
      if (s_err) atomicOr(&reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes)->strm_err, s_err);
    } else if (gridDim.x == 1) {
      publish_completion(hc, it.req_seq, s_err, t1 - t0);
    } else {
      Ctrl *me = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
      Completion *cp = &me->comp[it.req_slot];
      if (s_err) atomicOr(&cp->retcode, s_err);
      atomicMin(&cp->t_start, t0);
      atomicMax(&cp->t_end, t1);
      __threadfence();
      if (atomicAdd(&cp->done_ctas, 1u) == gridDim.x - 1) {
        // last CTA out: publish to the host and recycle the record
        __threadfence();
        const uint32_t rc = cp->retcode;
        const unsigned long long dur = cp->t_end - cp->t_start;
        cp->retcode = 0;
        cp->done_ctas = 0;
        cp->t_start = ~0ull;
        cp->t_end = 0;
        publish_completion(hc, it.req_seq, rc, dur);
      }
    }
  }
}

cudaError_t launch_call(const DevWorld &w, const WorkItem &item, HostCompletion *hc_dev, cudaStream_t stream) {
  k_call<<<item.n_ctas, k::BLOCK, 0, stream>>>(w, item, hc_dev);
  return cudaGetLastError();
}

__global__ void k_reset_ctrl(DevWorld w) {
  Ctrl *me = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
  // everything after the exchange memory is protocol state
  uint32_t *p = reinterpret_cast<uint32_t *>(&me->sig[0][0]);
  const size_t n = (sizeof(Ctrl) - offsetof(Ctrl, sig)) / 4;
  for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) p[i] = 0;
  if (blockIdx.x == 0 && threadIdx.x < N_REQ_SLOTS) {
    __syncthreads();
  }
}
__global__ void k_init_comp(DevWorld w) {
  Ctrl *me = reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes);
  for (int i = threadIdx.x; i < N_REQ_SLOTS; i += blockDim.x) me->comp[i].t_start = ~0ull;
}

cudaError_t launch_reset_ctrl(const DevWorld &w, cudaStream_t stream) {
  k_reset_ctrl<<<8, 256, 0, stream>>>(w);
  k_init_comp<<<1, 256, 0, stream>>>(w);
  return cudaGetLastError();
}


// ---------------------------------------------------------------- stream port
// FIFO -> local buffer / local buffer -> FIFO of rank dst_rank (my own: RES_STREAM results; a peer's:
// stream_put).  One CTA each; the bodies are the device API's Data port (accl/device/api.cuh).
__global__ void __launch_bounds__(512) k_stream_pop(DevWorld w, uint64_t dst_off, uint64_t bytes, uint32_t timeout_us) {
  char *heap = w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes;
  device::Data port(w);
  const uint32_t e = port.pull(heap + dst_off, bytes, static_cast<uint64_t>(timeout_us) * 1000ull);
  if (e && threadIdx.x == 0) atomicOr(&reinterpret_cast<Ctrl *>(heap)->strm_err, e);
}

__global__ void __launch_bounds__(512) k_stream_push(DevWorld w, uint32_t dst_rank, uint64_t src_off, uint64_t bytes,
                                                     uint32_t timeout_us) {
  char *heap = w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes;
  device::Data port(w);
  const uint32_t e = port.push(heap + src_off, bytes, static_cast<int>(dst_rank), static_cast<uint64_t>(timeout_us) * 1000ull);
  if (e && threadIdx.x == 0) atomicOr(&reinterpret_cast<Ctrl *>(heap)->strm_err, e);
}

cudaError_t launch_stream_pop(const DevWorld &w, uint64_t dst_off, uint64_t bytes, uint32_t timeout_us, cudaStream_t stream) {
  k_stream_pop<<<1, 512, 0, stream>>>(w, dst_off, bytes, timeout_us);
  return cudaGetLastError();
}
cudaError_t launch_stream_push(const DevWorld &w, uint32_t dst_rank, uint64_t src_off, uint64_t bytes, uint32_t timeout_us,
                               cudaStream_t stream) {
  k_stream_push<<<1, 512, 0, stream>>>(w, dst_rank, src_off, bytes, timeout_us);
  return cudaGetLastError();
}

void preload_engine_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_stream_pop);
  cudaFuncGetAttributes(&a, k_stream_push);
  cudaFuncGetAttributes(&a, k_engine);
  cudaFuncGetAttributes(&a, k_engine_prepare);
  cudaFuncGetAttributes(&a, k_call);
  cudaFuncGetAttributes(&a, k_reset_ctrl);
  cudaFuncGetAttributes(&a, k_init_comp);
}

// --------------------------------------------------------------------- host
struct Engine::Impl {
  HostRing *ring = nullptr, *ring_dev = nullptr;
  cudaStream_t stream = nullptr;
  unsigned long long submitted = 0;
  int nworkers = 0;
  std::mutex m;
};

static_assert(sizeof(PlanCfg) <= 32, "PlanCfg must fit plan_cfg_words");

Engine::Engine(CudaDevice &dev) : dev_(dev), impl_(new Impl()) {
  ACCL_CUDART(cudaSetDevice(dev_.device()));
  ACCL_CUDART(cudaHostAlloc(reinterpret_cast<void **>(&impl_->ring), sizeof(HostRing), cudaHostAllocMapped | cudaHostAllocPortable));
  std::memset(impl_->ring, 0, sizeof(HostRing));
  ACCL_CUDART(cudaHostGetDevicePointer(reinterpret_cast<void **>(&impl_->ring_dev), impl_->ring, 0));
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  ACCL_CUDART(cudaStreamCreateWithPriority(&impl_->stream, cudaStreamNonBlocking, hi));
  impl_->nworkers = dev_.config().max_ctas;
  impl_->ring->idle_us = static_cast<uint32_t>(dev_.config().engine_idle_us);
  impl_->ring->state = ENG_STOPPED;
}

Engine::~Engine() {
  try {
    stop();
  } catch (...) {
  }
  if (impl_) {
    cudaSetDevice(dev_.device());
    if (impl_->stream) cudaStreamDestroy(impl_->stream);
    // the ring is pinned host memory: cudaFreeHost synchronises the device, which is fine here
    if (impl_->ring) cudaFreeHost(impl_->ring);
  }
  delete impl_;
}

void Engine::launch_locked() {
  HostRing *r = impl_->ring;
  ACCL_CUDART(cudaSetDevice(dev_.device()));
  r->stop = 0;
  r->state = ENG_RUNNING; // optimistic: the kernel re-asserts it
  std::atomic_thread_fence(std::memory_order_seq_cst);
  k_engine_prepare<<<1, 1, 0, impl_->stream>>>(dev_.world(), dev_.plan_cfg(), dev_.timeout_us());
  k_engine<<<impl_->nworkers + 1, k::BLOCK, 0, impl_->stream>>>(dev_.world(), impl_->ring_dev, impl_->nworkers);
  ACCL_CUDART(cudaGetLastError());
}

// make sure an engine instance is (or will be) alive for everything submitted so far
void Engine::ensure_running_locked() {
  HostRing *r = impl_->ring;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  for (;;) {
    const uint32_t st = r->state;
    if (st == ENG_RUNNING) return;
    if (st == ENG_STOPPED) {
      launch_locked();
      return;
    }
    std::this_thread::yield(); // EXITING: it either resumes or stops within microseconds
  }
}

void Engine::stop() {
  if (!impl_ || !impl_->ring) return;
  std::lock_guard<std::mutex> g(impl_->m);
  HostRing *r = impl_->ring;
  cudaSetDevice(dev_.device());
  if (r->state != ENG_STOPPED) {
    r->stop = 1;
    std::atomic_thread_fence(std::memory_order_seq_cst);
  }
  cudaStreamSynchronize(impl_->stream);
  r->state = ENG_STOPPED;
  r->stop = 0;
}

void Engine::pin() {
  std::lock_guard<std::mutex> g(impl_->m);
  impl_->ring->pins = impl_->ring->pins + 1;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  ensure_running_locked();
}

void Engine::unpin() {
  std::lock_guard<std::mutex> g(impl_->m);
  if (impl_->ring->pins) impl_->ring->pins = impl_->ring->pins - 1;
}

void Engine::submit(const WorkItem &w, HostCompletion *hc, cudaStream_t s) {
  std::lock_guard<std::mutex> g(impl_->m);
  HostRing *r = impl_->ring;
  auto &drv = DriverApi::get();
  Ctrl *ctrl = reinterpret_cast<Ctrl *>(dev_.heap().local());
  // ring full?  the engine retires in order: wait until the oldest entry has been fetched
  // (host_fetched is mirrored in done_count for host-issued calls only approximately, so
  // bound the number of outstanding submits by the ring size using the completion records)
  const unsigned long long seq = impl_->submitted;
  WorkItem item = w;
  item.hc_ptr = reinterpret_cast<uint64_t>(hc);
  item.dev_ticket = 0;
  item.host_seq = seq + 1;
  item.flags |= WF_ENGINE;
  std::memcpy(const_cast<WorkItem *>(&r->slots[seq % RING_SLOTS].item), &item, sizeof(WorkItem));
  std::atomic_thread_fence(std::memory_order_seq_cst);
  impl_->submitted = seq + 1;
  r->submitted = seq + 1;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  ensure_running_locked();
  if (s) {
    // stream-ordered: the doorbell fires when the user's stream gets here; the stream then
    // waits until the engine has retired this call (host_done counts retired host entries)
    ACCL_CU(drv.cuStreamWriteValue64(s, reinterpret_cast<CUdeviceptr>(&ctrl->host_tail), seq + 1, 0));
    ACCL_CU(drv.cuStreamWaitValue64(s, reinterpret_cast<CUdeviceptr>(&ctrl->host_done), seq + 1, CU_STREAM_WAIT_VALUE_GEQ));
  } else {
    r->tail_direct = seq + 1;
  }
}

} // namespace cuda
} // namespace accl
