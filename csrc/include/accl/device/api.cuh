// Device-side API: lets a compute kernel ("plugin") issue collective calls to
// the persistent engine itself — no host on the path — and exchange data with
// it through the device-side stream port.
//
// Same surface as the reference's HLS bindings (driver/hls/accl_hls.h:82-541):
// `Command` binds (communicator address, arithmetic-config address, default
// compression / stream flags); `start_call` emits the same 15-word descriptor
// (scenario, len, comm, root_src_dst, function, tag, datapath_cfg,
// compression_flags, stream_flags, addr a/b/c), `finalize_call` returns the
// engine's status word; convenience wrappers copy / combine / send /
// stream_put / recv / bcast / scatter / gather / all_gather / reduce /
// reduce_scatter / all_reduce.  Instead of pushing words into an AXI stream,
// one thread takes a ticket in the engine's device ring (MPMC: any number of
// kernels / CTAs may issue concurrently — the client arbiter) and publishes the
// descriptor with release semantics; completion is an acquire spin on the
// slot's status word.
//
// Addresses are byte offsets into the caller's symmetric heap
// (`Buffer.address` on the host side).
#pragma once
#include "accl/cuda/devtypes.hpp"
#include "accl/device/primitives.cuh"

namespace accl {
namespace device {

using Ticket = unsigned long long;

class Command {
public:
  // heap_base: this rank's heap (DevWorld.window + rank * heap_bytes);
  // comm_adr / dpcfg_adr: ACCL::get_communicator_addr / get_arithmetic_config_addr
  __device__ Command(void *heap_base, uint32_t comm_adr, uint32_t dpcfg_adr, uint32_t cflags = 0, uint32_t sflags = 0)
      : ctrl_(static_cast<cuda::Ctrl *>(heap_base)),
        ring_(reinterpret_cast<cuda::EngineArea *>(static_cast<char *>(heap_base) + cuda::CTRL_BYTES / 2)),
        comm_((comm_adr - exchmem::COMM_BASE) / exchmem::COMM_STRIDE), dpcfg_(dpcfg_adr), cflags_(cflags), sflags_(sflags) {}

  // Call from ONE thread.  Returns the ticket to pass to finalize_call.
  __device__ Ticket start_call(uint32_t scenario, uint32_t len, uint32_t comm, uint32_t root_src_dst, uint32_t function,
                               uint32_t tag, uint32_t datapath_cfg, uint32_t compression_flags, uint32_t stream_flags,
                               uint64_t addra, uint64_t addrb, uint64_t addrc) {
    const Ticket t = atomicAdd(&ctrl_->cmd_tail, 1ull);
    // ring full: wait until the engine has consumed the entry that used this slot
    while (t >= dev::ld_acquire_sys(&ctrl_->cmd_fetched) + cuda::RING_SLOTS) dev::nanosleep(200);
    CallDesc d;
    d.scenario = scenario;
    d.count = len;
    d.comm = comm;
    d.root_src_dst = root_src_dst;
    d.function = function;
    d.tag = tag;
    d.arithcfg = datapath_cfg;
    d.compression_flags = compression_flags;
    d.stream_host_flags = stream_flags;
    d.set_addr(0, addra);
    d.set_addr(1, addrb);
    d.set_addr(2, addrc);
    d.aux = 0;
    cuda::WorkItem *slot = &ring_->cmd_ring[t % cuda::RING_SLOTS].item;
    slot->desc = d;
    slot->flags = 0; // not planned: the engine's control CTA decodes and plans the descriptor
    slot->hc_ptr = 0;
    __threadfence();
    dev::st_release_sys(&ctrl_->cmd_ready[t % cuda::RING_SLOTS], t + 1);
    return t + 1;
  }

  // Blocks until the engine has retired the call; returns its error word.
  __device__ uint32_t finalize_call(Ticket ticket) {
    const unsigned long long *st = &ctrl_->cmd_status[(ticket - 1) % cuda::RING_SLOTS];
    unsigned long long v;
    uint32_t spins = 0;
    while (((v = dev::ld_acquire_sys(st)) & 0xFFFFFFFFull) != (ticket & 0xFFFFFFFFull))
      if (++spins > 8) dev::nanosleep(100);
    return static_cast<uint32_t>(v >> 32);
  }

  // ---- convenience wrappers (start + finalize), argument order as in accl_hls.h
  __device__ uint32_t copy(uint32_t len, uint64_t src, uint64_t dst) {
    return run(operation::copy, len, 0, 0, TAG_ANY, src, 0, dst);
  }
  __device__ uint32_t combine(uint32_t len, reduceFunction fn, uint64_t op0, uint64_t op1, uint64_t res) {
    return run(operation::combine, len, 0, static_cast<uint32_t>(fn), TAG_ANY, op0, op1, res);
  }
  __device__ uint32_t send(uint32_t len, uint32_t tag, uint32_t dst_rank, uint64_t src) {
    return run(operation::send, len, dst_rank, 0, tag, src, 0, 0);
  }
  __device__ uint32_t stream_put(uint32_t len, uint32_t stream_id, uint32_t dst_rank, uint64_t src) {
    if (stream_id < STREAM_ID_MIN || stream_id > STREAM_ID_MAX) return CONFIG_SWITCH_ERROR;
    return finalize_call(start_call(static_cast<uint32_t>(operation::send), len, comm_, dst_rank, 0, stream_id, dpcfg_, cflags_,
                                    sflags_ | static_cast<uint32_t>(streamFlags::RES_STREAM), src, 0, 0));
  }
  __device__ uint32_t recv(uint32_t len, uint32_t tag, uint32_t src_rank, uint64_t dst) {
    return run(operation::recv, len, src_rank, 0, tag, 0, 0, dst);
  }
  __device__ uint32_t bcast(uint32_t len, uint32_t root, uint64_t buf) {
    return run(operation::bcast, len, root, 0, TAG_ANY, buf, 0, 0);
  }
  __device__ uint32_t scatter(uint32_t len, uint32_t root, uint64_t src, uint64_t dst) {
    return run(operation::scatter, len, root, 0, TAG_ANY, src, 0, dst);
  }
  __device__ uint32_t gather(uint32_t len, uint32_t root, uint64_t src, uint64_t dst) {
    return run(operation::gather, len, root, 0, TAG_ANY, src, 0, dst);
  }
  __device__ uint32_t all_gather(uint32_t len, uint64_t src, uint64_t dst) {
    return run(operation::allgather, len, 0, 0, TAG_ANY, src, 0, dst);
  }
  __device__ uint32_t reduce(uint32_t len, uint32_t root, reduceFunction fn, uint64_t src, uint64_t dst) {
    return run(operation::reduce, len, root, static_cast<uint32_t>(fn), TAG_ANY, src, 0, dst);
  }
  __device__ uint32_t reduce_scatter(uint32_t len, reduceFunction fn, uint64_t src, uint64_t dst) {
    return run(operation::reduce_scatter, len, 0, static_cast<uint32_t>(fn), TAG_ANY, src, 0, dst);
  }
  __device__ uint32_t all_reduce(uint32_t len, reduceFunction fn, uint64_t src, uint64_t dst) {
    return run(operation::allreduce, len, 0, static_cast<uint32_t>(fn), TAG_ANY, src, 0, dst);
  }
  __device__ uint32_t barrier() { return run(operation::barrier, 0, 0, 0, TAG_ANY, 0, 0, 0); }
  __device__ uint32_t nop() { return run(operation::nop, 0, 0, 0, TAG_ANY, 0, 0, 0); }

  // asynchronous flavours: issue now, finalize later (lets a kernel overlap compute with the collective)
  __device__ Ticket all_reduce_async(uint32_t len, reduceFunction fn, uint64_t src, uint64_t dst) {
    return start_call(static_cast<uint32_t>(operation::allreduce), len, comm_, 0, static_cast<uint32_t>(fn), TAG_ANY, dpcfg_,
                      cflags_, sflags_, src, 0, dst);
  }
  __device__ Ticket reduce_scatter_async(uint32_t len, reduceFunction fn, uint64_t src, uint64_t dst) {
    return start_call(static_cast<uint32_t>(operation::reduce_scatter), len, comm_, 0, static_cast<uint32_t>(fn), TAG_ANY, dpcfg_,
                      cflags_, sflags_, src, 0, dst);
  }
  // non-blocking probe of an asynchronous call: true once retired (then *retcode holds the error word)
  // Last thing a client kernel does (ONE thread, after its last finalize_call): tells the engine that this device-side
  // client has finished, so the resident kernel may park again when idle.  The host side of the pair is
  // Engine::client_begin(), called right before the client kernel is launched.
  __device__ void client_done() {
    __threadfence();
    atomicAdd(&ctrl_->clients_done, 1u);
  }

  __device__ bool test_call(Ticket ticket, uint32_t *retcode) {
    const unsigned long long v = dev::ld_acquire_sys(&ctrl_->cmd_status[(ticket - 1) % cuda::RING_SLOTS]);
    if ((v & 0xFFFFFFFFull) != (ticket & 0xFFFFFFFFull)) return false;
    *retcode = static_cast<uint32_t>(v >> 32);
    return true;
  }

private:
  __device__ uint32_t run(operation op, uint32_t len, uint32_t root, uint32_t fn, uint32_t tag, uint64_t a, uint64_t b,
                          uint64_t c) {
    return finalize_call(start_call(static_cast<uint32_t>(op), len, comm_, root, fn, tag, dpcfg_, cflags_, sflags_, a, b, c));
  }
  cuda::Ctrl *ctrl_;
  cuda::EngineArea *ring_;
  uint32_t comm_, dpcfg_, cflags_, sflags_;
};

// Data port of a user kernel: one of the rank's stream FIFOs (reference ACCLData, accl_hls.h:455-541: push / pull
// of 64-byte words on the streams between the user kernel and the CCLO, routed by TDEST = stream id,
// dma_mover.cpp:312,644).  Stream id s is served by FIFO s % N_STRM_PORTS, so consumers of different ids do
// not interleave; id 0 = the default port of stream operands without an id.  A FIFO is a byte ring in the
// symmetric heap; `push` may target a peer's ring (what `stream_put` lowers to).  All methods are
// cooperative over ONE CTA (every thread of the block must call them with the same arguments).
class Data {
public:
  __device__ explicit Data(const cuda::DevWorld &w, uint32_t stream_id = 0) : w_(w), port_(stream_id % cuda::N_STRM_PORTS) {}

  // Wait for `bytes` in my FIFO and copy them to dst (any address space visible to the GPU).
  // Returns 0 or KRNL_TIMEOUT_STS_ERROR.
  __device__ uint32_t pull(void *dst, uint64_t bytes, uint64_t timeout_ns = 10ull * 1000 * 1000 * 1000) {
    char *heap = w_.window + static_cast<uint64_t>(w_.rank) * w_.heap_bytes;
    cuda::Ctrl *me = reinterpret_cast<cuda::Ctrl *>(heap);
    __shared__ unsigned long long s_tail;
    __shared__ int s_ok;
    __syncthreads();
    cuda::StrmPort *sp = &me->strm[port_];
    if (threadIdx.x == 0) {
      s_tail = sp->tail;
      s_ok = wait_ge(&sp->head, s_tail + bytes, timeout_ns) ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) return KRNL_TIMEOUT_STS_ERROR;
    const char *fifo = heap + w_.strm_off + static_cast<uint64_t>(port_) * w_.strm_cap;
    char *d = static_cast<char *>(dst);
    const unsigned long long tail = s_tail;
    const uint64_t mask = w_.strm_cap - 1;
    if (((tail | bytes | reinterpret_cast<uint64_t>(d)) & 15) == 0) {
      for (uint64_t i = static_cast<uint64_t>(threadIdx.x) * 16; i < bytes; i += static_cast<uint64_t>(blockDim.x) * 16)
        *reinterpret_cast<uint4 *>(d + i) = *reinterpret_cast<const uint4 *>(fifo + ((tail + i) & mask));
    } else {
      for (uint64_t i = threadIdx.x; i < bytes; i += blockDim.x) d[i] = fifo[(tail + i) & mask];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) dev::st_release_sys(&sp->tail, tail + bytes);
    return 0;
  }

  // Append `bytes` from src to the FIFO of rank dst_rank (default: my own).  Producers on different GPUs
  // reserve windows with a system-scope atomic and publish in reservation order.
  __device__ uint32_t push(const void *src, uint64_t bytes, int dst_rank = -1, uint64_t timeout_ns = 10ull * 1000 * 1000 * 1000) {
    const uint32_t dr = dst_rank < 0 ? w_.rank : static_cast<uint32_t>(dst_rank);
    char *dheap = w_.window + static_cast<uint64_t>(dr) * w_.heap_bytes;
    cuda::StrmPort *dp = &reinterpret_cast<cuda::Ctrl *>(dheap)->strm[port_];
    __shared__ unsigned long long s_start;
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long start;
      asm volatile("atom.relaxed.sys.global.add.u64 %0, [%1], %2;"
                   : "=l"(start) : "l"(&dp->reserve), "l"(static_cast<unsigned long long>(bytes)) : "memory");
      s_start = start;
      s_ok = bytes <= w_.strm_cap;
      uint32_t spins = 0;
      uint64_t t0 = 0;
      while (s_ok && start + bytes > dev::ld_acquire_sys(&dp->tail) + w_.strm_cap) {
        if (++spins > 16) dev::nanosleep(200);
        if ((spins & 0xFF) == 0) {
          const uint64_t now = dev::globaltimer_ns();
          if (!t0) t0 = now;
          else if (now - t0 > timeout_ns) s_ok = 0;
        }
      }
    }
    __syncthreads();
    char *fifo = dheap + w_.strm_off + static_cast<uint64_t>(port_) * w_.strm_cap;
    const char *s = static_cast<const char *>(src);
    const unsigned long long start = s_start;
    const uint64_t mask = w_.strm_cap - 1;
    const int ok = s_ok;
    if (ok) {
      if (((start | bytes | reinterpret_cast<uint64_t>(s)) & 15) == 0) {
        for (uint64_t i = static_cast<uint64_t>(threadIdx.x) * 16; i < bytes; i += static_cast<uint64_t>(blockDim.x) * 16)
          *reinterpret_cast<uint4 *>(fifo + ((start + i) & mask)) = *reinterpret_cast<const uint4 *>(s + i);
      } else {
        for (uint64_t i = threadIdx.x; i < bytes; i += blockDim.x) fifo[(start + i) & mask] = s[i];
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      // publish in reservation order (even after a timeout, so later producers are not blocked forever)
      wait_ge(&dp->head, start, timeout_ns);
      dev::st_release_sys(&dp->head, start + bytes);
    }
    return ok ? 0u : static_cast<uint32_t>(KRNL_TIMEOUT_STS_ERROR);
  }

  // bytes currently readable in my FIFO (one thread)
  __device__ uint64_t available() const {
    const cuda::Ctrl *me = reinterpret_cast<const cuda::Ctrl *>(w_.window + static_cast<uint64_t>(w_.rank) * w_.heap_bytes);
    return dev::ld_acquire_sys(&me->strm[port_].head) - dev::ld_acquire_sys(&me->strm[port_].tail);
  }

private:
  static __device__ __forceinline__ bool wait_ge(const unsigned long long *p, unsigned long long target, uint64_t timeout_ns) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (dev::ld_acquire_sys(p) < target) {
      if (++spins > 16) dev::nanosleep(200);
      if ((spins & 0xFF) == 0) {
        const uint64_t now = dev::globaltimer_ns();
        if (!t0) t0 = now;
        else if (now - t0 > timeout_ns) return false;
      }
    }
    return true;
  }
  cuda::DevWorld w_;
  uint32_t port_;
};

// ---------------------------------------------------------------------------------------------------------------
// Data port for TENSOR-CORE producers ("producer tile writes ... directly into peers"): a kernel that
// has a finished output tile staged in shared memory (128B-swizzled box of `owner_shard_map`, a CUtensorMap over the
// owner rank's shard in the peer-mapped symmetric heap) hands it to the reduce-scatter with one instruction: the TMA
// unit adds the box into the owner's memory over NVLink (cp.reduce.async.bulk.tensor .add, SASS UTMAREDG) — element
// type and accumulation precision are those of the tensor map (bf16 or fp32).  One thread issues; the staging buffer
// may be reused once emit_wait_read<N>() says at most N emitted boxes are still being read.
__device__ __forceinline__ void reduce_scatter_emit_tile(const void *owner_shard_map, const void *smem_tile, int col, int row) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(owner_shard_map)),
               "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_tile))), "r"(col), "r"(row)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void emit_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// all boxes emitted by this thread have been added at their owners
__device__ __forceinline__ void emit_flush() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

} // namespace device
} // namespace accl
