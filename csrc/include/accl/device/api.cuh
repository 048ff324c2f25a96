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
        comm_((comm_adr - exchmem::COMM_BASE) / exchmem::COMM_STRIDE), dpcfg_(dpcfg_adr), cflags_(cflags), sflags_(sflags) {}

  // Call from ONE thread.  Returns the ticket to pass to finalize_call.
  __device__ Ticket start_call(uint32_t scenario, uint32_t len, uint32_t comm, uint32_t root_src_dst, uint32_t function,
                               uint32_t tag, uint32_t datapath_cfg, uint32_t compression_flags, uint32_t stream_flags,
                               uint64_t addra, uint64_t addrb, uint64_t addrc) {
    const Ticket t = atomicAdd(&ctrl_->dev_tail, 1ull);
    // ring full: wait until the engine has consumed the entry that used this slot
    while (t >= dev::ld_acquire_sys(&ctrl_->dev_fetched) + cuda::RING_SLOTS) dev::nanosleep(200);
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
    ctrl_->dev_ring[t % cuda::RING_SLOTS] = d;
    __threadfence();
    dev::st_release_sys(&ctrl_->dev_ready[t % cuda::RING_SLOTS], t + 1);
    return t + 1;
  }

  // Blocks until the engine has retired the call; returns its error word.
  __device__ uint32_t finalize_call(Ticket ticket) {
    const unsigned long long *st = &ctrl_->dev_status[(ticket - 1) % cuda::RING_SLOTS];
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

  // asynchronous flavour: issue now, finalize later (lets a kernel overlap compute)
  __device__ Ticket all_reduce_async(uint32_t len, reduceFunction fn, uint64_t src, uint64_t dst) {
    return start_call(static_cast<uint32_t>(operation::allreduce), len, comm_, 0, static_cast<uint32_t>(fn), TAG_ANY, dpcfg_,
                      cflags_, sflags_, src, 0, dst);
  }

private:
  __device__ uint32_t run(operation op, uint32_t len, uint32_t root, uint32_t fn, uint32_t tag, uint64_t a, uint64_t b,
                          uint64_t c) {
    return finalize_call(start_call(static_cast<uint32_t>(op), len, comm_, root, fn, tag, dpcfg_, cflags_, sflags_, a, b, c));
  }
  cuda::Ctrl *ctrl_;
  uint32_t comm_, dpcfg_, cflags_, sflags_;
};

} // namespace device
} // namespace accl
