// run_work: execute one WorkItem with `nctas` cooperating CTAs.
#pragma once
#include "collectives.cuh"
#include "staged.cuh"
#include "compress.cuh"

namespace accl {
namespace cuda {
namespace k {

__device__ __noinline__ void run_work(const DevWorld &w, const WorkItem &it, int cta, int nctas, uint32_t *s_err) {
  __shared__ uint64_t s_off0[ACCL_MAX_RANKS], s_off2[ACCL_MAX_RANKS];
  __shared__ const char *s_slots[ACCL_MAX_RANKS];
  __shared__ uint32_t s_tmp;
  __shared__ PtrTable s_tab;
  Ctx c{w, it, cta, nctas, reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes), s_err,
        static_cast<uint64_t>(it.timeout_us) * 1000ull, &s_tab};
  const operation op = static_cast<operation>(it.desc.scenario);
  const bool eager = it.algo == ALGO_EAGER;
  if (it.algo == ALGO_LL || it.algo == ALGO_STAGED) {
    // one-way staged exchange (one hop, no meetings): staged.cuh
    switch (op) {
    case operation::allreduce: stg_collective(c, SP_ALLREDUCE, &s_tmp); return;
    case operation::reduce_scatter: stg_collective(c, SP_REDUCE_SCATTER, &s_tmp); return;
    case operation::allgather: stg_collective(c, SP_ALLGATHER, &s_tmp); return;
    case operation::bcast: stg_collective(c, SP_BCAST, &s_tmp); return;
    case operation::scatter: stg_collective(c, SP_SCATTER, &s_tmp); return;
    case operation::gather: stg_collective(c, SP_GATHER, &s_tmp); return;
    case operation::reduce: stg_collective(c, SP_REDUCE, &s_tmp); return;
    case operation::alltoall: stg_collective(c, SP_ALLTOALL, &s_tmp); return;
    default: break;
    }
  }
  if (it.algo == ALGO_WIRE) {
    wire_collective(c);
    return;
  }
  switch (op) {
  case operation::nop:
  case operation::config: break;
  case operation::copy: op_copy(c); break;
  case operation::combine: op_combine(c); break;
  case operation::send:
    if (eager) egr_send(c, &s_tmp);
    else rv_send(c, s_off0);
    break;
  case operation::recv:
    if (eager) egr_recv(c, &s_tmp);
    else rv_recv(c, s_off0);
    break;
  case operation::allreduce:
    if (eager) egr_collective(c, EP_ALLREDUCE, &s_tmp, s_slots);
    else if (it.algo == ALGO_P2P_ONESHOT) rv_allreduce_oneshot(c, s_off0, s_off2);
    else rv_allreduce(c, s_off0, s_off2);
    break;
  case operation::reduce_scatter:
    if (eager) egr_collective(c, EP_REDUCE_SCATTER, &s_tmp, s_slots);
    else rv_reduce_scatter(c, s_off0, s_off2);
    break;
  case operation::reduce:
    if (eager) egr_collective(c, EP_REDUCE, &s_tmp, s_slots);
    else rv_reduce(c, s_off0, s_off2);
    break;
  case operation::allgather:
    if (eager) egr_collective(c, EP_ALLGATHER, &s_tmp, s_slots);
    else rv_move(c, EP_ALLGATHER, s_off0, s_off2);
    break;
  case operation::bcast:
    if (eager) egr_collective(c, EP_BCAST, &s_tmp, s_slots);
    else rv_move(c, EP_BCAST, s_off0, s_off2);
    break;
  case operation::scatter:
    if (eager) egr_collective(c, EP_SCATTER, &s_tmp, s_slots);
    else rv_move(c, EP_SCATTER, s_off0, s_off2);
    break;
  case operation::gather:
    if (eager) egr_collective(c, EP_GATHER, &s_tmp, s_slots);
    else rv_move(c, EP_GATHER, s_off0, s_off2);
    break;
  case operation::alltoall:
    if (eager) egr_collective(c, EP_ALLTOALL, &s_tmp, s_slots);
    else rv_move(c, EP_ALLTOALL, s_off0, s_off2);
    break;
  case operation::barrier:
    chan_sync(c, false, 0, 0, nullptr, nullptr);
    break;
  default:
    if (threadIdx.x == 0) atomicOr(s_err, COLLECTIVE_NOT_IMPLEMENTED);
    break;
  }
}

// The data phase of a rendezvous collective with the buffer offsets already exchanged: what the engine's worker
// CTAs execute for an MV_BODY move (nothing in here waits for a peer).
__device__ __noinline__ void run_body(const DevWorld &w, const WorkItem &it, const uint64_t *off0, const uint64_t *off2, int cta,
                                      int nctas, uint32_t *s_err) {
  __shared__ PtrTable s_tab;
  Ctx c{w, it, cta, nctas, reinterpret_cast<Ctrl *>(w.window + static_cast<uint64_t>(w.rank) * w.heap_bytes), s_err,
        static_cast<uint64_t>(it.timeout_us) * 1000ull, &s_tab};
  switch (static_cast<operation>(it.desc.scenario)) {
  case operation::copy: op_copy(c); break;
  case operation::combine: op_combine(c); break;
  case operation::allreduce:
    if (it.algo == ALGO_P2P_ONESHOT) rvb_allreduce_oneshot(c, off0, off2);
    else rvb_allreduce(c, off0, off2);
    break;
  case operation::reduce_scatter: rvb_reduce_scatter(c, off0, off2); break;
  case operation::reduce: rvb_reduce(c, off0, off2, false); break;
  case operation::allgather: rvb_move(c, EP_ALLGATHER, off0, off2, false); break;
  case operation::bcast: rvb_move(c, EP_BCAST, off0, off2, false); break;
  case operation::scatter: rvb_move(c, EP_SCATTER, off0, off2, false); break;
  case operation::gather: rvb_move(c, EP_GATHER, off0, off2, false); break;
  case operation::alltoall: rvb_move(c, EP_ALLTOALL, off0, off2, false); break;
  default:
    if (threadIdx.x == 0) atomicOr(s_err, COLLECTIVE_NOT_IMPLEMENTED);
    break;
  }
}

} // namespace k
} // namespace cuda
} // namespace accl
