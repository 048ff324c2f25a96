// run_work: execute one WorkItem with `nctas` cooperating CTAs.
#pragma once
#include "collectives.cuh"

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

} // namespace k
} // namespace cuda
} // namespace accl
