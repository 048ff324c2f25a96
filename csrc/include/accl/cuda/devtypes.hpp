// Structures shared by the CUDA backend's host code and its sm_100a kernels:
// geometry of the symmetric heap's control block, the work item a kernel
// executes, completion records and the persistent engine's rings.
//
// Mapping to the reference: `Ctrl` is the device-resident equivalent of the
// CCLO's exchange memory + RX-buffer metadata + rendezvous mailbox
// (kernels/cclo/fw/.../ccl_offload_control.h:85-98,292-323); `WorkItem` is the
// decoded 15-word call (ccl_offload_control.c:2319-2373); `Completion` is
// RETVAL + PERFCNT (:2291-2306).
#pragma once
#include <cstdint>

#include "accl/cclo.hpp"
#include "accl/constants.hpp"
#include "accl/exchmem.hpp"

namespace accl {
namespace cuda {

constexpr int MAX_CH = 160;       // sync channels == max CTAs cooperating on one call (one per SM and a few spare)
constexpr int EGR_CH = 16;        // channels usable by eager (slot-based) transfers
constexpr int EGR_DEPTH_MAX = 16; // slots per (channel, src): the eager RX buffers
constexpr int N_REQ_SLOTS = 256;  // completion records
constexpr int RING_SLOTS = 128;   // command ring depth (host ring and device ring each)
constexpr uint64_t CTRL_BYTES = 1ull << 20; // control block at the start of every heap
constexpr uint64_t INVALID_OFF = ~0ull;

enum Algo : uint32_t {
  ALGO_AUTO = 0,
  ALGO_LOCAL = 1,        // world == 1 / local-only primitive
  ALGO_EAGER = 2,        // push into the peers' eager slots, consume locally (one network hop)
  ALGO_NVLS = 3,         // multimem.ld_reduce / multimem.st through the NVSwitch
  ALGO_P2P = 4,          // peer loads/stores on the mapped heaps (one-shot or two-shot by op)
  ALGO_P2P_ONESHOT = 5   // every rank pulls everything (small allreduce without slots)
};

struct SyncRec { // rendezvous "address exchange" record, written by a peer next to its signal
  uint64_t off0;
  uint64_t off2;
  uint32_t kind; // operation code of the sender's call: mismatches are protocol errors
  uint32_t pad;
};

struct EgrHdr { // per-slot message header (the reference's eth_header, minus routing)
  uint32_t tag;
  uint32_t bytes;
  uint32_t elems;
  uint32_t kind; // scenario | wire dtype << 8
};

struct Completion {
  uint32_t retcode;
  uint32_t done_ctas;
  unsigned long long t_start; // %globaltimer ns (min over CTAs)
  unsigned long long t_end;   // max over CTAs
  uint32_t seq;               // written last: request sequence this record belongs to
  uint32_t pad;
};

// Layout of the first CTRL_BYTES of every rank's heap (same offsets everywhere).
struct Ctrl {
  uint32_t exch[exchmem::SIZE_WORDS]; // exchange memory (host <-> engine configuration block)
  // ---- written by peers
  uint32_t sig[MAX_CH][ACCL_MAX_RANKS];
  SyncRec rec[MAX_CH][ACCL_MAX_RANKS];
  uint32_t egr_sig[EGR_CH][ACCL_MAX_RANKS]; // messages arrived from src on channel
  uint32_t egr_ack[EGR_CH][ACCL_MAX_RANKS]; // my messages consumed by dst (credits)
  EgrHdr egr_hdr[EGR_CH][EGR_DEPTH_MAX][ACCL_MAX_RANKS];
  // ---- local protocol state (single writer: the CTA owning the channel)
  uint32_t sent[MAX_CH][ACCL_MAX_RANKS];
  uint32_t expect[MAX_CH][ACCL_MAX_RANKS];
  uint32_t egr_sent[EGR_CH][ACCL_MAX_RANKS];
  uint32_t egr_expect[EGR_CH][ACCL_MAX_RANKS];
  // ---- completion records
  Completion comp[N_REQ_SLOTS];
  // ---- persistent engine (engine.cu): configuration, rings, doorbells
  uint32_t plan_cfg_words[8];     // PlanCfg image (plan.hpp) for device-side planning
  uint32_t engine_timeout_us;
  uint32_t engine_exit;           // workers leave when set
  unsigned long long host_tail;   // stream-ordered doorbell (cuStreamWriteValue64 target)
  unsigned long long host_fetched; // host-ring entries consumed so far (survives relaunches)
  unsigned long long dev_tail;    // device-side producers take tickets here
  unsigned long long dev_fetched;
  unsigned long long issue_tail;  // control -> workers
  unsigned long long done_count;  // completed work items (in-order retirement)
  unsigned long long host_done;   // host-ring entries retired (cuStreamWaitValue64 target)
  unsigned long long dev_ready[RING_SLOTS];  // slot published: ticket + 1
  unsigned long long dev_status[RING_SLOTS]; // slot finished: (ticket + 1) | retcode << 32
  CallDesc dev_ring[RING_SLOTS];
  // ---- device-side stream port (OP0_STREAM / RES_STREAM operands, stream_put)
  unsigned long long strm_head;    // bytes published (readable by the consumer)
  unsigned long long strm_tail;    // bytes consumed
  unsigned long long strm_reserve; // bytes reserved by producers (local kernels or peers doing stream_put)
  uint32_t strm_err;               // sticky error bits raised by stream helper kernels
  uint32_t strm_pad;
#ifdef ACCL_EXPERIMENTAL_BCAST_FLAGS
  // one-way "chunk landed" counters of the flag-driven pipelined broadcast (docs/roadmap.md #2)
  uint32_t step_sig[MAX_CH][ACCL_MAX_RANKS];  // written by peers: chunks src has delivered to me on this channel
  uint32_t step_seen[MAX_CH][ACCL_MAX_RANKS]; // local: how many of them earlier calls already consumed
  uint32_t step_sent[MAX_CH][ACCL_MAX_RANKS]; // local: chunks I have delivered to dst
#endif
#ifdef ACCL_PHASE_TIMING
  // opt-in instrumentation (channel 0 only): where a call's time goes.  Read with CudaDevice::debug_state().
  unsigned long long dbg_calls;      // calls executed
  unsigned long long dbg_kernel_ns;  // sum of kernel body durations (run_work entry -> exit)
  unsigned long long dbg_sync_ns;    // of which spent inside chan_sync / pair_sync (flag round trips)
  unsigned long long dbg_syncs;      // number of meetings
  unsigned long long dbg_wait_ns;    // flag waits of any kind (wait_ge), summed over the waiting threads of channel 0
  unsigned long long dbg_waits;      // number of such waits
#endif
};
static_assert(sizeof(Ctrl) <= CTRL_BYTES / 2, "control block too large");

struct DevWorld {
  char *window;        // rank r's heap lives at window + r * heap_bytes
  char *mc;            // NVLS multicast alias of heap offset 0 on every rank (nullptr: none)
  uint64_t heap_bytes;
  uint32_t world;      // ranks sharing the heap
  uint32_t rank;       // my global rank
  uint64_t egr_off;    // heap offset of the eager slot area
  uint32_t egr_depth;  // slots per (channel, src)
  uint32_t egr_slot_bytes;
  uint64_t strm_off;   // heap offset of the device-side stream FIFO (same on every rank)
  uint64_t strm_cap;   // its capacity in bytes (power of two)
#ifdef ACCL_EXPERIMENTAL_REDUCE_PUSH
  uint64_t scr_off;    // heap offset of the collective scratch region (same on every rank)
  uint64_t scr_bytes;
#endif
};

struct WorkItem {
  CallDesc desc;
  uint32_t comm_size, comm_rank;
  uint8_t members[ACCL_MAX_RANKS]; // communicator rank -> global rank
  uint32_t algo;
  uint32_t n_ctas;
  uint32_t udtype, cdtype, ratio_log, arith_compressed;
  uint32_t req_slot, req_seq;
  uint32_t timeout_us;
  uint32_t flags;
  uint64_t scratch_off, scratch_bytes; // per-call scratch inside the heap (compression / staging)
  uint64_t hc_ptr;     // engine: device-visible HostCompletion of a host-issued call (0: none)
  uint64_t dev_ticket; // engine: ticket + 1 of a device-issued call (0: host-issued)
  uint64_t host_seq;   // engine: position + 1 in the host ring (0: device-issued)
};
constexpr int ISSUE_SLOTS = 4;

enum WorkFlags : uint32_t {
  WF_USE_MC = 1u << 0,
  WF_ENGINE = 1u << 1,
  WF_CHAIN = 1u << 2 // not the last kernel of a lowered call: park the error word instead of publishing completion
};

// eager slot addressing inside a heap
ACCL_HD uint64_t egr_slot_off(const DevWorld &w, uint32_t ch, uint32_t slot, uint32_t src) {
  return w.egr_off + ((static_cast<uint64_t>(ch) * w.egr_depth + slot) * w.world + src) * w.egr_slot_bytes;
}
ACCL_HD uint64_t egr_area_bytes(uint32_t world, uint32_t depth, uint32_t slot_bytes) {
  return static_cast<uint64_t>(EGR_CH) * depth * world * slot_bytes;
}

} // namespace cuda
} // namespace accl
