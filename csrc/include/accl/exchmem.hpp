// Exchange memory: the one configuration block shared by the host API and
// the engine (emulator control thread or GPU control CTA).  The host writes
// it through CCLO::write / reads through CCLO::read; the engine reads
// communicators, arithmetic configs, eager-buffer geometry and tuning values
// from it and updates sequence numbers, RETCODE and PERFCNT.
//
// Concept and usage follow the reference (8 KB block, layout in
// kernels/cclo/fw/sw_apps/ccl_offload_control/src/ccl_offload_control.h:85-98,
// 292-323; written by accl.cpp:1131-1208 and communicator.cpp:25-52).  The
// word layout below is this library's own: fixed-offset sections instead of a
// bump-allocated tail, 64-bit addresses everywhere, bf16/fp8-capable arith
// entries.  All offsets are BYTE offsets; all fields are 32-bit words.
#pragma once
#include <cstdint>

#include "accl/common.hpp"
#include "accl/constants.hpp"

namespace accl {
namespace exchmem {

constexpr uint32_t SIZE_BYTES = 8192;
constexpr uint32_t SIZE_WORDS = SIZE_BYTES / 4;

// ---- scalar registers
constexpr uint32_t HWID = 0x000;       // capability bits (enum capability)
constexpr uint32_t CFGRDY = 0x004;     // 1 once initialize() finished
constexpr uint32_t RETCODE = 0x008;    // error word of the last finished call
constexpr uint32_t PERFCNT_LO = 0x00C; // duration of the last call, ns
constexpr uint32_t PERFCNT_HI = 0x010;
constexpr uint32_t TIMEOUT = 0x014;    // engine wait budget (µs on GPU, polls in emu)
constexpr uint32_t MAX_EAGER_SIZE = 0x018;
constexpr uint32_t MAX_RENDEZVOUS_SIZE = 0x01C;
constexpr uint32_t EAGER_RX_BUF_SIZE = 0x020;
constexpr uint32_t EAGER_RX_BUF_COUNT = 0x024;
constexpr uint32_t GATHER_FLAT_TREE_MAX_FANIN = 0x028;
constexpr uint32_t GATHER_FLAT_TREE_MAX_COUNT = 0x02C;
constexpr uint32_t BCAST_FLAT_TREE_MAX_RANKS = 0x030;
constexpr uint32_t REDUCE_FLAT_TREE_MAX_RANKS = 0x034;
constexpr uint32_t REDUCE_FLAT_TREE_MAX_COUNT = 0x038;
constexpr uint32_t NUM_COMMUNICATORS = 0x03C;
constexpr uint32_t NUM_ARITHCFG = 0x040;
constexpr uint32_t PKT_ENABLED = 0x044;     // data plane enabled (cfgFunc::enable_pkt)
constexpr uint32_t SPARE_BUF_SIZE = 0x048;  // bytes of each rendezvous scratch buffer
constexpr uint32_t ONE_HOP_SCHEDULES = 0x04C; // emulator: 1 = all-gather / reduce-scatter / all-reduce as one-hop exchanges and
                                              // rooted collectives in their flat forms (the schedules the B200 backend runs)
constexpr uint32_t SPARE_BUF_BASE = 0x050;  // 3 x {addr lo, addr hi}
constexpr uint32_t NUM_SPARE_BUFS = 3;

// ---- arithmetic configurations: ARITHCFG_WORDS (8) words each
constexpr uint32_t ARITHCFG_BASE = 0x080;
constexpr uint32_t ARITHCFG_STRIDE = 32;
constexpr uint32_t MAX_ARITHCFG = 28;
// word indices inside one entry
enum ArithWord { AC_UNCOMPRESSED_BYTES = 0, AC_COMPRESSED_BYTES = 1, AC_RATIO_LOG = 2, AC_COMPRESSOR = 3,
                 AC_DECOMPRESSOR = 4, AC_ARITH_COMPRESSED = 5, AC_FN_SUM = 6, AC_FN_MAX = 7 };

// ---- eager RX buffer table: 8 words each
constexpr uint32_t RXBUF_BASE = 0x400;
constexpr uint32_t RXBUF_STRIDE = 32;
constexpr uint32_t MAX_RXBUFS = 64;
enum RxWord { RX_STATUS = 0, RX_ADDR_LO = 1, RX_ADDR_HI = 2, RX_MAX_LEN = 3, RX_TAG = 4, RX_LEN = 5, RX_SRC = 6, RX_SEQ = 7 };
enum RxStatus : uint32_t { RX_IDLE = 0, RX_ENQUEUED = 1, RX_RESERVED = 2, RX_ERROR = 3 };

// ---- communicators: header {size, local_rank} + 6 words per rank
constexpr uint32_t COMM_BASE = 0xC00;
constexpr uint32_t COMM_RANK_WORDS = 6;
constexpr uint32_t COMM_STRIDE = (2 + ACCL_MAX_RANKS * COMM_RANK_WORDS) * 4; // 392 bytes
enum CommRankWord { CR_ADDR = 0,       // "ip": device ordinal / encoded address
                    CR_PORT = 1,
                    CR_INBOUND_SEQ = 2,
                    CR_OUTBOUND_SEQ = 3,
                    CR_SESSION = 4,    // global rank id of this member (route key)
                    CR_MAX_SEG = 5 };
static_assert(COMM_BASE + ACCL_MAX_COMMUNICATORS * COMM_STRIDE <= SIZE_BYTES, "exchange memory overflow");
static_assert(ARITHCFG_BASE + MAX_ARITHCFG * ARITHCFG_STRIDE <= RXBUF_BASE, "arith table overflow");
static_assert(RXBUF_BASE + MAX_RXBUFS * RXBUF_STRIDE <= COMM_BASE, "rx table overflow");

ACCL_HD uint32_t comm_offset(uint32_t comm_idx) { return COMM_BASE + comm_idx * COMM_STRIDE; }
ACCL_HD uint32_t comm_rank_offset(uint32_t comm_idx, uint32_t r, uint32_t word) {
  return comm_offset(comm_idx) + (2 + r * COMM_RANK_WORDS + word) * 4;
}
ACCL_HD uint32_t arith_offset(uint32_t idx, uint32_t word) { return ARITHCFG_BASE + idx * ARITHCFG_STRIDE + word * 4; }
ACCL_HD uint32_t rxbuf_offset(uint32_t idx, uint32_t word) { return RXBUF_BASE + idx * RXBUF_STRIDE + word * 4; }

} // namespace exchmem
} // namespace accl
