// Decoded call context shared by the emulator's data mover (engine.cpp) and
// firmware handlers (firmware.cpp).
#pragma once
#include "accl/emu/engine.hpp"
#include "accl/emu/softfloat.hpp"

namespace accl {
namespace emu {

struct Engine::Ctx {
  EmuCall *call = nullptr;
  operation op = operation::nop;
  uint32_t count = 0, root = 0, tag = TAG_ANY, func = 0;
  uint32_t cflags = 0, sflags = 0, hflags = 0;
  uint64_t a0 = 0, a1 = 0, a2 = 0;
  CommView comm;
  ArithView ar;
  uint32_t max_eager = 0, max_rndzv = 0, rxbuf_size = 0, spare_size = 0;
  uint64_t spare[3] = {0, 0, 0};
  bool eager = true; // protocol for this call
  bool eth_c = false;

  bool op0_stream() const { return sflags & static_cast<uint32_t>(streamFlags::OP0_STREAM); }
  bool res_stream() const { return sflags & static_cast<uint32_t>(streamFlags::RES_STREAM); }
  bool op0_c() const { return cflags & static_cast<uint32_t>(compressionFlags::OP0_COMPRESSED); }
  bool op1_c() const { return cflags & static_cast<uint32_t>(compressionFlags::OP1_COMPRESSED); }
  bool res_c() const { return cflags & static_cast<uint32_t>(compressionFlags::RES_COMPRESSED); }
  reduceFunction fn() const { return static_cast<reduceFunction>(func); }
  // bytes of n elements as stored in an operand buffer
  uint64_t op_bytes(bool compressed, uint64_t n) const {
    return repr_bytes(compressed ? ar.c : ar.u, n, ar.ratio_log);
  }
  uint32_t ubytes() const { return dtype_bytes(ar.u); }
  uint32_t stream_id() const { return tag == TAG_ANY ? 0 : tag; }
};

// Runs `fn` only if this step has not completed in an earlier attempt of a
// parked call.  fn returns false when it cannot make progress yet (and must
// then have had no side effects).
struct Steps {
  uint32_t &done;
  uint32_t idx = 0;
  bool blocked = false;
  explicit Steps(uint32_t &d) : done(d) {}
  template <typename F> bool operator()(F &&fn) {
    if (blocked) return false;
    if (idx++ < done) return true;
    if (!fn()) {
      blocked = true;
      return false;
    }
    ++done;
    return true;
  }
};

} // namespace emu
} // namespace accl
