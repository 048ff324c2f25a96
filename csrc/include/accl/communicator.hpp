// Communicators: an ordered set of ranks with per-peer sequence counters,
// stored in exchange memory where the engine reads (and advances) them.
//
// Reference: Communicator / rank_t (driver/xrt/include/accl/communicator.hpp,
// src/communicator.cpp:25-117).  A rank here is identified by the GPU it
// drives (device ordinal) and its global rank id ("session" in the table:
// the route key the engine uses to reach that peer's heap) instead of an
// IP/port/TCP-session triple; ip/port are kept as free-form fields so JSON
// rank files of the reference still load.
#pragma once
#include <string>
#include <vector>

#include "accl/cclo.hpp"
#include "accl/constants.hpp"

namespace accl {

struct rank_t {
  std::string ip = "127.0.0.1"; // informational on NVLink; socket address in the emulator
  int port = 0;
  int session_id = 0;           // global rank id (route key)
  addr_t max_segment_size = 0;  // largest eager segment this peer accepts
  rank_t() = default;
  rank_t(std::string ip_, int port_, int session, addr_t max_seg)
      : ip(std::move(ip_)), port(port_), session_id(session), max_segment_size(max_seg) {}
};

class Communicator {
public:
  // Serialises the table into exchange memory at `comm_index`.
  Communicator(CCLO *cclo, const std::vector<rank_t> &ranks, unsigned int local_rank, unsigned int comm_index);

  unsigned int local_rank() const { return local_rank_; }
  unsigned int size() const { return static_cast<unsigned int>(ranks_.size()); }
  unsigned int index() const { return index_; }
  // byte offset of this communicator's table in exchange memory; handed to
  // device-side kernels (get_communicator_addr)
  addr_t communicators_addr() const;
  const std::vector<rank_t> &get_ranks() const { return ranks_; }

  // re-read sequence numbers etc. from the engine
  void readback();
  std::string dump();

private:
  CCLO *cclo_;
  std::vector<rank_t> ranks_;
  unsigned int local_rank_;
  unsigned int index_;
  std::vector<uint32_t> inbound_seq_, outbound_seq_;
};

} // namespace accl
