#include "accl/communicator.hpp"

#include <sstream>

#include "accl/common.hpp"
#include "accl/exchmem.hpp"

namespace accl {

Communicator::Communicator(CCLO *cclo, const std::vector<rank_t> &ranks, unsigned int local_rank,
                           unsigned int comm_index)
    : cclo_(cclo), ranks_(ranks), local_rank_(local_rank), index_(comm_index) {
  if (comm_index >= static_cast<unsigned int>(ACCL_MAX_COMMUNICATORS))
    throw std::out_of_range("communicator index out of range");
  if (ranks.size() > static_cast<size_t>(ACCL_MAX_RANKS)) throw std::out_of_range("too many ranks in communicator");
  const uint32_t base = exchmem::comm_offset(comm_index);
  cclo_->write(base, static_cast<val_t>(ranks_.size()));
  cclo_->write(base + 4, local_rank_);
  for (uint32_t r = 0; r < ranks_.size(); ++r) {
    uint32_t ip = 0;
    try {
      ip = ip_encode(ranks_[r].ip);
    } catch (...) {
      ip = 0;
    }
    cclo_->write(exchmem::comm_rank_offset(comm_index, r, exchmem::CR_ADDR), ip);
    cclo_->write(exchmem::comm_rank_offset(comm_index, r, exchmem::CR_PORT), static_cast<val_t>(ranks_[r].port));
    cclo_->write(exchmem::comm_rank_offset(comm_index, r, exchmem::CR_INBOUND_SEQ), 0);
    cclo_->write(exchmem::comm_rank_offset(comm_index, r, exchmem::CR_OUTBOUND_SEQ), 0);
    cclo_->write(exchmem::comm_rank_offset(comm_index, r, exchmem::CR_SESSION),
                 static_cast<val_t>(ranks_[r].session_id));
    cclo_->write(exchmem::comm_rank_offset(comm_index, r, exchmem::CR_MAX_SEG),
                 static_cast<val_t>(ranks_[r].max_segment_size));
  }
  inbound_seq_.assign(ranks_.size(), 0);
  outbound_seq_.assign(ranks_.size(), 0);
}

addr_t Communicator::communicators_addr() const { return exchmem::comm_offset(index_); }

void Communicator::readback() {
  for (uint32_t r = 0; r < ranks_.size(); ++r) {
    inbound_seq_[r] = cclo_->read(exchmem::comm_rank_offset(index_, r, exchmem::CR_INBOUND_SEQ));
    outbound_seq_[r] = cclo_->read(exchmem::comm_rank_offset(index_, r, exchmem::CR_OUTBOUND_SEQ));
  }
}

std::string Communicator::dump() {
  readback();
  std::ostringstream o;
  o << "local rank: " << local_rank_ << " \t number of ranks: " << ranks_.size() << "\n";
  for (uint32_t r = 0; r < ranks_.size(); ++r) {
    o << "> rank " << r << " (ip " << ranks_[r].ip << ":" << ranks_[r].port << " ; session " << ranks_[r].session_id
      << " ; max segment size " << ranks_[r].max_segment_size << ") : \t <- inbound_seq " << inbound_seq_[r]
      << ", -> outbound_seq " << outbound_seq_[r] << "\n";
  }
  return o.str();
}

} // namespace accl
