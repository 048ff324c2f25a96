// Out-of-band bootstrap transport: the only thing ranks share before the
// symmetric heap exists.
//
// Role in the reference: MPI in the test harness (rank/size, barriers, QP
// exchange: driver/utils/accl_network_utils/accl_network_utils.cpp:39-58,
// 272-333) plus `generate_ranks` (:394-449).  Here it is part of the library so
// that the C++ API is usable without MPI or Python: ranks are either threads
// of one process (LocalOob), processes of a torchrun-style job (TcpOob, using
// MASTER_ADDR/MASTER_PORT), or anything able to allgather bytes (CallbackOob,
// used by the Python layer to ride on torch.distributed).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "accl/communicator.hpp"

namespace accl {

class Oob {
public:
  virtual ~Oob() = default;
  virtual int rank() const = 0;
  virtual int size() const = 0;
  // True when all ranks live in this process (driver handles can be shared
  // by value, no file-descriptor passing needed).
  virtual bool same_process() const = 0;
  // `out` receives size()*bytes, rank-major.
  virtual void allgather(const void *in, void *out, size_t bytes) = 0;

  void barrier();
  template <typename T> std::vector<T> allgather_value(const T &v) {
    std::vector<T> out(static_cast<size_t>(size()));
    allgather(&v, out.data(), sizeof(T));
    return out;
  }
  // Value of rank `root`, on every rank.
  template <typename T> T bcast_value(const T &v, int root = 0) {
    return allgather_value(v)[static_cast<size_t>(root)];
  }
};

// N ranks as threads of one process.  create() returns one endpoint per rank;
// each endpoint must be used by exactly one thread at a time.
class LocalOob : public Oob {
public:
  static std::vector<std::shared_ptr<Oob>> create(int world_size);
  struct Shared;
  LocalOob(std::shared_ptr<Shared> s, int rank) : s_(std::move(s)), rank_(rank) {}
  int rank() const override { return rank_; }
  int size() const override;
  bool same_process() const override { return true; }
  void allgather(const void *in, void *out, size_t bytes) override;

private:
  std::shared_ptr<Shared> s_;
  int rank_;
};

// One rank per process; rank 0 runs a tiny TCP rendezvous server.
class TcpOob : public Oob {
public:
  TcpOob(int rank, int size, const std::string &addr, int port,
         int timeout_ms = 60000);
  ~TcpOob() override;
  // Reads RANK / WORLD_SIZE / MASTER_ADDR / ACCL_PORT (or MASTER_PORT+port_offset).
  static std::shared_ptr<Oob> from_env(int port_offset = 37);
  int rank() const override { return rank_; }
  int size() const override { return size_; }
  bool same_process() const override { return false; }
  void allgather(const void *in, void *out, size_t bytes) override;

private:
  int rank_, size_;
  int listen_fd_ = -1;
  std::vector<int> peers_; // rank 0: fd per rank; others: peers_[0] = root fd
};

// Bootstrap over a user-supplied allgather (e.g. torch.distributed).
class CallbackOob : public Oob {
public:
  using AllgatherFn = std::function<void(const void *, void *, size_t)>;
  CallbackOob(int rank, int size, AllgatherFn fn)
      : rank_(rank), size_(size), fn_(std::move(fn)) {}
  int rank() const override { return rank_; }
  int size() const override { return size_; }
  bool same_process() const override { return false; }
  void allgather(const void *in, void *out, size_t bytes) override {
    fn_(in, out, bytes);
  }

private:
  int rank_, size_;
  AllgatherFn fn_;
};

// Pass open file descriptors between the processes of a job over abstract
// AF_UNIX sockets (SCM_RIGHTS).  Used to share CUDA VMM allocation handles
// and the NVLS multicast object.  all_to_all_fds(): every rank contributes
// one fd and receives everybody's (its own slot is a dup of its own fd).
std::vector<int> exchange_fds(Oob &oob, int my_fd, const std::string &channel);

// ---- rank tables (reference accl_network_utils::generate_ranks / get_ips, accl_network_utils.cpp:394-449)
// IPs from a JSON configuration file of the form {"ips": ["10.0.0.1", "10.0.0.2", ...]} (the reference's -c file).
std::vector<std::string> get_ips(const std::string &config_file);
// local: everybody on 127.0.0.1; otherwise the reference's 10.10.10.<rank + 1> convention
std::vector<std::string> get_ips(bool local, int world_size);
// rank i = {ips[i], base_port + i, session id i, max eager segment rxbuf_size}
std::vector<rank_t> generate_ranks(const std::vector<std::string> &ips, int base_port, addr_t rxbuf_size);
std::vector<rank_t> generate_ranks(bool local, int world_size, int base_port, addr_t rxbuf_size);
std::vector<rank_t> generate_ranks(const std::string &config_file, int base_port, addr_t rxbuf_size);

} // namespace accl
