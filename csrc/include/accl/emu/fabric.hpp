// The emulator's "network": moves Packets (a fixed message header + payload)
// between engines.  InProcFabric connects engines living in one process
// (ranks as threads); SocketFabric connects one engine per process over
// loopback TCP so torchrun-style multi-process jobs run without a GPU.
//
// Reference counterparts: the 64-byte ACCL message header `eth_header`
// (kernels/cclo/hls/eth_intf/eth_intf.h:114-151), the dummy protocol stacks
// (kernels/plugins/dummy_tcp_stack, dummy_cyt_rdma_stack) and the ZMQ pub/sub
// "Ethernet" of the emulator (test/model/zmq/zmq_server.cpp:31-190).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace accl {
namespace emu {

enum class MsgType : uint32_t {
  EGR_MSG = 0,        // eager payload, lands in a spare RX buffer (or a stream when strm != 0)
  RNDZVS_MSG = 1,     // one-sided write of payload to `vaddr`
  RNDZVS_INIT = 2,    // receiver -> sender: "my buffer for (tag) is at vaddr"
  RNDZVS_WR_DONE = 3, // sender -> receiver: the write for (tag) has completed
  RNDZVS_CANCEL = 4   // receiver -> sender: forget my address note for (tag, vaddr) — the recv that posted it timed out
};

struct MsgHeader {
  uint32_t count = 0;     // payload bytes (EGR/RNDZVS_MSG) or announced bytes (INIT)
  uint32_t tag = 0;
  uint32_t src = 0;       // sender's global rank id
  uint32_t seqn = 0;
  uint32_t strm = 0;      // !=0: deliver to this device-side stream instead of an RX buffer
  uint32_t dst = 0;       // receiver's global rank id
  uint32_t msg_type = 0;
  uint32_t host = 0;      // target buffer lives in host memory
  uint64_t vaddr = 0;     // rendezvous target address
  uint32_t comm_sig = 0;  // signature of the communicator the message belongs to
  uint32_t elems = 0;     // element count of the payload (block-scaled wires need it)
  uint32_t dtypes = 0;    // wire dtype | element dtype << 8 | ratio_log << 16 (sanity check at the receiver)
};

struct Packet {
  MsgHeader hdr;
  std::vector<uint8_t> payload;
};

class Fabric {
public:
  using Handler = std::function<void(Packet &&)>;
  virtual ~Fabric() = default;
  // deliver packets addressed to `global_rank` to `h` (called from fabric threads)
  virtual void attach(int global_rank, Handler h) = 0;
  virtual void detach(int global_rank) = 0;
  virtual void send(Packet &&p) = 0; // routed by p.hdr.dst
  virtual int world_size() const = 0;
  virtual const char *name() const = 0;
};

// All ranks in one process; send() calls the destination handler directly
// (the handler only enqueues, so this never blocks on the peer's progress).
class InProcFabric : public Fabric {
public:
  explicit InProcFabric(int world) : handlers_(static_cast<size_t>(world)) {}
  void attach(int r, Handler h) override;
  void detach(int r) override;
  void send(Packet &&p) override;
  int world_size() const override { return static_cast<int>(handlers_.size()); }
  const char *name() const override { return "inproc"; }

private:
  std::mutex m_;
  std::condition_variable cv_;
  std::vector<Handler> handlers_;
};

// One rank per process, full mesh of loopback TCP connections.  Rank r
// listens on base_port + r.
class SocketFabric : public Fabric {
public:
  SocketFabric(int my_rank, int world, const std::string &addr, int base_port);
  ~SocketFabric() override;
  void attach(int r, Handler h) override;
  void detach(int r) override;
  void send(Packet &&p) override;
  int world_size() const override { return world_; }
  const char *name() const override { return "socket"; }

private:
  void rx_loop(int fd);
  int connect_to(int peer);
  int me_, world_;
  std::string addr_;
  int base_port_;
  int listen_fd_ = -1;
  std::vector<int> tx_fd_;
  std::vector<std::unique_ptr<std::mutex>> tx_m_;
  std::vector<std::thread> rx_threads_;
  std::thread accept_thread_;
  std::atomic<bool> stop_{false};
  std::mutex hm_;
  std::condition_variable hcv_;
  Handler handler_;
};

} // namespace emu
} // namespace accl
