// The engine as a separate process, and the driver-side device that talks to it.
//
//   driver process                         engine process (build/bin/cclo_emu)
//   ACCL -> RemoteDevice  ==== TCP ====>  EngineServer -> emu::Engine <-> SocketFabric <-> other engines
//
// Counterpart of the reference's split between SimDevice / SimBuffer in the driver
// (driver/xrt/src/simdevice.cpp, include/accl/simbuffer.hpp) and the `cclo_emu` process with its
// ZMQ control server (test/model/emulator/cclo_emu.cpp:510-537, test/model/zmq/zmq_server.cpp):
// the same request kinds (MMIO read / write, device-memory alloc / read / write, call, stream
// push / pull), here as fixed 32-byte binary frames over a loopback TCP connection instead of JSON
// over ZMQ, and with asynchronous completion events so any number of calls may be in flight.
#pragma once
#include <atomic>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "accl/cclo.hpp"
#include "accl/emu/engine.hpp"
#include "accl/request.hpp"

namespace accl {
namespace emu {

namespace wire {
enum Type : uint32_t {
  MMIO_READ = 1,   // a = byte offset                 -> a = value
  MMIO_WRITE = 2,  // a = byte offset, b = value
  MEM_ALLOC = 3,   // a = bytes, b = host arena?       -> a = address
  MEM_FREE = 4,    // a = address
  MEM_WRITE = 5,   // a = address, payload
  MEM_READ = 6,    // a = address, b = bytes           -> payload
  CALL = 7,        // a = call id, payload = CallDesc  -> ack; later EVENT_DONE
  KRNL_PUSH = 8,   // payload (kernel -> engine stream)
  KRNL_PULL = 9,   // a = stream id, b = bytes, aux = timeout ms -> a = ok, payload
  DEBUG_STATE = 10, //                                  -> payload = text
  LOOPBACK = 11,   // a = on / off
  SHUTDOWN = 12,
  REPLY = 0x8000,      // reply to request `seq`
  EVENT_DONE = 0x8001, // a = call id, b = retcode, aux = unused, payload = u64 duration ns
};
struct Frame {
  uint32_t type = 0;
  uint32_t seq = 0;
  uint64_t a = 0;
  uint64_t b = 0;
  uint32_t len = 0; // payload bytes following the frame
  uint32_t aux = 0;
};
static_assert(sizeof(Frame) == 32, "wire frame is 32 bytes");
} // namespace wire

// Serves one driver connection on addr:port until SHUTDOWN or disconnect.
class EngineServer {
public:
  EngineServer(std::shared_ptr<Engine> engine, const std::string &addr, int port);
  ~EngineServer();
  void serve(); // blocks
  int port() const { return port_; }

private:
  // the driver connection; shared with completion hooks that may outlive the server object
  // (calls still queued in the engine when the driver disconnects)
  struct Conn {
    std::mutex m;
    int fd = -1;
    void send(const wire::Frame &f, const void *payload);
    void close();
  };
  std::shared_ptr<Engine> engine_;
  std::shared_ptr<Conn> conn_;
  int listen_fd_ = -1, port_;
};

// CCLO backend living in the driver: every operation is a request to the engine process.
class RemoteDevice : public CCLO {
public:
  RemoteDevice(const std::string &addr, int port, int global_rank, int world, int connect_timeout_s = 60);
  ~RemoteDevice() override;

  ACCLRequest *call(const Options &options) override;
  ACCLRequest *start(const Options &options) override;
  val_t read(addr_t offset) override;
  void write(addr_t offset, val_t val) override;
  void wait(ACCLRequest *request) override;
  bool wait(ACCLRequest *request, std::chrono::milliseconds timeout) override;
  bool test(ACCLRequest *request) override;
  void free_request(ACCLRequest *request) override;
  val_t get_retcode(ACCLRequest *request) override;
  uint64_t get_duration(ACCLRequest *request) override;
  deviceType get_device_type() override { return deviceType::emulator; }
  std::string describe() override;
  void printDebug() override;
  std::shared_ptr<BufferStorage> allocate(size_t bytes, bufferKind kind) override;
  std::shared_ptr<BufferStorage> wrap_host(void *host_ptr, size_t bytes) override;
  void attach(int world_size, int local_rank) override;

  // raw services (also used by the storage objects and the stream-port helpers)
  uint64_t mem_alloc(size_t bytes, bool host);
  void mem_free(uint64_t addr);
  void mem_write(uint64_t addr, const void *src, size_t len);
  void mem_read(uint64_t addr, void *dst, size_t len);
  void kernel_push(const void *data, size_t bytes);
  bool kernel_pull(uint32_t strm, void *out, size_t bytes, int timeout_ms);
  void set_kernel_loopback(bool on);
  std::string debug_state();
  void shutdown_engine(); // ask the engine process to exit

private:
  struct Pending {
    bool done = false;
    wire::Frame reply;
    std::vector<uint8_t> payload;
  };
  wire::Frame rpc(wire::Frame f, const void *payload, std::vector<uint8_t> *reply_payload = nullptr);
  void reader_loop();

  int fd_ = -1, rank_, world_, port_;
  std::string addr_;
  std::mutex tx_m_, p_m_;
  std::condition_variable p_cv_;
  std::map<uint32_t, Pending> pending_;
  uint32_t next_seq_ = 1;
  std::atomic<bool> stop_{false}, broken_{false};
  std::shared_ptr<std::atomic<bool>> alive_ = std::make_shared<std::atomic<bool>>(true);
  std::thread reader_;
  RequestRegistry requests_;
  std::mutex calls_m_;
  std::map<uint64_t, std::shared_ptr<BaseRequest>> calls_;
  uint64_t next_call_ = 1;
};

} // namespace emu
} // namespace accl
