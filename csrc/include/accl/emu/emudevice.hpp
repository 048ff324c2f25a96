// EmuDevice: the CCLO backend that drives a CPU engine model (engine.hpp).
// Ranks are either threads of one process sharing an InProcFabric, or one
// process each over a SocketFabric.
//
// Reference counterpart: SimDevice + SimBuffer (driver/xrt/src/simdevice.cpp,
// include/accl/simbuffer.hpp) talking ZMQ to a separate cclo_emu process;
// here the engine lives in the driver's process and only the rank-to-rank
// "network" may cross process boundaries.
#pragma once
#include <memory>

#include "accl/cclo.hpp"
#include "accl/emu/engine.hpp"
#include "accl/request.hpp"

namespace accl {
namespace emu {

class EmuDevice : public CCLO {
public:
  EmuDevice(std::shared_ptr<Fabric> fabric, int global_rank, int world, size_t dev_mem_bytes = 256u << 20,
            size_t host_mem_bytes = 256u << 20);
  ~EmuDevice() override;

  ACCLRequest *call(const Options &options) override;
  ACCLRequest *start(const Options &options) override;
  val_t read(addr_t offset) override { return engine_->read_exch(static_cast<uint32_t>(offset)); }
  void write(addr_t offset, val_t val) override { engine_->write_exch(static_cast<uint32_t>(offset), val); }
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

  Engine &engine() { return *engine_; }
  std::shared_ptr<Engine> engine_ptr() { return engine_; }

private:
  std::shared_ptr<Fabric> fabric_;
  std::shared_ptr<Engine> engine_;
  RequestRegistry requests_;
  int rank_, world_;
};

// Convenience: N in-process devices sharing one fabric (ranks as threads).
std::vector<std::unique_ptr<EmuDevice>> make_inproc_world(int world, size_t dev_mem_bytes = 256u << 20);

} // namespace emu
} // namespace accl
