#include "accl/emu/emudevice.hpp"

#include <cstring>
#include <sstream>

#include "accl/common.hpp"

namespace accl {
namespace emu {

namespace {
// host mirror + simulated device allocation
class EmuStorage : public BufferStorage {
public:
  EmuStorage(std::shared_ptr<Engine> e, size_t bytes, bufferKind kind, void *wrap)
      : engine_(std::move(e)), bytes_(bytes), kind_(kind) {
    addr_ = engine_->mem_alloc(std::max<size_t>(bytes, 1), kind == bufferKind::host_only);
    if (wrap) host_ = static_cast<uint8_t *>(wrap);
    else if (kind == bufferKind::device) {
      own_.assign(bytes, 0);
      host_ = own_.data();
    }
  }
  ~EmuStorage() override {
    try {
      engine_->mem_free(addr_);
    } catch (...) {
    }
  }
  void *host_ptr() override {
    if (host_) return host_;
    // p2p / host-only storage: the host looks straight into simulated memory
    if (staging_.size() != bytes_) staging_.assign(bytes_, 0);
    return staging_.data();
  }
  addr_t device_addr() const override { return addr_; }
  size_t bytes() const override { return bytes_; }
  bufferKind kind() const override { return kind_; }
  void to_device(size_t off, size_t len) override {
    if (len) engine_->mem_write(addr_ + off, static_cast<uint8_t *>(host_ptr()) + off, len);
  }
  void from_device(size_t off, size_t len) override {
    if (len) engine_->mem_read(addr_ + off, static_cast<uint8_t *>(host_ptr()) + off, len);
  }
  bool is_simulated() const override { return true; }

private:
  std::shared_ptr<Engine> engine_;
  size_t bytes_;
  bufferKind kind_;
  uint64_t addr_ = 0;
  uint8_t *host_ = nullptr;
  std::vector<uint8_t> own_, staging_;
};
} // namespace

EmuDevice::EmuDevice(std::shared_ptr<Fabric> fabric, int global_rank, int world, size_t dev_mem, size_t host_mem)
    : fabric_(std::move(fabric)), engine_(std::make_shared<Engine>(global_rank, world, fabric_, dev_mem, host_mem)),
      rank_(global_rank), world_(world) {}

EmuDevice::~EmuDevice() = default;

void EmuDevice::attach(int world_size, int local_rank) {
  if (world_size > world_) throw std::invalid_argument("EmuDevice: communicator larger than the fabric");
  (void)local_rank;
}

ACCLRequest *EmuDevice::start(const Options &options) {
  // chained requests: the engine executes in order, so honouring `waitfor`
  // only needs the dependencies to have been issued; wait for foreign ones
  for (ACCLRequest *dep : options.waitfor)
    if (dep) wait(dep);
  auto req = std::make_shared<BaseRequest>(options);
  req->desc = make_call_desc(options);
  ACCLRequest *h = requests_.add(req);
  EmuCall c;
  c.desc = req->desc;
  c.req = req;
  engine_->submit(std::move(c));
  return h;
}

ACCLRequest *EmuDevice::call(const Options &options) {
  ACCLRequest *h = start(options);
  wait(h);
  return h;
}

void EmuDevice::wait(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("wait: unknown request");
  r->wait();
}
bool EmuDevice::wait(ACCLRequest *request, std::chrono::milliseconds timeout) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("wait: unknown request");
  return r->wait(timeout);
}
bool EmuDevice::test(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("test: unknown request");
  return r->test();
}
void EmuDevice::free_request(ACCLRequest *request) { requests_.erase(request); }
val_t EmuDevice::get_retcode(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("get_retcode: unknown request");
  return r->retcode();
}
uint64_t EmuDevice::get_duration(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("get_duration: unknown request");
  return r->duration_ns();
}

std::string EmuDevice::describe() {
  std::ostringstream o;
  o << "EmuDevice rank " << rank_ << "/" << world_ << " fabric=" << fabric_->name();
  return o.str();
}
void EmuDevice::printDebug() { ACCL_INFO_LOG(engine_->debug_state()); }

std::shared_ptr<BufferStorage> EmuDevice::allocate(size_t bytes, bufferKind kind) {
  return std::make_shared<EmuStorage>(engine_, bytes, kind, nullptr);
}
std::shared_ptr<BufferStorage> EmuDevice::wrap_host(void *host_ptr, size_t bytes) {
  return std::make_shared<EmuStorage>(engine_, bytes, bufferKind::device, host_ptr);
}

std::vector<std::unique_ptr<EmuDevice>> make_inproc_world(int world, size_t dev_mem_bytes) {
  auto fabric = std::make_shared<InProcFabric>(world);
  std::vector<std::unique_ptr<EmuDevice>> v;
  for (int r = 0; r < world; ++r) v.emplace_back(new EmuDevice(fabric, r, world, dev_mem_bytes));
  return v;
}

} // namespace emu
} // namespace accl
