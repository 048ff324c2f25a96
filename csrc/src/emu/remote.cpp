#include "accl/emu/remote.hpp"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cstring>
#include <sstream>
#include <stdexcept>

#include "accl/common.hpp"

namespace accl {
namespace emu {

using wire::Frame;

namespace {
bool tx_all(int fd, const void *buf, size_t n) {
  const char *p = static_cast<const char *>(buf);
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) return false;
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}
bool rx_all(int fd, void *buf, size_t n) {
  char *p = static_cast<char *>(buf);
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) return false;
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}
sockaddr_in make_addr(const std::string &addr, int port) {
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons(static_cast<uint16_t>(port));
  inet_pton(AF_INET, addr.c_str(), &sa.sin_addr);
  return sa;
}
} // namespace

// ------------------------------------------------------------------ server
EngineServer::EngineServer(std::shared_ptr<Engine> engine, const std::string &addr, int port)
    : engine_(std::move(engine)), conn_(std::make_shared<Conn>()), port_(port) {
  listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in sa = make_addr(addr, port);
  if (::bind(listen_fd_, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) != 0)
    throw std::runtime_error("EngineServer: cannot bind control port " + std::to_string(port));
  ::listen(listen_fd_, 1);
}

EngineServer::~EngineServer() {
  conn_->close();
  if (listen_fd_ >= 0) ::close(listen_fd_);
}

void EngineServer::Conn::send(const Frame &f, const void *payload) {
  std::lock_guard<std::mutex> g(m);
  if (fd < 0) return;
  if (!tx_all(fd, &f, sizeof(f)) || (f.len && !tx_all(fd, payload, f.len))) ACCL_DEBUG_LOG("EngineServer: driver went away");
}

void EngineServer::Conn::close() {
  std::lock_guard<std::mutex> g(m);
  if (fd >= 0) ::close(fd);
  fd = -1;
}

void EngineServer::serve() {
  const int fd = ::accept(listen_fd_, nullptr, nullptr);
  if (fd < 0) throw std::runtime_error("EngineServer: accept failed");
  int on = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &on, sizeof(on));
  {
    std::lock_guard<std::mutex> g(conn_->m);
    conn_->fd = fd;
  }
  std::vector<uint8_t> payload;
  for (;;) {
    Frame f;
    if (!rx_all(fd, &f, sizeof(f))) break; // driver disconnected
    if (f.len > (256u << 20)) {            // not one of ours (the driver chunks at 64 MiB): drop the connection
      ACCL_ERROR_LOG("EngineServer: frame of " << f.len << " bytes refused");
      break;
    }
    payload.resize(f.len);
    if (f.len && !rx_all(fd, payload.data(), f.len)) break;
    Frame r;
    r.type = wire::REPLY;
    r.seq = f.seq;
    std::vector<uint8_t> out;
    bool quit = false;
    try {
      switch (f.type) {
      case wire::MMIO_READ: r.a = engine_->read_exch(static_cast<uint32_t>(f.a)); break;
      case wire::MMIO_WRITE: engine_->write_exch(static_cast<uint32_t>(f.a), static_cast<uint32_t>(f.b)); break;
      case wire::MEM_ALLOC: r.a = engine_->mem_alloc(static_cast<size_t>(f.a), f.b != 0); break;
      case wire::MEM_FREE: engine_->mem_free(f.a); break;
      case wire::MEM_WRITE: engine_->mem_write(f.a, payload.data(), payload.size()); break;
      case wire::MEM_READ:
        out.resize(static_cast<size_t>(f.b));
        engine_->mem_read(f.a, out.data(), out.size());
        break;
      case wire::CALL: {
        if (payload.size() != sizeof(CallDesc)) throw std::runtime_error("CALL: bad descriptor size");
        EmuCall c;
        std::memcpy(&c.desc, payload.data(), sizeof(CallDesc));
        auto req = std::make_shared<BaseRequest>(CCLO::Options{});
        c.req = req;
        const uint64_t id = f.a;
        // runs on the engine's control thread after the request has been completed
        c.on_done = [conn = conn_, id, req](uint32_t rc) {
          Frame e;
          e.type = wire::EVENT_DONE;
          e.a = id;
          e.b = rc;
          const uint64_t dur = req->duration_ns();
          e.len = sizeof(dur);
          conn->send(e, &dur);
        };
        engine_->submit(std::move(c));
        break;
      }
      case wire::KRNL_PUSH: engine_->kernel_push(payload.data(), payload.size()); break;
      case wire::KRNL_PULL:
        out.resize(static_cast<size_t>(f.b));
        r.a = engine_->kernel_pull(static_cast<uint32_t>(f.a), out.data(), out.size(), static_cast<int>(f.aux)) ? 1 : 0;
        if (!r.a) out.clear();
        break;
      case wire::DEBUG_STATE: {
        const std::string s = engine_->debug_state();
        out.assign(s.begin(), s.end());
        break;
      }
      case wire::LOOPBACK: engine_->set_kernel_loopback(f.a != 0); break;
      case wire::SHUTDOWN: quit = true; break;
      default: throw std::runtime_error("unknown request type " + std::to_string(f.type));
      }
    } catch (const std::exception &e) {
      r.aux = 1; // error flag; payload = message
      const std::string s = e.what();
      out.assign(s.begin(), s.end());
    }
    r.len = static_cast<uint32_t>(out.size());
    conn_->send(r, out.data());
    if (quit) break;
  }
  conn_->close();
}

// ------------------------------------------------------------------ client
namespace {
class RemoteStorage : public BufferStorage {
public:
  RemoteStorage(RemoteDevice *d, std::shared_ptr<std::atomic<bool>> alive, size_t bytes, bufferKind kind, void *wrap)
      : dev_(d), alive_(std::move(alive)), bytes_(bytes), kind_(kind) {
    addr_ = dev_->mem_alloc(std::max<size_t>(bytes, 1), kind == bufferKind::host_only);
    if (wrap) host_ = static_cast<uint8_t *>(wrap);
    else {
      own_.assign(bytes, 0);
      host_ = own_.data();
    }
  }
  ~RemoteStorage() override {
    try {
      if (alive_->load()) dev_->mem_free(addr_); // a buffer may outlive its device (interpreter shutdown order)
    } catch (...) {
    }
  }
  void *host_ptr() override { return host_; }
  addr_t device_addr() const override { return addr_; }
  size_t bytes() const override { return bytes_; }
  bufferKind kind() const override { return kind_; }
  void to_device(size_t off, size_t len) override {
    if (len) dev_->mem_write(addr_ + off, host_ + off, len);
  }
  void from_device(size_t off, size_t len) override {
    if (len) dev_->mem_read(addr_ + off, host_ + off, len);
  }
  bool is_simulated() const override { return true; }

private:
  RemoteDevice *dev_;
  std::shared_ptr<std::atomic<bool>> alive_;
  size_t bytes_;
  bufferKind kind_;
  uint64_t addr_ = 0;
  uint8_t *host_ = nullptr;
  std::vector<uint8_t> own_;
};
} // namespace

RemoteDevice::RemoteDevice(const std::string &addr, int port, int global_rank, int world, int connect_timeout_s)
    : rank_(global_rank), world_(world), port_(port), addr_(addr) {
  sockaddr_in sa = make_addr(addr, port);
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(connect_timeout_s);
  for (;;) {
    fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (::connect(fd_, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) == 0) break;
    ::close(fd_);
    fd_ = -1;
    if (std::chrono::steady_clock::now() > deadline)
      throw std::runtime_error("RemoteDevice: no engine process on " + addr + ":" + std::to_string(port));
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
  int on = 1;
  setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &on, sizeof(on));
  reader_ = std::thread([this] { reader_loop(); });
}

RemoteDevice::~RemoteDevice() {
  alive_->store(false);
  stop_ = true;
  if (fd_ >= 0) ::shutdown(fd_, SHUT_RDWR);
  if (reader_.joinable()) reader_.join();
  if (fd_ >= 0) ::close(fd_);
}

void RemoteDevice::reader_loop() {
  std::vector<uint8_t> payload;
  while (!stop_) {
    Frame f;
    if (!rx_all(fd_, &f, sizeof(f))) break;
    payload.resize(f.len);
    if (f.len && !rx_all(fd_, payload.data(), f.len)) break;
    if (f.type == wire::EVENT_DONE) {
      std::shared_ptr<BaseRequest> req;
      {
        std::lock_guard<std::mutex> g(calls_m_);
        auto it = calls_.find(f.a);
        if (it != calls_.end()) {
          req = it->second;
          calls_.erase(it);
        }
      }
      uint64_t dur = 0;
      if (payload.size() >= sizeof(dur)) std::memcpy(&dur, payload.data(), sizeof(dur));
      if (req) req->complete(static_cast<val_t>(f.b), dur);
    } else if (f.type == wire::REPLY) {
      std::lock_guard<std::mutex> g(p_m_);
      Pending &p = pending_[f.seq];
      p.reply = f;
      p.payload = payload;
      p.done = true;
      p_cv_.notify_all();
    }
  }
  // connection lost: fail everything that is still waiting
  broken_ = true;
  {
    std::lock_guard<std::mutex> g(calls_m_);
    for (auto &kv : calls_) kv.second->complete(NOT_READY_ERROR, 0);
    calls_.clear();
  }
  std::lock_guard<std::mutex> g(p_m_);
  p_cv_.notify_all();
}

Frame RemoteDevice::rpc(Frame f, const void *payload, std::vector<uint8_t> *reply_payload) {
  uint32_t seq;
  {
    std::lock_guard<std::mutex> g(p_m_);
    seq = next_seq_++;
    pending_[seq] = Pending{};
  }
  f.seq = seq;
  {
    std::lock_guard<std::mutex> g(tx_m_);
    if (broken_ || !tx_all(fd_, &f, sizeof(f)) || (f.len && !tx_all(fd_, payload, f.len)))
      throw std::runtime_error("RemoteDevice: engine process is gone");
  }
  std::unique_lock<std::mutex> lk(p_m_);
  p_cv_.wait(lk, [&] { return pending_[seq].done || broken_.load(); });
  Pending p = std::move(pending_[seq]);
  pending_.erase(seq);
  lk.unlock();
  if (!p.done) throw std::runtime_error("RemoteDevice: engine process is gone");
  if (p.reply.aux) throw std::runtime_error("engine: " + std::string(p.payload.begin(), p.payload.end()));
  if (reply_payload) *reply_payload = std::move(p.payload);
  return p.reply;
}

val_t RemoteDevice::read(addr_t offset) {
  Frame f;
  f.type = wire::MMIO_READ;
  f.a = offset;
  return static_cast<val_t>(rpc(f, nullptr).a);
}
void RemoteDevice::write(addr_t offset, val_t val) {
  Frame f;
  f.type = wire::MMIO_WRITE;
  f.a = offset;
  f.b = val;
  rpc(f, nullptr);
}
uint64_t RemoteDevice::mem_alloc(size_t bytes, bool host) {
  Frame f;
  f.type = wire::MEM_ALLOC;
  f.a = bytes;
  f.b = host ? 1 : 0;
  return rpc(f, nullptr).a;
}
void RemoteDevice::mem_free(uint64_t addr) {
  Frame f;
  f.type = wire::MEM_FREE;
  f.a = addr;
  rpc(f, nullptr);
}
void RemoteDevice::mem_write(uint64_t addr, const void *src, size_t len) {
  const size_t CH = 64u << 20; // frame length is 32 bits
  for (size_t off = 0; off < len; off += CH) {
    Frame f;
    f.type = wire::MEM_WRITE;
    f.a = addr + off;
    f.len = static_cast<uint32_t>(std::min(CH, len - off));
    rpc(f, static_cast<const uint8_t *>(src) + off);
  }
}
void RemoteDevice::mem_read(uint64_t addr, void *dst, size_t len) {
  const size_t CH = 64u << 20;
  for (size_t off = 0; off < len; off += CH) {
    Frame f;
    f.type = wire::MEM_READ;
    f.a = addr + off;
    f.b = std::min(CH, len - off);
    std::vector<uint8_t> out;
    rpc(f, nullptr, &out);
    if (out.size() != f.b) throw std::runtime_error("RemoteDevice: short memory read");
    std::memcpy(static_cast<uint8_t *>(dst) + off, out.data(), out.size());
  }
}
void RemoteDevice::kernel_push(const void *data, size_t bytes) {
  Frame f;
  f.type = wire::KRNL_PUSH;
  f.len = static_cast<uint32_t>(bytes);
  rpc(f, data);
}
bool RemoteDevice::kernel_pull(uint32_t strm, void *out, size_t bytes, int timeout_ms) {
  Frame f;
  f.type = wire::KRNL_PULL;
  f.a = strm;
  f.b = bytes;
  f.aux = static_cast<uint32_t>(timeout_ms);
  std::vector<uint8_t> data;
  const Frame r = rpc(f, nullptr, &data);
  if (!r.a || data.size() != bytes) return false;
  std::memcpy(out, data.data(), bytes);
  return true;
}
void RemoteDevice::set_kernel_loopback(bool on) {
  Frame f;
  f.type = wire::LOOPBACK;
  f.a = on ? 1 : 0;
  rpc(f, nullptr);
}
std::string RemoteDevice::debug_state() {
  Frame f;
  f.type = wire::DEBUG_STATE;
  std::vector<uint8_t> out;
  rpc(f, nullptr, &out);
  return std::string(out.begin(), out.end());
}
void RemoteDevice::shutdown_engine() {
  Frame f;
  f.type = wire::SHUTDOWN;
  try {
    rpc(f, nullptr);
  } catch (...) {
  }
}

void RemoteDevice::attach(int world_size, int) {
  if (world_size > world_) throw std::invalid_argument("RemoteDevice: communicator larger than the fabric");
}

ACCLRequest *RemoteDevice::start(const Options &options) {
  for (ACCLRequest *dep : options.waitfor)
    if (dep) wait(dep);
  auto req = std::make_shared<BaseRequest>(options);
  req->desc = make_call_desc(options);
  ACCLRequest *h = requests_.add(req);
  uint64_t id;
  {
    std::lock_guard<std::mutex> g(calls_m_);
    id = next_call_++;
    calls_[id] = req;
  }
  Frame f;
  f.type = wire::CALL;
  f.a = id;
  f.len = sizeof(CallDesc);
  rpc(f, &req->desc);
  return h;
}

ACCLRequest *RemoteDevice::call(const Options &options) {
  ACCLRequest *h = start(options);
  wait(h);
  return h;
}

void RemoteDevice::wait(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("wait: unknown request");
  r->wait();
}
bool RemoteDevice::wait(ACCLRequest *request, std::chrono::milliseconds timeout) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("wait: unknown request");
  return r->wait(timeout);
}
bool RemoteDevice::test(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("test: unknown request");
  return r->test();
}
void RemoteDevice::free_request(ACCLRequest *request) { requests_.erase(request); }
val_t RemoteDevice::get_retcode(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("get_retcode: unknown request");
  return r->retcode();
}
uint64_t RemoteDevice::get_duration(ACCLRequest *request) {
  auto r = requests_.find(request);
  if (!r) throw std::invalid_argument("get_duration: unknown request");
  return r->duration_ns();
}

std::string RemoteDevice::describe() {
  std::ostringstream o;
  o << "RemoteDevice rank " << rank_ << "/" << world_ << " engine=" << addr_ << ":" << port_;
  return o.str();
}
void RemoteDevice::printDebug() { ACCL_INFO_LOG(debug_state()); }

std::shared_ptr<BufferStorage> RemoteDevice::allocate(size_t bytes, bufferKind kind) {
  return std::make_shared<RemoteStorage>(this, alive_, bytes, kind, nullptr);
}
std::shared_ptr<BufferStorage> RemoteDevice::wrap_host(void *host_ptr, size_t bytes) {
  return std::make_shared<RemoteStorage>(this, alive_, bytes, bufferKind::device, host_ptr);
}

} // namespace emu
} // namespace accl
