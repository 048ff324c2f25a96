#include "accl/emu/fabric.hpp"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <stdexcept>

#include "accl/common.hpp"

namespace accl {
namespace emu {

// ------------------------------------------------------------ InProcFabric
void InProcFabric::attach(int r, Handler h) {
  std::lock_guard<std::mutex> g(m_);
  handlers_.at(static_cast<size_t>(r)) = std::move(h);
  cv_.notify_all();
}

void InProcFabric::detach(int r) {
  std::lock_guard<std::mutex> g(m_);
  handlers_.at(static_cast<size_t>(r)) = nullptr;
}

void InProcFabric::send(Packet &&p) {
  Handler h;
  {
    std::unique_lock<std::mutex> lk(m_);
    const size_t d = p.hdr.dst;
    if (d >= handlers_.size()) throw std::out_of_range("InProcFabric: bad destination rank");
    // a peer may still be constructing its engine: wait for it briefly
    if (!cv_.wait_for(lk, std::chrono::seconds(30), [&] { return static_cast<bool>(handlers_[d]); }))
      throw std::runtime_error("InProcFabric: destination rank never attached");
    h = handlers_[d];
  }
  h(std::move(p));
}

// ------------------------------------------------------------ SocketFabric
namespace {
bool send_all(int fd, const void *buf, size_t n) {
  const char *p = static_cast<const char *>(buf);
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) return false;
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}
bool recv_all(int fd, void *buf, size_t n) {
  char *p = static_cast<char *>(buf);
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) return false;
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}
} // namespace

SocketFabric::SocketFabric(int my_rank, int world, const std::string &addr, int base_port)
    : me_(my_rank), world_(world), addr_(addr), base_port_(base_port), tx_fd_(static_cast<size_t>(world), -1) {
  for (int i = 0; i < world; ++i) tx_m_.emplace_back(new std::mutex());
  listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons(static_cast<uint16_t>(base_port + my_rank));
  inet_pton(AF_INET, addr.c_str(), &sa.sin_addr);
  if (::bind(listen_fd_, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) != 0)
    throw std::runtime_error("SocketFabric: cannot bind port " + std::to_string(base_port + my_rank));
  ::listen(listen_fd_, world + 4);
  accept_thread_ = std::thread([this] {
    while (!stop_) {
      pollfd pfd{listen_fd_, POLLIN, 0};
      if (::poll(&pfd, 1, 100) <= 0) continue;
      int fd = ::accept(listen_fd_, nullptr, nullptr);
      if (fd < 0) continue;
      int on = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &on, sizeof(on));
      rx_threads_.emplace_back([this, fd] { rx_loop(fd); });
    }
  });
}

SocketFabric::~SocketFabric() {
  stop_ = true;
  for (int &fd : tx_fd_)
    if (fd >= 0) {
      ::shutdown(fd, SHUT_RDWR);
      ::close(fd);
      fd = -1;
    }
  if (accept_thread_.joinable()) accept_thread_.join();
  if (listen_fd_ >= 0) ::close(listen_fd_);
  for (auto &t : rx_threads_)
    if (t.joinable()) t.join();
}

void SocketFabric::attach(int r, Handler h) {
  if (r != me_) throw std::invalid_argument("SocketFabric: a process hosts exactly one rank");
  std::lock_guard<std::mutex> g(hm_);
  handler_ = std::move(h);
  hcv_.notify_all();
}

void SocketFabric::detach(int) {
  std::lock_guard<std::mutex> g(hm_);
  handler_ = nullptr;
}

void SocketFabric::rx_loop(int fd) {
  while (!stop_) {
    pollfd pfd{fd, POLLIN, 0};
    int pr = ::poll(&pfd, 1, 100);
    if (pr == 0) continue;
    if (pr < 0) break;
    Packet p;
    uint64_t len = 0;
    if (!recv_all(fd, &p.hdr, sizeof(p.hdr)) || !recv_all(fd, &len, sizeof(len))) break;
    p.payload.resize(static_cast<size_t>(len));
    if (len && !recv_all(fd, p.payload.data(), static_cast<size_t>(len))) break;
    Handler h;
    {
      std::unique_lock<std::mutex> lk(hm_);
      hcv_.wait_for(lk, std::chrono::seconds(30), [&] { return static_cast<bool>(handler_) || stop_.load(); });
      h = handler_;
    }
    if (h) h(std::move(p));
  }
  ::close(fd);
}

int SocketFabric::connect_to(int peer) {
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons(static_cast<uint16_t>(base_port_ + peer));
  inet_pton(AF_INET, addr_.c_str(), &sa.sin_addr);
  auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(60);
  for (;;) {
    int fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (::connect(fd, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) == 0) {
      int on = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &on, sizeof(on));
      return fd;
    }
    ::close(fd);
    if (std::chrono::steady_clock::now() > deadline)
      throw std::runtime_error("SocketFabric: cannot reach rank " + std::to_string(peer));
    std::this_thread::sleep_for(std::chrono::milliseconds(10));
  }
}

void SocketFabric::send(Packet &&p) {
  const int d = static_cast<int>(p.hdr.dst);
  if (d < 0 || d >= world_) throw std::out_of_range("SocketFabric: bad destination rank");
  if (d == me_) { // loopback without touching the network
    Handler h;
    {
      std::lock_guard<std::mutex> g(hm_);
      h = handler_;
    }
    if (h) h(std::move(p));
    return;
  }
  std::lock_guard<std::mutex> g(*tx_m_[static_cast<size_t>(d)]);
  int &fd = tx_fd_[static_cast<size_t>(d)];
  if (fd < 0) fd = connect_to(d);
  uint64_t len = p.payload.size();
  if (!send_all(fd, &p.hdr, sizeof(p.hdr)) || !send_all(fd, &len, sizeof(len)) ||
      (len && !send_all(fd, p.payload.data(), static_cast<size_t>(len))))
    throw std::runtime_error("SocketFabric: send to rank " + std::to_string(d) + " failed");
}

} // namespace emu
} // namespace accl
