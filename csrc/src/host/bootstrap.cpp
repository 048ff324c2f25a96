#include "accl/bootstrap.hpp"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <mutex>
#include <random>
#include <stdexcept>
#include <thread>

namespace accl {

void Oob::barrier() {
  char c = 0;
  std::vector<char> all(static_cast<size_t>(size()));
  allgather(&c, all.data(), 1);
}

// ---------------------------------------------------------------- LocalOob
struct LocalOob::Shared {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  std::vector<char> stage;
  int arrived = 0, departed = 0;
  uint64_t generation = 0;
};

std::vector<std::shared_ptr<Oob>> LocalOob::create(int world_size) {
  auto s = std::make_shared<Shared>();
  s->n = world_size;
  std::vector<std::shared_ptr<Oob>> v;
  for (int r = 0; r < world_size; ++r)
    v.push_back(std::make_shared<LocalOob>(s, r));
  return v;
}

int LocalOob::size() const { return s_->n; }

void LocalOob::allgather(const void *in, void *out, size_t bytes) {
  std::unique_lock<std::mutex> lk(s_->m);
  // wait until the previous round has fully drained
  s_->cv.wait(lk, [&] { return s_->departed == 0 || s_->arrived < s_->n; });
  if (s_->arrived == 0) s_->stage.assign(bytes * static_cast<size_t>(s_->n), 0);
  if (s_->stage.size() != bytes * static_cast<size_t>(s_->n))
    throw std::runtime_error("LocalOob::allgather: size mismatch across ranks");
  std::memcpy(s_->stage.data() + bytes * static_cast<size_t>(rank_), in, bytes);
  uint64_t gen = s_->generation;
  if (++s_->arrived == s_->n) {
    s_->generation++;
    s_->cv.notify_all();
  } else {
    s_->cv.wait(lk, [&] { return s_->generation != gen; });
  }
  std::memcpy(out, s_->stage.data(), s_->stage.size());
  if (++s_->departed == s_->n) {
    s_->arrived = 0;
    s_->departed = 0;
    s_->cv.notify_all();
  } else {
    // block re-entry of fast ranks until everyone copied out
    s_->cv.wait(lk, [&] { return s_->departed == 0; });
  }
}

// ------------------------------------------------------------------ TcpOob
namespace {
void send_all(int fd, const void *buf, size_t n) {
  const char *p = static_cast<const char *>(buf);
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) throw std::runtime_error("TcpOob: send failed");
    p += k;
    n -= static_cast<size_t>(k);
  }
}
void recv_all(int fd, void *buf, size_t n) {
  char *p = static_cast<char *>(buf);
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) throw std::runtime_error("TcpOob: recv failed (peer gone?)");
    p += k;
    n -= static_cast<size_t>(k);
  }
}
} // namespace

TcpOob::TcpOob(int rank, int size, const std::string &addr, int port,
               int timeout_ms)
    : rank_(rank), size_(size) {
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons(static_cast<uint16_t>(port));
  if (inet_pton(AF_INET, addr.c_str(), &sa.sin_addr) != 1)
    throw std::runtime_error("TcpOob: bad IPv4 address " + addr);
  int one = 1;
  if (rank == 0) {
    listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (::bind(listen_fd_, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) != 0)
      throw std::runtime_error("TcpOob: bind failed on port " + std::to_string(port));
    ::listen(listen_fd_, size);
    peers_.assign(static_cast<size_t>(size), -1);
    for (int i = 1; i < size; ++i) {
      int fd = ::accept(listen_fd_, nullptr, nullptr);
      if (fd < 0) throw std::runtime_error("TcpOob: accept failed");
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      int32_t r = -1;
      recv_all(fd, &r, sizeof(r));
      if (r <= 0 || r >= size || peers_[static_cast<size_t>(r)] != -1)
        throw std::runtime_error("TcpOob: bad hello from peer");
      peers_[static_cast<size_t>(r)] = fd;
    }
  } else {
    auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    int fd = -1;
    for (;;) {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      if (::connect(fd, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) == 0) break;
      ::close(fd);
      if (std::chrono::steady_clock::now() > deadline)
        throw std::runtime_error("TcpOob: timed out connecting to rank 0");
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    int32_t r = rank;
    send_all(fd, &r, sizeof(r));
    peers_.assign(1, fd);
  }
}

TcpOob::~TcpOob() {
  for (int fd : peers_)
    if (fd >= 0) ::close(fd);
  if (listen_fd_ >= 0) ::close(listen_fd_);
}

std::shared_ptr<Oob> TcpOob::from_env(int port_offset) {
  auto env = [](const char *k, const char *d) {
    const char *v = std::getenv(k);
    return std::string(v ? v : d);
  };
  int rank = std::atoi(env("RANK", "0").c_str());
  int size = std::atoi(env("WORLD_SIZE", "1").c_str());
  std::string addr = env("MASTER_ADDR", "127.0.0.1");
  if (addr == "localhost") addr = "127.0.0.1";
  int port = std::getenv("ACCL_PORT")
                 ? std::atoi(std::getenv("ACCL_PORT"))
                 : std::atoi(env("MASTER_PORT", "29500").c_str()) + port_offset;
  return std::make_shared<TcpOob>(rank, size, addr, port);
}

void TcpOob::allgather(const void *in, void *out, size_t bytes) {
  char *o = static_cast<char *>(out);
  if (rank_ == 0) {
    std::memcpy(o, in, bytes);
    for (int r = 1; r < size_; ++r)
      recv_all(peers_[static_cast<size_t>(r)], o + bytes * static_cast<size_t>(r), bytes);
    for (int r = 1; r < size_; ++r)
      send_all(peers_[static_cast<size_t>(r)], o, bytes * static_cast<size_t>(size_));
  } else {
    send_all(peers_[0], in, bytes);
    recv_all(peers_[0], o, bytes * static_cast<size_t>(size_));
  }
}

// ------------------------------------------------------------ fd exchange
namespace {
socklen_t abstract_addr(sockaddr_un &sa, const std::string &name) {
  std::memset(&sa, 0, sizeof(sa));
  sa.sun_family = AF_UNIX;
  // abstract namespace: leading NUL, no filesystem entry to clean up
  size_t n = std::min(name.size(), sizeof(sa.sun_path) - 2);
  std::memcpy(sa.sun_path + 1, name.data(), n);
  return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}
} // namespace

std::vector<int> exchange_fds(Oob &oob, int my_fd, const std::string &channel) {
  const int P = oob.size(), me = oob.rank();
  std::vector<int> out(static_cast<size_t>(P), -1);
  if (oob.same_process()) {
    auto all = oob.allgather_value<int>(my_fd);
    for (int r = 0; r < P; ++r)
      out[static_cast<size_t>(r)] = all[static_cast<size_t>(r)] >= 0 ? ::dup(all[static_cast<size_t>(r)]) : -1;
    oob.barrier(); // nobody closes its fd before every rank has dup'ed it
    return out;
  }
  uint64_t job = 0;
  if (me == 0) {
    std::random_device rd;
    job = (static_cast<uint64_t>(rd()) << 32) ^ rd() ^ static_cast<uint64_t>(::getpid());
  }
  job = oob.bcast_value(job, 0);
  auto has = oob.allgather_value<int>(my_fd >= 0 ? 1 : 0);
  auto name_of = [&](int r) {
    return "accl." + std::to_string(job) + "." + channel + "." + std::to_string(r);
  };
  int lfd = ::socket(AF_UNIX, SOCK_STREAM, 0);
  if (lfd < 0) throw std::runtime_error("exchange_fds: socket() failed");
  sockaddr_un sa;
  socklen_t sl = abstract_addr(sa, name_of(me));
  if (::bind(lfd, reinterpret_cast<sockaddr *>(&sa), sl) != 0)
    throw std::runtime_error("exchange_fds: bind failed");
  ::listen(lfd, P + 1);
  oob.barrier();

  if (my_fd >= 0) {
    out[static_cast<size_t>(me)] = ::dup(my_fd);
    for (int p = 0; p < P; ++p) {
      if (p == me) continue;
      int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
      sockaddr_un pa;
      socklen_t pl = abstract_addr(pa, name_of(p));
      if (::connect(fd, reinterpret_cast<sockaddr *>(&pa), pl) != 0)
        throw std::runtime_error("exchange_fds: connect to peer failed");
      int32_t src = me;
      iovec iov{&src, sizeof(src)};
      alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))] = {};
      msghdr msg{};
      msg.msg_iov = &iov;
      msg.msg_iovlen = 1;
      msg.msg_control = ctrl;
      msg.msg_controllen = sizeof(ctrl);
      cmsghdr *c = CMSG_FIRSTHDR(&msg);
      c->cmsg_level = SOL_SOCKET;
      c->cmsg_type = SCM_RIGHTS;
      c->cmsg_len = CMSG_LEN(sizeof(int));
      std::memcpy(CMSG_DATA(c), &my_fd, sizeof(int));
      if (::sendmsg(fd, &msg, 0) < 0)
        throw std::runtime_error("exchange_fds: sendmsg failed");
      ::close(fd);
    }
  }
  int expect = 0;
  for (int p = 0; p < P; ++p)
    if (p != me && has[static_cast<size_t>(p)]) ++expect;
  for (int i = 0; i < expect; ++i) {
    int fd = ::accept(lfd, nullptr, nullptr);
    if (fd < 0) throw std::runtime_error("exchange_fds: accept failed");
    int32_t src = -1;
    iovec iov{&src, sizeof(src)};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))] = {};
    msghdr msg{};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    if (::recvmsg(fd, &msg, 0) <= 0)
      throw std::runtime_error("exchange_fds: recvmsg failed");
    cmsghdr *c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_type != SCM_RIGHTS || src < 0 || src >= P)
      throw std::runtime_error("exchange_fds: malformed fd message");
    int got = -1;
    std::memcpy(&got, CMSG_DATA(c), sizeof(int));
    out[static_cast<size_t>(src)] = got;
    ::close(fd);
  }
  oob.barrier();
  ::close(lfd);
  return out;
}

// ------------------------------------------------------------ rank tables
std::vector<std::string> get_ips(const std::string &config_file) {
  std::ifstream f(config_file);
  if (!f.good()) throw std::runtime_error("cannot open rank configuration file " + config_file);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string txt = ss.str();
  // minimal JSON: find the "ips" key and take every quoted string up to the closing bracket
  size_t k = txt.find("\"ips\"");
  if (k == std::string::npos) throw std::runtime_error("IPs not specified in config file");
  size_t open = txt.find('[', k), close = txt.find(']', k);
  if (open == std::string::npos || close == std::string::npos || close < open)
    throw std::runtime_error("malformed \"ips\" array in " + config_file);
  std::vector<std::string> ips;
  size_t p = open;
  while (true) {
    size_t a = txt.find('"', p + 1);
    if (a == std::string::npos || a > close) break;
    size_t b = txt.find('"', a + 1);
    if (b == std::string::npos || b > close) throw std::runtime_error("malformed \"ips\" array in " + config_file);
    ips.push_back(txt.substr(a + 1, b - a - 1));
    p = b;
  }
  if (ips.empty()) throw std::runtime_error("IPs not specified in config file");
  return ips;
}

std::vector<std::string> get_ips(bool local, int world_size) {
  std::vector<std::string> ips;
  for (int i = 0; i < world_size; ++i) ips.push_back(local ? "127.0.0.1" : "10.10.10." + std::to_string(i + 1));
  return ips;
}

std::vector<rank_t> generate_ranks(const std::vector<std::string> &ips, int base_port, addr_t rxbuf_size) {
  std::vector<rank_t> ranks;
  for (size_t i = 0; i < ips.size(); ++i)
    ranks.emplace_back(ips[i], base_port + static_cast<int>(i), static_cast<int>(i), rxbuf_size);
  return ranks;
}
std::vector<rank_t> generate_ranks(bool local, int world_size, int base_port, addr_t rxbuf_size) {
  return generate_ranks(get_ips(local, world_size), base_port, rxbuf_size);
}
std::vector<rank_t> generate_ranks(const std::string &config_file, int base_port, addr_t rxbuf_size) {
  return generate_ranks(get_ips(config_file), base_port, rxbuf_size);
}

} // namespace accl
