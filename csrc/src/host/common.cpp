#include "accl/common.hpp"

#ifdef ACCL_WITH_CUDA
#include <nvtx3/nvToolsExt.h>
#endif

#include <arpa/inet.h>

#include <cstdlib>
#include <fstream>
#include <memory>
#include <stdexcept>

namespace accl {

int Log::rank_from_env() {
  for (const char *k : {"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}) {
    const char *v = std::getenv(k);
    if (v && *v) return std::atoi(v);
  }
  return -1;
}

Log::Log() {
  if (const char *v = std::getenv("ACCL_LOG_LEVEL")) level_ = std::atoi(v);
  else if (const char *d = std::getenv("ACCL_DEBUG")) level_ = std::atoi(d) ? 4 : 1;
  rank_ = rank_from_env();
  if (const char *f = std::getenv("ACCL_LOG_FILE")) {
    static std::unique_ptr<std::ofstream> file;
    file.reset(new std::ofstream(std::string(f) + std::to_string(rank_ < 0 ? 0 : rank_) + ".log"));
    if (file->good()) sink_ = file.get();
  }
  if (!sink_) sink_ = &std::cerr;
}

Log &Log::get() {
  static Log l;
  return l;
}

void Log::write(LogLevel lvl, const std::string &msg) {
  static const char *names[] = {"ERROR", "WARN", "INFO", "VERB", "DEBUG", "TRACE"};
  std::lock_guard<std::mutex> g(m_);
  (*sink_) << "[accl";
  if (rank_ >= 0) (*sink_) << " r" << rank_;
  (*sink_) << " " << names[static_cast<int>(lvl)] << "] " << msg << std::endl;
}

// ------------------------------------------------------------------ Tracer
bool Tracer::enabled_ = false;
bool Tracer::nvtx_ = false;

Tracer::Tracer() : t0_(std::chrono::steady_clock::now()) {
  rank_ = Log::rank_from_env();
  if (rank_ < 0) rank_ = 0;
  if (const char *p = std::getenv("ACCL_TRACE")) {
    if (*p) {
      path_ = std::string(p) + std::to_string(rank_) + ".json";
      enabled_ = true;
    }
  }
#ifdef ACCL_WITH_CUDA
  if (const char *n = std::getenv("ACCL_NVTX")) nvtx_ = std::atoi(n) != 0;
#endif
}

Tracer::~Tracer() {
  try {
    flush();
  } catch (...) {
  }
}

Tracer &Tracer::get() {
  static Tracer t;
  return t;
}

uint64_t Tracer::now_ns() const {
  return static_cast<uint64_t>(
      std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0_).count());
}

void Tracer::issue(const void *key, int rank, const char *op, unsigned count, unsigned comm, uint64_t issue_cost_ns) {
  std::lock_guard<std::mutex> g(m_);
  events_.push_back(Ev{op, rank, count, comm, now_ns() - issue_cost_ns, issue_cost_ns, 0, 0, false});
  open_.emplace_back(key, events_.size() - 1);
}

void Tracer::complete(const void *key, uint32_t retcode, uint64_t device_ns) {
  std::lock_guard<std::mutex> g(m_);
  for (size_t i = open_.size(); i-- > 0;)
    if (open_[i].first == key) {
      Ev &e = events_[open_[i].second];
      e.retcode = retcode;
      e.device_ns = device_ns;
      e.done = true;
      open_.erase(open_.begin() + static_cast<long>(i));
      return;
    }
}

void Tracer::flush() {
  std::lock_guard<std::mutex> g(m_);
  if (!enabled_ || path_.empty()) return;
  std::ofstream f(path_);
  if (!f.good()) return;
  f << "{\"displayTimeUnit\": \"ns\", \"traceEvents\": [\n";
  bool first = true;
  for (const Ev &e : events_) {
    if (!first) f << ",\n";
    first = false;
    // complete event: ts / dur in microseconds (fractional allowed)
    f << "{\"name\": \"" << e.op << "\", \"ph\": \"X\", \"pid\": " << e.rank << ", \"tid\": " << e.comm
      << ", \"ts\": " << (static_cast<double>(e.t_issue_ns) / 1e3) << ", \"dur\": "
      << (static_cast<double>(e.done ? e.device_ns : e.issue_cost_ns) / 1e3) << ", \"args\": {\"count\": " << e.count
      << ", \"retcode\": " << e.retcode << ", \"issue_us\": " << (static_cast<double>(e.issue_cost_ns) / 1e3)
      << ", \"engine_ns\": " << e.device_ns << ", \"completed\": " << (e.done ? "true" : "false") << "}}";
  }
  f << "\n]}\n";
}

void Tracer::range_push(const char *name) {
#ifdef ACCL_WITH_CUDA
  if (nvtx_) nvtxRangePushA(name);
#else
  (void)name;
#endif
}
void Tracer::range_pop() {
#ifdef ACCL_WITH_CUDA
  if (nvtx_) nvtxRangePop();
#endif
}

uint32_t ip_encode(const std::string &ip) {
  in_addr a{};
  if (inet_pton(AF_INET, ip.c_str(), &a) != 1) throw std::invalid_argument("bad IPv4 address: " + ip);
  return ntohl(a.s_addr);
}

std::string ip_decode(uint32_t ip) {
  in_addr a{};
  a.s_addr = htonl(ip);
  char buf[INET_ADDRSTRLEN];
  inet_ntop(AF_INET, &a, buf, sizeof(buf));
  return buf;
}

} // namespace accl
