#include "accl/common.hpp"

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
