// Logging, timing and small helpers shared by every layer.
//
// Reference counterparts: `debug()` under ACCL_DEBUG and the per-rank log
// files (driver/xrt/include/accl/common.hpp:38-42, src/common.cpp:91-136), the
// host `Timer` (driver/xrt/include/accl/timing.hpp:31-99) and the emulator's
// `Log` class (test/log/log.hpp:27-131).  Here one logger serves host API,
// emulator and CUDA backend; level comes from ACCL_LOG_LEVEL (0=error …
// 5=trace; ACCL_DEBUG=1 is shorthand for 4) and ACCL_LOG_FILE=<prefix> sends
// each rank to <prefix><rank>.log.
#pragma once
#include <chrono>
#include <cstdint>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define ACCL_HD __host__ __device__ __forceinline__
#else
#define ACCL_HD inline
#endif

namespace accl {

enum class LogLevel : int { error = 0, warning = 1, info = 2, verbose = 3, debug = 4, trace = 5 };

class Log {
public:
  static Log &get();
  int level() const { return level_; }
  void set_level(int l) { level_ = l; }
  void set_rank(int r) { rank_ = r; }
  // identify the calling process' rank from the launcher environment
  // (RANK, OMPI_COMM_WORLD_RANK, PMI_RANK), as the reference does
  static int rank_from_env();
  void write(LogLevel lvl, const std::string &msg);

private:
  Log();
  int level_ = 1;
  int rank_ = -1;
  std::mutex m_;
  std::ostream *sink_ = nullptr;
};

#define ACCL_LOG(lvl, expr)                                                  \
  do {                                                                       \
    if (static_cast<int>(lvl) <= ::accl::Log::get().level()) {               \
      std::ostringstream _accl_os;                                           \
      _accl_os << expr;                                                      \
      ::accl::Log::get().write(lvl, _accl_os.str());                         \
    }                                                                        \
  } while (0)
#define ACCL_DEBUG_LOG(expr) ACCL_LOG(::accl::LogLevel::debug, expr)
#define ACCL_INFO_LOG(expr) ACCL_LOG(::accl::LogLevel::info, expr)
#define ACCL_WARN_LOG(expr) ACCL_LOG(::accl::LogLevel::warning, expr)
#define ACCL_ERROR_LOG(expr) ACCL_LOG(::accl::LogLevel::error, expr)

// Wall-clock stopwatch (host side); device-side durations come from the
// engine's per-command %globaltimer stamps instead.
class Timer {
public:
  void start() { t0_ = clock::now(); running_ = true; }
  void end() { t1_ = clock::now(); running_ = false; }
  void reset() { t0_ = t1_ = clock::time_point(); running_ = false; }
  // microseconds between start() and end() (or now, if still running)
  uint64_t elapsed() const {
    auto e = running_ ? clock::now() : t1_;
    return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::microseconds>(e - t0_).count());
  }
  uint64_t elapsed_ns() const {
    auto e = running_ ? clock::now() : t1_;
    return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(e - t0_).count());
  }

private:
  using clock = std::chrono::steady_clock;
  clock::time_point t0_, t1_;
  bool running_ = false;
};

// Per-call tracing.  ACCL_TRACE=<prefix> writes <prefix><rank>.json in Chrome / Perfetto
// trace-event format when the process exits (or on flush()): one complete event per call with the host
// issue time as timestamp and the ENGINE-measured duration (perf counter / %globaltimer) as length, plus
// count, communicator, return code and the host-side issue cost.  ACCL_NVTX=1 additionally brackets every
// call issue in an NVTX range (CUDA builds) so that nsys / ncu timelines show the collective names.
class Tracer {
public:
  static Tracer &get();
  static bool enabled() { return enabled_; }
  static bool nvtx() { return nvtx_; }
  // key identifies the request until complete() is called for it
  void issue(const void *key, int rank, const char *op, unsigned count, unsigned comm, uint64_t issue_cost_ns);
  void complete(const void *key, uint32_t retcode, uint64_t device_ns);
  void flush();
  uint64_t now_ns() const;
  void range_push(const char *name);
  void range_pop();

private:
  Tracer();
  ~Tracer();
  struct Ev {
    const char *op;
    int rank;
    unsigned count, comm;
    uint64_t t_issue_ns, issue_cost_ns, device_ns;
    uint32_t retcode;
    bool done;
  };
  static bool enabled_, nvtx_;
  std::mutex m_;
  std::string path_;
  std::vector<Ev> events_;
  std::vector<std::pair<const void *, size_t>> open_; // request key -> index into events_
  std::chrono::steady_clock::time_point t0_;
  int rank_ = 0;
};

// dotted-quad <-> u32, kept for rank tables that carry addresses
uint32_t ip_encode(const std::string &ip);
std::string ip_decode(uint32_t ip);

template <typename T> constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T> constexpr T round_up_to(T a, T b) { return ceil_div(a, b) * b; }

} // namespace accl
