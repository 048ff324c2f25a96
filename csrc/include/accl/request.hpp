// Request bookkeeping shared by the backends: one BaseRequest per started
// call (status, return word, engine-measured duration, blocking wait with
// timeout) and a registry that maps the opaque ACCLRequest ids handed to the
// user back to them.
//
// Same semantics as the reference's BaseRequest / FPGAQueue
// (driver/xrt/include/accl/acclrequest.hpp:39-211), except that more than one
// call may be in flight per device (the engines have real queues).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "accl/cclo.hpp"

namespace accl {

enum class operationStatus : int { QUEUED = 0, EXECUTING = 1, COMPLETED = 2 };

class BaseRequest {
public:
  explicit BaseRequest(const CCLO::Options &o) : options(o) {}
  virtual ~BaseRequest() = default;

  CCLO::Options options;
  CallDesc desc{};

  operationStatus status() const { return status_.load(std::memory_order_acquire); }
  void set_status(operationStatus s) { status_.store(s, std::memory_order_release); }
  val_t retcode() const { return retcode_; }
  uint64_t duration_ns() const { return duration_ns_; }

  // engine side: publish the outcome and wake waiters
  void complete(val_t retcode, uint64_t duration_ns) {
    {
      std::lock_guard<std::mutex> g(m_);
      retcode_ = retcode;
      duration_ns_ = duration_ns;
      status_.store(operationStatus::COMPLETED, std::memory_order_release);
    }
    cv_.notify_all();
  }
  // host side
  virtual void wait() {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return status_.load() == operationStatus::COMPLETED; });
  }
  virtual bool wait(std::chrono::milliseconds timeout) {
    std::unique_lock<std::mutex> lk(m_);
    return cv_.wait_for(lk, timeout, [&] { return status_.load() == operationStatus::COMPLETED; });
  }
  virtual bool test() { return status() == operationStatus::COMPLETED; }

protected:
  std::atomic<operationStatus> status_{operationStatus::QUEUED};
  val_t retcode_ = 0;
  uint64_t duration_ns_ = 0;
  std::mutex m_;
  std::condition_variable cv_;
};

// id <-> request map; ids are what the public API exposes as ACCLRequest*
class RequestRegistry {
public:
  ACCLRequest *add(std::shared_ptr<BaseRequest> r) {
    std::lock_guard<std::mutex> g(m_);
    auto id = std::make_unique<ACCLRequest>(next_++);
    ACCLRequest *key = id.get();
    entries_[key] = Entry{std::move(id), std::move(r)};
    return key;
  }
  std::shared_ptr<BaseRequest> find(ACCLRequest *h) {
    std::lock_guard<std::mutex> g(m_);
    auto it = entries_.find(h);
    return it == entries_.end() ? nullptr : it->second.req;
  }
  void erase(ACCLRequest *h) {
    std::lock_guard<std::mutex> g(m_);
    entries_.erase(h);
  }
  size_t size() {
    std::lock_guard<std::mutex> g(m_);
    return entries_.size();
  }

private:
  struct Entry {
    std::unique_ptr<ACCLRequest> id;
    std::shared_ptr<BaseRequest> req;
  };
  std::mutex m_;
  std::unordered_map<ACCLRequest *, Entry> entries_;
  ACCLRequest next_ = 1;
};

// Blocking FIFO used between API threads and engine threads.
template <typename T> class WorkQueue {
public:
  void push(T v) {
    {
      std::lock_guard<std::mutex> g(m_);
      q_.push_back(std::move(v));
    }
    cv_.notify_one();
  }
  bool pop(T &out, std::chrono::milliseconds timeout) {
    std::unique_lock<std::mutex> lk(m_);
    if (!cv_.wait_for(lk, timeout, [&] { return !q_.empty(); })) return false;
    out = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  bool try_pop(T &out) {
    std::lock_guard<std::mutex> g(m_);
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  size_t size() {
    std::lock_guard<std::mutex> g(m_);
    return q_.size();
  }

private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<T> q_;
};

} // namespace accl
