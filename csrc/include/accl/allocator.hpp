// First-fit range allocator with coalescing, used for the emulator's device
// memory and for carving user buffers out of the symmetric NVLink heap.
// (Reference: the per-bank bump allocator of SimBuffer,
// driver/xrt/include/accl/simbuffer.hpp:76-104 — which never frees.)
#pragma once
#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>

namespace accl {

class RangeAllocator {
public:
  RangeAllocator(uint64_t base, uint64_t size) : base_(base), size_(size) { free_[base] = size; }

  uint64_t alloc(uint64_t bytes, uint64_t align = 256) {
    if (bytes == 0) bytes = align;
    std::lock_guard<std::mutex> g(m_);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      const uint64_t start = it->first, len = it->second;
      const uint64_t a = (start + align - 1) / align * align;
      if (a + bytes > start + len) continue;
      free_.erase(it);
      if (a > start) free_[start] = a - start;
      if (a + bytes < start + len) free_[a + bytes] = start + len - (a + bytes);
      used_[a] = bytes;
      in_use_ += bytes;
      return a;
    }
    throw std::bad_alloc();
  }

  void free(uint64_t addr) {
    std::lock_guard<std::mutex> g(m_);
    auto u = used_.find(addr);
    if (u == used_.end()) throw std::invalid_argument("RangeAllocator::free: unknown address");
    uint64_t start = addr, len = u->second;
    in_use_ -= len;
    used_.erase(u);
    auto next = free_.lower_bound(start);
    if (next != free_.end() && start + len == next->first) {
      len += next->second;
      next = free_.erase(next);
    }
    if (next != free_.begin()) {
      auto prev = std::prev(next);
      if (prev->first + prev->second == start) {
        prev->second += len;
        return;
      }
    }
    free_[start] = len;
  }

  uint64_t base() const { return base_; }
  uint64_t size() const { return size_; }
  uint64_t in_use() const { return in_use_; }

private:
  uint64_t base_, size_, in_use_ = 0;
  std::mutex m_;
  std::map<uint64_t, uint64_t> free_; // start -> length
  std::map<uint64_t, uint64_t> used_;
};

} // namespace accl
