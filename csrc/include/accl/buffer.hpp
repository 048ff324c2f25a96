// Buffers: a host-visible side (byte_array) paired with a device-side
// allocation inside the backend's memory (symmetric NVLink heap on B200,
// simulated device memory in the emulator).  All engine operands are device
// addresses; the host side exists for the "data lives on the host unless told
// otherwise" convention of the API (sync_to_device before / sync_from_device
// after a call).
//
// API parity with the reference hierarchy BaseBuffer / Buffer<T> /
// {XRT,Sim,Coyote,Dummy}Buffer (driver/xrt/include/accl/buffer.hpp:32-203,
// simbuffer.hpp, xrtbuffer.hpp, dummybuffer.hpp).  Differences: one concrete
// typed wrapper `Buffer<T>` over a backend-provided storage object, and a
// device-resident "wrap" mode for torch tensors.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>

#include "accl/constants.hpp"

namespace accl {

enum class bufferKind : uint32_t {
  device = 0,    // host mirror + device allocation (default)
  host_only = 1, // pinned host memory; staged through device scratch by the engine
  p2p = 2,       // device allocation directly visible to peers/host (no mirror copy)
  dummy = 3
};

// Backend-specific storage behind a buffer.
class BufferStorage {
public:
  virtual ~BufferStorage() = default;
  virtual void *host_ptr() = 0;          // may be nullptr for device-only storage
  virtual addr_t device_addr() const = 0; // engine address (heap offset / emulated address)
  virtual void *device_ptr() const { return nullptr; } // real pointer when one exists (CUDA)
  virtual size_t bytes() const = 0;
  virtual bufferKind kind() const = 0;
  // byte-range copies between the two sides
  virtual void to_device(size_t offset, size_t len) = 0;
  virtual void from_device(size_t offset, size_t len) = 0;
  virtual bool is_simulated() const = 0;
};

class BaseBuffer {
public:
  BaseBuffer(std::shared_ptr<BufferStorage> st, size_t byte_offset, size_t size_bytes, dataType type)
      : storage_(std::move(st)), offset_(byte_offset), size_(size_bytes), type_(type) {}
  virtual ~BaseBuffer() = default;

  virtual void sync_from_device() { if (storage_) storage_->from_device(offset_, size_); }
  virtual void sync_to_device() { if (storage_) storage_->to_device(offset_, size_); }
  virtual void free_buffer() { storage_.reset(); }

  size_t size() const { return size_; } // bytes
  dataType type() const { return type_; }
  void *byte_array() const {
    if (!storage_ || !storage_->host_ptr()) return nullptr;
    return static_cast<char *>(storage_->host_ptr()) + offset_;
  }
  // engine-visible address of the first element
  addr_t address() const { return storage_ ? storage_->device_addr() + offset_ : 0; }
  // raw device pointer (CUDA backend), nullptr otherwise
  void *device_ptr() const {
    if (!storage_ || !storage_->device_ptr()) return nullptr;
    return static_cast<char *>(storage_->device_ptr()) + offset_;
  }
  bool is_simulated() const { return storage_ ? storage_->is_simulated() : true; }
  bool is_host_only() const { return storage_ && storage_->kind() == bufferKind::host_only; }
  bool is_dummy() const { return !storage_ || storage_->kind() == bufferKind::dummy; }
  size_t length() const { unsigned b = dtype_bytes(type_); return b ? size_ / b : 0; }

  // element-range view [start, end) sharing the same storage
  std::unique_ptr<BaseBuffer> slice(size_t start, size_t end) const {
    const size_t eb = dtype_bytes(type_);
    if (end < start || end * eb > size_) throw std::out_of_range("BaseBuffer::slice out of range");
    return std::unique_ptr<BaseBuffer>(new BaseBuffer(storage_, offset_ + start * eb, (end - start) * eb, type_));
  }
  const std::shared_ptr<BufferStorage> &storage() const { return storage_; }
  size_t storage_offset() const { return offset_; }

protected:
  std::shared_ptr<BufferStorage> storage_;
  size_t offset_;
  size_t size_;
  dataType type_;
};

template <typename T> class Buffer : public BaseBuffer {
public:
  Buffer(std::shared_ptr<BufferStorage> st, size_t byte_offset, size_t length, dataType type)
      : BaseBuffer(std::move(st), byte_offset, length * sizeof(T), type) {
    if (sizeof(T) != dtype_bytes(type) && type != dataType::none)
      throw std::invalid_argument("Buffer<T>: sizeof(T) does not match dataType");
  }
  T *buffer() const { return static_cast<T *>(byte_array()); }
  size_t length() const { return size_ / sizeof(T); }
  T &operator[](size_t i) { return buffer()[i]; }
  const T &operator[](size_t i) const { return buffer()[i]; }
  std::unique_ptr<Buffer<T>> slice(size_t start, size_t end) const {
    if (end < start || end > length()) throw std::out_of_range("Buffer::slice out of range");
    return std::unique_ptr<Buffer<T>>(new Buffer<T>(storage_, offset_ + start * sizeof(T), end - start, type_));
  }
};

// Placeholder operand for calls that do not use all three address slots
// (reference: dummybuffer.hpp:33-61).
class DummyStorage : public BufferStorage {
public:
  void *host_ptr() override { return nullptr; }
  addr_t device_addr() const override { return 0; }
  size_t bytes() const override { return 0; }
  bufferKind kind() const override { return bufferKind::dummy; }
  void to_device(size_t, size_t) override {}
  void from_device(size_t, size_t) override {}
  bool is_simulated() const override { return false; }
};

class DummyBuffer : public BaseBuffer {
public:
  DummyBuffer() : BaseBuffer(std::make_shared<DummyStorage>(), 0, 0, dataType::none) {}
};

} // namespace accl
