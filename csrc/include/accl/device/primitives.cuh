// sm_100a building blocks for inter-GPU protocols: system-scope
// acquire/release flag accesses on peer-mapped HBM, NVLS multimem
// load-reduce / store / reduce through the NVSwitch multicast address, and
// 16-byte streaming accesses.
//
// These replace the reference's transport hardware: packetizer/depacketizer
// + POE (kernels/cclo/hls/eth_intf/*.cpp) become plain stores into a peer's
// memory followed by a release-flag; the `reduce_ops` arithmetic plugin
// (kernels/plugins/reduce_ops/reduce_ops.cpp:31-107) becomes either the
// in-switch reduction of multimem.ld_reduce or the SM functors in reduce.cuh.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace accl {
namespace dev {

// ------------------------------------------------------------ flag accesses
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_acquire_gpu(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint64_t *p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint64_t *p, uint64_t v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_release_sys_add(uint32_t *p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_sys_add(uint32_t *p, uint32_t v) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acqrel_gpu(uint32_t *p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
// `unsigned long long` flavours (CUDA atomics use that type; uint64_t is `unsigned long` on LP64)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  return ld_acquire_sys(reinterpret_cast<const uint64_t *>(p));
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
  return ld_relaxed_sys(reinterpret_cast<const uint64_t *>(p));
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
  return ld_acquire_gpu(reinterpret_cast<const uint64_t *>(p));
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  st_release_sys(reinterpret_cast<uint64_t *>(p), static_cast<uint64_t>(v));
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  st_relaxed_sys(reinterpret_cast<uint64_t *>(p), static_cast<uint64_t>(v));
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
  st_release_gpu(reinterpret_cast<uint64_t *>(p), static_cast<uint64_t>(v));
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_sc_sys() { asm volatile("fence.sc.sys;" ::: "memory"); }
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void nanosleep(unsigned ns) { asm volatile("nanosleep.u32 %0;" ::"r"(ns)); }

// one signal that lands in the same flag word of every rank (NVLS)
__device__ __forceinline__ void multimem_red_release_add(uint32_t *mc_flag, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_flag), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_red_relaxed_add(uint32_t *mc_flag, uint32_t v) {
  asm volatile("multimem.red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(mc_flag), "r"(v) : "memory");
}

// --------------------------------------------------------- 16-byte data path
struct alignas(16) Vec16 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Vec16 ld_stream(const void *p) {
  Vec16 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
// coherent (peer data may have been written during this kernel)
__device__ __forceinline__ Vec16 ld_volatile16(const void *p) {
  Vec16 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ Vec16 ld_relaxed_sys16(const void *p) {
  Vec16 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_stream(void *p, const Vec16 &v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_relaxed_sys16(void *p, const Vec16 &v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// ------------------------------------------------------------ NVLS multimem
// In-switch reduction: one load returns the element-wise reduction over all
// ranks' copies of the addressed 16 bytes.
enum class McType { f32, f16, bf16 };
enum class McOp { add, max, min };

__device__ __forceinline__ Vec16 multimem_ld_reduce_add_f32(const void *mc) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ Vec16 multimem_ld_reduce_add_bf16(const void *mc) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ Vec16 multimem_ld_reduce_add_f16(const void *mc) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ Vec16 multimem_ld_reduce_max_bf16(const void *mc) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.max.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ Vec16 multimem_ld_reduce_max_f16(const void *mc) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.max.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
// 64-bit integer / double variants go one element at a time
__device__ __forceinline__ double multimem_ld_reduce_add_f64(const void *mc) {
  double v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ int32_t multimem_ld_reduce_add_s32(const void *mc) {
  int32_t v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.s32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ int32_t multimem_ld_reduce_max_s32(const void *mc) {
  int32_t v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ int64_t multimem_ld_reduce_max_s64(const void *mc) {
  int64_t v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s64 %0, [%1];" : "=l"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t multimem_ld_reduce_add_u64(const void *mc) {
  uint64_t v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(mc) : "memory");
  return v;
}

// broadcast store: lands at the same offset of every rank's heap
__device__ __forceinline__ void multimem_st16(void *mc, const Vec16 &v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
// in-switch accumulate into every rank's copy (GEMM->reduce epilogues)
__device__ __forceinline__ void multimem_red_add_f32x4(void *mc, const Vec16 &v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_red_add_bf16x8(void *mc, const Vec16 &v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// accumulate into ONE peer's memory (owner-rank reduce-scatter epilogue)
__device__ __forceinline__ void red_add_f32x4(void *p, const Vec16 &v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_bf16x8(void *p, const Vec16 &v) {
  asm volatile("red.relaxed.sys.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

} // namespace dev
} // namespace accl
