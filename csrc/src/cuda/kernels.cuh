// sm_100a device code of the collective engine: inter-GPU synchronisation on
// flag pads, eager slot transfers, NVLS (multimem) and peer-to-peer
// collective bodies with the reduction fused into the transfer.  `run_work`
// executes one WorkItem cooperatively on `nctas` CTAs; it is called both by
// the per-call kernel (direct launch) and by the persistent engine kernel.
//
// Reference counterparts: the DMP data loops and RX-buffer matching
// (kernels/cclo/hls/dma_mover/dma_mover.cpp:433-927, rxbuf_offload/*.cpp), the
// arithmetic and cast plugins (kernels/plugins/reduce_ops, hp_compression) and
// the firmware collectives (ccl_offload_control.c:531-2218).  Algorithms are
// NVSwitch-native instead: one-shot through eager slots for small messages,
// two-shot with in-switch reduction (multimem.ld_reduce + multimem.st) or peer
// loads/stores for large ones; no rings or trees — every peer is one hop away.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "accl/cuda/devtypes.hpp"
#include "accl/device/primitives.cuh"

namespace accl {
namespace cuda {
namespace k {

using namespace accl::dev;

constexpr int BLOCK = 512;

__device__ __forceinline__ uint32_t esize_of(uint32_t dtype) { return dtype_bytes(static_cast<dataType>(dtype)); }

// ------------------------------------------------------------------ context
struct Ctx {
  const DevWorld &w;
  const WorkItem &it;
  int cta, nctas;
  Ctrl *me;           // my control block
  uint32_t *err;      // shared-memory error word of this CTA
  uint64_t timeout_ns;
  struct PtrTable *tab; // shared-memory pointer table of this CTA
  __device__ char *heap(uint32_t grank) const { return w.window + static_cast<uint64_t>(grank) * w.heap_bytes; }
  __device__ Ctrl *ctrl(uint32_t grank) const { return reinterpret_cast<Ctrl *>(heap(grank)); }
  __device__ PadBank &pads() const { return me->pad[it.bank]; }
  __device__ PadBank &pads_of(uint32_t grank) const { return ctrl(grank)->pad[it.bank]; }
  __device__ StageBank &stg() const { return me->stg[it.bank]; }
  __device__ StageBank &stg_of(uint32_t grank) const { return ctrl(grank)->stg[it.bank]; }
  __device__ uint32_t kind_word() const { return (it.desc.scenario & 0xFFu) | (it.comm_sig << 8); }
  __device__ uint32_t P() const { return it.comm_size; }
  __device__ uint32_t r() const { return it.comm_rank; }
  __device__ uint32_t g(uint32_t comm_rank) const { return it.members[comm_rank]; }
};

// spin until *p >= target (wrap-safe); false + error bit on timeout
__device__ __forceinline__ bool wait_ge(const uint32_t *p, uint32_t target, const Ctx &c, uint32_t errbit) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (static_cast<int32_t>(ld_acquire_sys(p) - target) < 0) {
    ++spins;
    if (spins > 32) nanosleep(spins > 4096 ? 256 : 32);
    if ((spins & 0x3FF) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > c.timeout_ns) {
        atomicOr(c.err, errbit);
        return false;
      }
    }
  }
  return true;
}

// Symmetric synchronisation of this CTA's channel with the same channel of
// every communicator peer.  Optionally carries (off0, off2) to the peers —
// the rendezvous address exchange — and returns theirs in shared arrays.
// All threads of the CTA must call it.
__device__ __forceinline__ void chan_sync(const Ctx &c, bool exchange, uint64_t my_off0, uint64_t my_off2,
                                          uint64_t *s_off0, uint64_t *s_off2) {
  __syncthreads(); // every thread's prior writes happen-before the release below
  const uint32_t t = threadIdx.x;
  const uint32_t ch = static_cast<uint32_t>(c.cta);
  if (t < c.P()) {
    const uint32_t peer = c.g(t);
    if (t == c.r()) {
      if (exchange) {
        s_off0[t] = my_off0;
        s_off2[t] = my_off2;
      }
    } else {
      PadBank &mine = c.pads();
      PadBank &theirs = c.pads_of(peer);
      const uint32_t v = mine.sent[ch][peer] + 1;
      mine.sent[ch][peer] = v;
      if (exchange) {
        SyncRec *rr = &theirs.rec[ch][c.w.rank];
        st_relaxed_sys(&rr->off0, my_off0);
        st_relaxed_sys(&rr->off2, my_off2);
        st_relaxed_sys(&rr->kind, c.kind_word());
      }
      st_release_sys(&theirs.sig[ch][c.w.rank], v);
      const uint32_t e = mine.expect[ch][peer] + 1;
      mine.expect[ch][peer] = e;
      if (wait_ge(&mine.sig[ch][peer], e, c, RECEIVE_TIMEOUT_ERROR) && exchange) {
        const SyncRec *mr = &mine.rec[ch][peer];
        s_off0[t] = ld_relaxed_sys(&mr->off0);
        s_off2[t] = ld_relaxed_sys(&mr->off2);
        // a peer in a different operation, or in a different communicator that hashes onto this bank
        if (ld_relaxed_sys(&mr->kind) != c.kind_word()) atomicOr(c.err, PACK_SEQ_NUMBER_ERROR);
      }
    }
  }
  __syncthreads();
}

// pairwise variant for send/recv in direct mode: only (me, peer) take part; kinds differ by design
// `post_rec` = false: signal only.  The second meeting of a transfer must NOT rewrite the record: the peer may not have
// read the first one yet (the receiver runs straight from its first into its second meeting).
__device__ __forceinline__ void pair_sync(const Ctx &c, uint32_t peer_comm_rank, uint64_t my_off, uint32_t my_kind,
                                          uint64_t *peer_off, uint32_t *peer_kind, bool post_rec = true) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t ch = static_cast<uint32_t>(c.cta);
    const uint32_t peer = c.g(peer_comm_rank);
    PadBank &mine = c.pads();
    PadBank &theirs = c.pads_of(peer);
    const uint32_t v = mine.sent[ch][peer] + 1;
    mine.sent[ch][peer] = v;
    if (post_rec) {
      SyncRec *rr = &theirs.rec[ch][c.w.rank];
      st_relaxed_sys(&rr->off0, my_off);
      st_relaxed_sys(&rr->kind, my_kind);
    }
    st_release_sys(&theirs.sig[ch][c.w.rank], v);
    const uint32_t e = mine.expect[ch][peer] + 1;
    mine.expect[ch][peer] = e;
    if (wait_ge(&mine.sig[ch][peer], e, c, RECEIVE_TIMEOUT_ERROR)) {
      const SyncRec *mr = &mine.rec[ch][peer];
      *peer_off = ld_relaxed_sys(&mr->off0);
      *peer_kind = ld_relaxed_sys(&mr->kind);
    } else {
      *peer_off = INVALID_OFF;
      *peer_kind = 0xFFFFFFFFu;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------- typed arithmetic
template <typename T> struct Tr;
template <> struct Tr<float> {
  using A = float;
  static __device__ __forceinline__ A up(float x) { return x; }
  static __device__ __forceinline__ float down(A x) { return x; }
};
template <> struct Tr<double> {
  using A = double;
  static __device__ __forceinline__ A up(double x) { return x; }
  static __device__ __forceinline__ double down(A x) { return x; }
};
template <> struct Tr<int32_t> {
  using A = int32_t;
  static __device__ __forceinline__ A up(int32_t x) { return x; }
  static __device__ __forceinline__ int32_t down(A x) { return x; }
};
template <> struct Tr<int64_t> {
  using A = int64_t;
  static __device__ __forceinline__ A up(int64_t x) { return x; }
  static __device__ __forceinline__ int64_t down(A x) { return x; }
};
template <> struct Tr<__half> {
  using A = float;
  static __device__ __forceinline__ A up(__half x) { return __half2float(x); }
  static __device__ __forceinline__ __half down(A x) { return __float2half_rn(x); }
};
template <> struct Tr<__nv_bfloat16> {
  using A = float;
  static __device__ __forceinline__ A up(__nv_bfloat16 x) { return __bfloat162float(x); }
  static __device__ __forceinline__ __nv_bfloat16 down(A x) { return __float2bfloat16_rn(x); }
};
template <> struct Tr<__nv_fp8_e4m3> {
  using A = float;
  static __device__ __forceinline__ A up(__nv_fp8_e4m3 x) { return static_cast<float>(x); }
  static __device__ __forceinline__ __nv_fp8_e4m3 down(A x) { return __nv_fp8_e4m3(x); }
};
template <> struct Tr<__nv_fp8_e5m2> {
  using A = float;
  static __device__ __forceinline__ A up(__nv_fp8_e5m2 x) { return static_cast<float>(x); }
  static __device__ __forceinline__ __nv_fp8_e5m2 down(A x) { return __nv_fp8_e5m2(x); }
};

struct OpSum {
  template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return a + b; }
};
struct OpMax {
  template <typename A> static __device__ __forceinline__ A apply(A a, A b) { return a > b ? a : b; }
};

template <typename T> struct VecOf {
  static constexpr int N = 16 / sizeof(T);
  using A = typename Tr<T>::A;
  static __device__ __forceinline__ void unpack(const Vec16 &v, A (&a)[N]) {
    const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = Tr<T>::up(e[i]);
  }
  static __device__ __forceinline__ Vec16 pack(const A (&a)[N]) {
    Vec16 v;
    T *e = reinterpret_cast<T *>(&v);
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = Tr<T>::down(a[i]);
    return v;
  }
};

// 16-bit floats: widen / narrow two lanes per 32-bit word (bf16 -> f32 is a shift, the narrowing
// is one cvt.rn.*x2 per pair) instead of eight scalar conversions per 16 bytes.  The SM reduce
// paths are ALU-limited for 16-bit types otherwise.
template <> struct VecOf<__nv_bfloat16> {
  static constexpr int N = 8;
  using A = float;
  static __device__ __forceinline__ void unpack(const Vec16 &v, A (&a)[N]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[2 * i] = __uint_as_float(w[i] << 16);
      a[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ Vec16 pack(const A (&a)[N]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 p = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t *>(&p);
    }
    return Vec16{w[0], w[1], w[2], w[3]};
  }
};
template <> struct VecOf<__half> {
  static constexpr int N = 8;
  using A = float;
  static __device__ __forceinline__ void unpack(const Vec16 &v, A (&a)[N]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
      a[2 * i] = f.x;
      a[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ Vec16 pack(const A (&a)[N]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 p = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t *>(&p);
    }
    return Vec16{w[0], w[1], w[2], w[3]};
  }
};

// Pointer table of one data-movement step.  Lives in shared memory so that
// rank-indexed accesses never turn into local-memory arrays.
struct PtrTable {
  const char *src[ACCL_MAX_RANKS];
  char *dst[ACCL_MAX_RANKS];
};

// Core reduce loop: for every element i < n, dst[q][i] = op over p of
// src[p][i], for q < ndst.  Sources and destinations may live on peers.  Work
// is split over (cta, nctas).  NP > 0 fixes the source count at compile time
// (loads of all peers are issued back to back); NP == 0 uses the runtime np.
template <typename T, typename Op, int NP>
__device__ __forceinline__ void reduce_core(const PtrTable &t, int np_rt, int ndst, size_t n, int cta, int nctas) {
  using V = VecOf<T>;
  const int np = NP > 0 ? NP : np_rt;
  bool aligned = true;
  for (int p = 0; p < np; ++p) aligned = aligned && (reinterpret_cast<uintptr_t>(t.src[p]) & 15) == 0;
  for (int q = 0; q < ndst; ++q) aligned = aligned && (reinterpret_cast<uintptr_t>(t.dst[q]) & 15) == 0;
  const size_t stride = static_cast<size_t>(nctas) * blockDim.x;
  const size_t tid = static_cast<size_t>(cta) * blockDim.x + threadIdx.x;
  size_t done = 0;
  if (aligned) {
    const size_t nvec = n / V::N;
    if (NP > 0) {
      const char *sp[NP > 0 ? NP : 1];
#pragma unroll
      for (int p = 0; p < NP; ++p) sp[p] = t.src[p];
      // vectors per thread in flight: enough bytes outstanding to cover the NVLink round trip
      constexpr int U = NP <= 2 ? 4 : (NP <= 4 ? 2 : 1);
      size_t i = tid;
      for (; i + (U - 1) * stride < nvec; i += U * stride) {
        Vec16 in[U][NP > 0 ? NP : 1];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int p = 0; p < NP; ++p) in[u][p] = ld_relaxed_sys16(sp[p] + (i + u * stride) * 16);
        Vec16 out[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          typename V::A acc[V::N], x[V::N];
          V::unpack(in[u][0], acc);
#pragma unroll
          for (int p = 1; p < NP; ++p) {
            V::unpack(in[u][p], x);
#pragma unroll
            for (int e = 0; e < V::N; ++e) acc[e] = Op::apply(acc[e], x[e]);
          }
          out[u] = V::pack(acc);
        }
        for (int q = 0; q < ndst; ++q) {
          char *d = t.dst[q];
#pragma unroll
          for (int u = 0; u < U; ++u) st_relaxed_sys16(d + (i + u * stride) * 16, out[u]);
        }
      }
      for (; i < nvec; i += stride) {
        typename V::A acc[V::N], x[V::N];
        V::unpack(ld_relaxed_sys16(sp[0] + i * 16), acc);
#pragma unroll
        for (int p = 1; p < NP; ++p) {
          V::unpack(ld_relaxed_sys16(sp[p] + i * 16), x);
#pragma unroll
          for (int e = 0; e < V::N; ++e) acc[e] = Op::apply(acc[e], x[e]);
        }
        const Vec16 o = V::pack(acc);
        for (int q = 0; q < ndst; ++q) st_relaxed_sys16(t.dst[q] + i * 16, o);
      }
    } else {
      for (size_t i = tid; i < nvec; i += stride) {
        typename V::A acc[V::N], x[V::N];
        V::unpack(ld_relaxed_sys16(t.src[0] + i * 16), acc);
        for (int p = 1; p < np; ++p) {
          V::unpack(ld_relaxed_sys16(t.src[p] + i * 16), x);
#pragma unroll
          for (int e = 0; e < V::N; ++e) acc[e] = Op::apply(acc[e], x[e]);
        }
        const Vec16 o = V::pack(acc);
        for (int q = 0; q < ndst; ++q) st_relaxed_sys16(t.dst[q] + i * 16, o);
      }
    }
    done = nvec * V::N;
  }
  for (size_t i = done + tid; i < n; i += stride) {
    typename Tr<T>::A acc = Tr<T>::up(reinterpret_cast<const volatile T *>(t.src[0])[i]);
    for (int p = 1; p < np; ++p) acc = Op::apply(acc, Tr<T>::up(reinterpret_cast<const volatile T *>(t.src[p])[i]));
    const T o = Tr<T>::down(acc);
    for (int q = 0; q < ndst; ++q) reinterpret_cast<T *>(t.dst[q])[i] = o;
  }
}

template <typename T, typename Op>
__device__ __forceinline__ void reduce_np(const PtrTable &t, int np, int ndst, size_t n, int cta, int nctas) {
  switch (np) {
  case 2: reduce_core<T, Op, 2>(t, np, ndst, n, cta, nctas); break;
  case 4: reduce_core<T, Op, 4>(t, np, ndst, n, cta, nctas); break;
  case 8: reduce_core<T, Op, 8>(t, np, ndst, n, cta, nctas); break;
  default: reduce_core<T, Op, 0>(t, np, ndst, n, cta, nctas); break;
  }
}

// The one place where (dtype, function) select typed code.  All threads of
// the CTA call it with a table that is already visible (after __syncthreads).
static __device__ __noinline__ void reduce_dispatch(const PtrTable *t, int np, int ndst, size_t n, uint32_t dtype, uint32_t func,
                                             int cta, int nctas, uint32_t *err) {
  const bool sum = func == static_cast<uint32_t>(reduceFunction::SUM);
#define ACCL_RD(TYPE)                                                         \
  if (sum) reduce_np<TYPE, OpSum>(*t, np, ndst, n, cta, nctas);               \
  else reduce_np<TYPE, OpMax>(*t, np, ndst, n, cta, nctas);
  switch (static_cast<dataType>(dtype)) {
  case dataType::float32: ACCL_RD(float); break;
  case dataType::float16: ACCL_RD(__half); break;
  case dataType::bfloat16: ACCL_RD(__nv_bfloat16); break;
  case dataType::float64: ACCL_RD(double); break;
  case dataType::int32: ACCL_RD(int32_t); break;
  case dataType::int64: ACCL_RD(int64_t); break;
  default:
    if (threadIdx.x == 0) atomicOr(err, ARITH_ERROR);
    break;
  }
#undef ACCL_RD
}

// byte copy t->src[0] -> t->dst[0..ndst) (push / broadcast), 16-byte path when aligned
static __device__ __noinline__ void copy_dispatch(const PtrTable *t, int ndst, size_t bytes, int cta, int nctas) {
  const char *src = t->src[0];
  bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int q = 0; q < ndst; ++q) aligned = aligned && (reinterpret_cast<uintptr_t>(t->dst[q]) & 15) == 0;
  const size_t stride = static_cast<size_t>(nctas) * blockDim.x;
  const size_t tid = static_cast<size_t>(cta) * blockDim.x + threadIdx.x;
  size_t done = 0;
  if (aligned) {
    const size_t nvec = bytes / 16;
    size_t i = tid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
      Vec16 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_relaxed_sys16(src + (i + u * stride) * 16);
      for (int q = 0; q < ndst; ++q) {
        char *d = t->dst[q];
#pragma unroll
        for (int u = 0; u < 4; ++u) st_relaxed_sys16(d + (i + u * stride) * 16, v[u]);
      }
    }
    for (; i < nvec; i += stride) {
      const Vec16 v = ld_relaxed_sys16(src + i * 16);
      for (int q = 0; q < ndst; ++q) st_relaxed_sys16(t->dst[q] + i * 16, v);
    }
    done = nvec * 16;
  }
  for (size_t i = done + tid; i < bytes; i += stride) {
    const char b = *reinterpret_cast<const volatile char *>(src + i);
    for (int q = 0; q < ndst; ++q) t->dst[q][i] = b;
  }
}

// single source / single destination copy without a table (whole-CTA helper)
__device__ __forceinline__ void copy_simple(char *dst, const char *src, size_t bytes, int cta, int nctas) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const size_t stride = static_cast<size_t>(nctas) * blockDim.x;
  const size_t tid = static_cast<size_t>(cta) * blockDim.x + threadIdx.x;
  size_t done = 0;
  if (aligned) {
    const size_t nvec = bytes / 16;
    size_t i = tid;
    for (; i + 7 * stride < nvec; i += 8 * stride) {
      Vec16 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ld_relaxed_sys16(src + (i + u * stride) * 16);
#pragma unroll
      for (int u = 0; u < 8; ++u) st_relaxed_sys16(dst + (i + u * stride) * 16, v[u]);
    }
    for (; i < nvec; i += stride) st_relaxed_sys16(dst + i * 16, ld_relaxed_sys16(src + i * 16));
    done = nvec * 16;
  }
  for (size_t i = done + tid; i < bytes; i += stride) dst[i] = *reinterpret_cast<const volatile char *>(src + i);
}

// ------------------------------------------------------------- NVLS bodies
enum class NvOp { add_f32, add_f16, add_bf16, max_f16, max_bf16, none };

__device__ __forceinline__ NvOp nvls_op(uint32_t dtype, uint32_t func) {
  const dataType t = static_cast<dataType>(dtype);
  const bool sum = func == static_cast<uint32_t>(reduceFunction::SUM);
  if (t == dataType::float32) return sum ? NvOp::add_f32 : NvOp::none; // no f32 max in the switch
  if (t == dataType::float16) return sum ? NvOp::add_f16 : NvOp::max_f16;
  if (t == dataType::bfloat16) return sum ? NvOp::add_bf16 : NvOp::max_bf16;
  return NvOp::none;
}

template <NvOp OP> __device__ __forceinline__ Vec16 nv_ld(const void *mc) {
  if (OP == NvOp::add_f32) return multimem_ld_reduce_add_f32(mc);
  if (OP == NvOp::add_f16) return multimem_ld_reduce_add_f16(mc);
  if (OP == NvOp::add_bf16) return multimem_ld_reduce_add_bf16(mc);
  if (OP == NvOp::max_f16) return multimem_ld_reduce_max_f16(mc);
  return multimem_ld_reduce_max_bf16(mc);
}

// out[i] = switch-reduce(in[i]) for vectors [v0, v1); out is a multicast
// address (two-shot allreduce) or a local address (reduce_scatter / reduce).
// U x 16 B per thread are in flight (in-switch reductions have a long round trip).
template <NvOp OP, bool MC_OUT, int U>
__device__ __forceinline__ void nvls_reduce_range(const char *mc_in, char *out, size_t v0, size_t v1, int cta, int nctas) {
  const size_t stride = static_cast<size_t>(nctas) * blockDim.x;
  size_t i = v0 + static_cast<size_t>(cta) * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < v1; i += U * stride) {
    Vec16 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = nv_ld<OP>(mc_in + (i + u * stride) * 16);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (MC_OUT) multimem_st16(out + (i + u * stride) * 16, v[u]);
      else st_stream(out + (i + u * stride) * 16, v[u]);
    }
  }
  for (; i < v1; i += stride) {
    const Vec16 v = nv_ld<OP>(mc_in + i * 16);
    if (MC_OUT) multimem_st16(out + i * 16, v);
    else st_stream(out + i * 16, v);
  }
}

template <NvOp OP, bool MC_OUT>
__device__ __forceinline__ void nvls_reduce_unroll(uint32_t unroll, const char *mc_in, char *out, size_t v0, size_t v1, int cta,
                                                   int nctas) {
  if (unroll == 4) nvls_reduce_range<OP, MC_OUT, 4>(mc_in, out, v0, v1, cta, nctas);
  else if (unroll == 2) nvls_reduce_range<OP, MC_OUT, 2>(mc_in, out, v0, v1, cta, nctas);
  else if (unroll == 16) nvls_reduce_range<OP, MC_OUT, 16>(mc_in, out, v0, v1, cta, nctas);
  else nvls_reduce_range<OP, MC_OUT, 8>(mc_in, out, v0, v1, cta, nctas);
}

template <bool MC_OUT>
__device__ __noinline__ void nvls_reduce_dispatch(NvOp op, uint32_t unroll, const char *mc_in, char *out, size_t v0, size_t v1, int cta,
                                                  int nctas) {
  switch (op) {
  case NvOp::add_f32: nvls_reduce_unroll<NvOp::add_f32, MC_OUT>(unroll, mc_in, out, v0, v1, cta, nctas); break;
  case NvOp::add_f16: nvls_reduce_unroll<NvOp::add_f16, MC_OUT>(unroll, mc_in, out, v0, v1, cta, nctas); break;
  case NvOp::add_bf16: nvls_reduce_unroll<NvOp::add_bf16, MC_OUT>(unroll, mc_in, out, v0, v1, cta, nctas); break;
  case NvOp::max_f16: nvls_reduce_unroll<NvOp::max_f16, MC_OUT>(unroll, mc_in, out, v0, v1, cta, nctas); break;
  case NvOp::max_bf16: nvls_reduce_unroll<NvOp::max_bf16, MC_OUT>(unroll, mc_in, out, v0, v1, cta, nctas); break;
  default: break;
  }
}

// local src -> multicast dst (allgather / bcast through the switch)
__device__ __forceinline__ void nvls_bcast_range(const char *src, char *mc_out, size_t nvec, int cta, int nctas) {
  const size_t stride = static_cast<size_t>(nctas) * blockDim.x;
  size_t i = static_cast<size_t>(cta) * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < nvec; i += 8 * stride) {
    Vec16 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ld_stream(src + (i + u * stride) * 16);
#pragma unroll
    for (int u = 0; u < 8; ++u) multimem_st16(mc_out + (i + u * stride) * 16, v[u]);
  }
  for (; i < nvec; i += stride) multimem_st16(mc_out + i * 16, ld_stream(src + i * 16));
}

// generic element conversion through float/double (cast lanes); slow path
__device__ __forceinline__ double load_as_double(const void *p, uint32_t dtype, size_t i) {
  switch (static_cast<dataType>(dtype)) {
  case dataType::float32: return reinterpret_cast<const float *>(p)[i];
  case dataType::float64: return reinterpret_cast<const double *>(p)[i];
  case dataType::float16: return __half2float(reinterpret_cast<const __half *>(p)[i]);
  case dataType::bfloat16: return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]);
  case dataType::float8_e4m3: return static_cast<float>(reinterpret_cast<const __nv_fp8_e4m3 *>(p)[i]);
  case dataType::float8_e5m2: return static_cast<float>(reinterpret_cast<const __nv_fp8_e5m2 *>(p)[i]);
  case dataType::int32: return reinterpret_cast<const int32_t *>(p)[i];
  case dataType::int64: return static_cast<double>(reinterpret_cast<const int64_t *>(p)[i]);
  default: return 0.0;
  }
}
__device__ __forceinline__ void store_from_double(void *p, uint32_t dtype, size_t i, double v) {
  switch (static_cast<dataType>(dtype)) {
  case dataType::float32: reinterpret_cast<float *>(p)[i] = static_cast<float>(v); break;
  case dataType::float64: reinterpret_cast<double *>(p)[i] = v; break;
  case dataType::float16: reinterpret_cast<__half *>(p)[i] = __float2half_rn(static_cast<float>(v)); break;
  case dataType::bfloat16: reinterpret_cast<__nv_bfloat16 *>(p)[i] = __float2bfloat16_rn(static_cast<float>(v)); break;
  case dataType::float8_e4m3: reinterpret_cast<__nv_fp8_e4m3 *>(p)[i] = __nv_fp8_e4m3(static_cast<float>(v)); break;
  case dataType::float8_e5m2: reinterpret_cast<__nv_fp8_e5m2 *>(p)[i] = __nv_fp8_e5m2(static_cast<float>(v)); break;
  case dataType::int32: reinterpret_cast<int32_t *>(p)[i] = static_cast<int32_t>(v); break;
  case dataType::int64: reinterpret_cast<int64_t *>(p)[i] = static_cast<int64_t>(v); break;
  default: break;
  }
}

} // namespace k
} // namespace cuda
} // namespace accl
