// One-way staged collectives (ALGO_LL / ALGO_STAGED): every rank PUSHES what its peers need into a staging
// region of the peer's heap and consumes what arrived in its own — one NVLink hop, no entry meeting (the staging
// region is always ready: double-buffered by message parity, guarded by an off-critical-path credit), no exit
// meeting (nobody touches a peer's user buffers).
//
//   ALGO_LL      the flag travels inside the data: every 8-byte store carries 4 payload bytes and the 4-byte
//                message sequence number, so the consumer needs no fence and no separate flag — it spins on the
//                data itself (the idea of NCCL's LL protocol).  2x the wire bytes: latency-bound sizes only.
//   ALGO_STAGED  plain 16-byte stores, then fence + one release flag per (channel, peer); the consumer copies /
//                reduces out of staging at HBM speed.
//
// Reference counterpart: the eager protocol — header + payload pushed into the receiver's spare RX buffers and
// matched by (src, seqn) (ccl_offload_control.c:611-648, dma_mover.cpp:579-611, rxbuf_offload/*.cpp), with the
// segmented ring algorithms of the firmware (allreduce :1888-2071, reduce_scatter :1782-1850, allgather :1402-1500)
// collapsed to one hop (all-gather, reduce-scatter, one-shot all-reduce) or two hops (reduce-scatter + all-gather
// all-reduce) because every peer is a neighbour on the NVSwitch.
//
// Message streams are per (bank, channel, ordered pair of ranks): `sent` / `recvd` count messages, message v uses
// staging parity v & 1 and may be written once the destination has acknowledged message v - 2.
#pragma once
#include "kernels.cuh"

namespace accl {
namespace cuda {
namespace k {

// ------------------------------------------------------------ LL primitives
struct alignas(16) LLLine {
  unsigned long long a, b; // each: payload word | flag << 32
};
__device__ __forceinline__ void ll_store(void *p, uint32_t w0, uint32_t w1, uint32_t flag) {
  const unsigned long long a = static_cast<unsigned long long>(w0) | (static_cast<unsigned long long>(flag) << 32);
  const unsigned long long b = static_cast<unsigned long long>(w1) | (static_cast<unsigned long long>(flag) << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ LLLine ll_load(const void *p) {
  LLLine l;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(l.a), "=l"(l.b) : "l"(p) : "memory");
  return l;
}
// spin until both halves of the line carry `flag`; false on timeout
__device__ __forceinline__ bool ll_wait(const void *p, uint32_t flag, uint2 &out, const Ctx &c) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  for (;;) {
    const LLLine l = ll_load(p);
    if (static_cast<uint32_t>(l.a >> 32) == flag && static_cast<uint32_t>(l.b >> 32) == flag) {
      out.x = static_cast<uint32_t>(l.a);
      out.y = static_cast<uint32_t>(l.b);
      return true;
    }
    if (++spins > 64) nanosleep(spins > 8192 ? 256 : 20);
    if ((spins & 0x3FF) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > c.timeout_ns) {
        atomicOr(c.err, RECEIVE_TIMEOUT_ERROR);
        return false;
      }
    }
  }
}

// 8 payload bytes of a local buffer (any alignment, zero padded past `limit`)
__device__ __forceinline__ uint2 ld8_local(const char *base, size_t off, size_t limit) {
  uint2 v{0, 0};
  if ((reinterpret_cast<uintptr_t>(base + off) & 7) == 0 && off + 8 <= limit) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(base + off) : "memory");
  } else {
    unsigned char b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; ++i)
      if (off + i < limit) b[i] = *reinterpret_cast<const volatile unsigned char *>(base + off + i);
    v.x = b[0] | (b[1] << 8) | (b[2] << 16) | (static_cast<uint32_t>(b[3]) << 24);
    v.y = b[4] | (b[5] << 8) | (b[6] << 16) | (static_cast<uint32_t>(b[7]) << 24);
  }
  return v;
}
__device__ __forceinline__ void st8_local(char *base, size_t off, size_t limit, uint2 v) {
  if ((reinterpret_cast<uintptr_t>(base + off) & 7) == 0 && off + 8 <= limit) {
    *reinterpret_cast<uint2 *>(base + off) = v;
  } else {
    const unsigned char b[8] = {static_cast<unsigned char>(v.x), static_cast<unsigned char>(v.x >> 8),
                                static_cast<unsigned char>(v.x >> 16), static_cast<unsigned char>(v.x >> 24),
                                static_cast<unsigned char>(v.y), static_cast<unsigned char>(v.y >> 8),
                                static_cast<unsigned char>(v.y >> 16), static_cast<unsigned char>(v.y >> 24)};
    for (int i = 0; i < 8; ++i)
      if (off + i < limit) base[off + i] = static_cast<char>(b[i]);
  }
}

// element-wise op on 8 bytes of T
template <typename T, typename Op> __device__ __forceinline__ uint2 reduce8(uint2 a, uint2 b) {
  constexpr int N = 8 / sizeof(T);
  using A = typename Tr<T>::A;
  uint2 r;
  const T *ea = reinterpret_cast<const T *>(&a);
  const T *eb = reinterpret_cast<const T *>(&b);
  T *er = reinterpret_cast<T *>(&r);
#pragma unroll
  for (int i = 0; i < N; ++i) er[i] = Tr<T>::down(Op::apply(static_cast<A>(Tr<T>::up(ea[i])), static_cast<A>(Tr<T>::up(eb[i]))));
  return r;
}
__device__ __forceinline__ uint2 reduce8_dyn(uint2 a, uint2 b, uint32_t dtype, bool sum) {
#define ACCL_R8(TYPE) return sum ? reduce8<TYPE, OpSum>(a, b) : reduce8<TYPE, OpMax>(a, b)
  switch (static_cast<dataType>(dtype)) {
  case dataType::float32: ACCL_R8(float);
  case dataType::float16: ACCL_R8(__half);
  case dataType::bfloat16: ACCL_R8(__nv_bfloat16);
  case dataType::float64: ACCL_R8(double);
  case dataType::int32: ACCL_R8(int32_t);
  case dataType::int64: ACCL_R8(int64_t);
  default: return a;
  }
#undef ACCL_R8
}

// ------------------------------------------------------------- geometry
struct StgGeom {
  uint64_t area_off;    // heap offset of the staging area in use (LL lines and plain payloads never share memory:
  uint64_t region;      //   a stale plain payload could otherwise look like an LL line of a later message)
  uint64_t slice_off;   // this CTA's slice inside every (bank, parity, src) region
  uint64_t slice_bytes;
};
// Slices are fixed per channel (not per call): a channel's slice only ever holds messages of that channel's own
// sequence, so a stale LL flag can never equal the sequence number a reader is waiting for.
__device__ __forceinline__ StgGeom stg_geom(const Ctx &c, bool ll) {
  const uint64_t region = ll ? c.w.ll_bytes : c.w.stg_bytes;
  const uint64_t per = (region / static_cast<uint64_t>(STG_CH)) & ~127ull;
  return StgGeom{ll ? c.w.ll_off : c.w.stg_off, region, static_cast<uint64_t>(c.cta) * per, per};
}
// where message `seq` from global rank `src` lands in the heap of global rank `dst` (this CTA's slice)
__device__ __forceinline__ char *stg_ptr(const Ctx &c, uint32_t dst_grank, uint32_t src_grank, uint32_t seq, const StgGeom &g) {
  return c.heap(dst_grank) + g.area_off + ((static_cast<uint64_t>(c.it.bank) * 2 + (seq & 1u)) * c.w.world + src_grank) * g.region +
         g.slice_off;
}

// The exchange patterns (who pushes what to whom) are those of the eager path.
enum StgPattern : uint32_t { SP_ALLREDUCE, SP_REDUCE_SCATTER, SP_ALLGATHER, SP_BCAST, SP_SCATTER, SP_GATHER, SP_REDUCE, SP_ALLTOALL };

struct StgRoles {
  uint32_t pat, P, me, root;
  __device__ bool pushes_to(uint32_t q) const {
    if (q == me) return false;
    switch (pat) {
    case SP_ALLREDUCE: case SP_ALLGATHER: case SP_REDUCE_SCATTER: case SP_ALLTOALL: return true;
    case SP_BCAST: case SP_SCATTER: return me == root;
    case SP_GATHER: case SP_REDUCE: return q == root;
    }
    return false;
  }
  __device__ bool expects_from(uint32_t q) const {
    if (q == me) return false;
    switch (pat) {
    case SP_ALLREDUCE: case SP_ALLGATHER: case SP_REDUCE_SCATTER: case SP_ALLTOALL: return true;
    case SP_BCAST: case SP_SCATTER: return me != root && q == root;
    case SP_GATHER: case SP_REDUCE: return me == root;
    }
    return false;
  }
  // does every destination get the same bytes of my source buffer?
  __device__ bool same_src() const {
    return pat == SP_ALLREDUCE || pat == SP_ALLGATHER || pat == SP_BCAST || pat == SP_GATHER || pat == SP_REDUCE;
  }
};

// credits + sequence numbers of the messages this CTA is about to push: thread q handles peer q.
// s_v[q] = sequence number of my next message to q (0: none)
__device__ __forceinline__ void stg_open(const Ctx &c, const StgRoles &ro, uint32_t *s_v, uint32_t *s_e) {
  const uint32_t t = threadIdx.x, ch = static_cast<uint32_t>(c.cta);
  if (t < ACCL_MAX_RANKS) {
    s_v[t] = 0;
    s_e[t] = 0;
  }
  __syncthreads();
  if (t < ro.P) {
    StageBank &sb = c.stg();
    const uint32_t peer = c.g(t);
    // three independent loads in flight together (each is an L2 round trip on the call's critical path)
    const uint32_t sent = *reinterpret_cast<volatile uint32_t *>(&sb.sent[ch][peer]);
    const uint32_t recvd = *reinterpret_cast<volatile uint32_t *>(&sb.recvd[ch][peer]);
    const uint32_t ack = ld_relaxed_sys(&sb.ack[ch][peer]);
    if (ro.pushes_to(t)) {
      const uint32_t v = sent + 1;
      // the staging region of parity v & 1 was last used by message v - 2: it must have been consumed
      if (v > 2 && static_cast<int32_t>(ack - (v - 2)) < 0) wait_ge(&sb.ack[ch][peer], v - 2, c, DEQUEUE_BUFFER_TIMEOUT_ERROR);
      s_v[t] = v;
      sb.sent[ch][peer] = v;
    }
    if (ro.expects_from(t)) {
      const uint32_t e = recvd + 1;
      s_e[t] = e;
      sb.recvd[ch][peer] = e;
    }
  }
  __syncthreads();
}

// hand the credits back: all threads are done reading the staging regions of this exchange
__device__ __forceinline__ void stg_close(const Ctx &c, const uint32_t *s_e) {
  __syncthreads();
  const uint32_t t = threadIdx.x, ch = static_cast<uint32_t>(c.cta);
  if (t < c.P() && s_e[t]) st_relaxed_sys(&c.stg_of(c.g(t)).ack[ch][c.w.rank], s_e[t]);
}

// ALGO_STAGED: raise the arrival flags after the payload stores of the whole CTA
__device__ __forceinline__ void stg_signal(const Ctx &c, const uint32_t *s_v) {
  __syncthreads();
  const uint32_t t = threadIdx.x, ch = static_cast<uint32_t>(c.cta);
  if (t < c.P() && s_v[t]) {
    fence_acq_rel_sys(); // payload was written by other threads of this CTA (ordered by the barrier above)
    st_release_sys(&c.stg_of(c.g(t)).sig[ch][c.w.rank], s_v[t]);
  }
}
// ALGO_STAGED: wait for the arrival flags of the expected messages
__device__ __forceinline__ void stg_await(const Ctx &c, const uint32_t *s_e) {
  const uint32_t t = threadIdx.x, ch = static_cast<uint32_t>(c.cta);
  if (t < c.P() && s_e[t]) wait_ge(&c.stg().sig[ch][c.g(t)], s_e[t], c, RECEIVE_TIMEOUT_ERROR);
  __syncthreads();
}

// byte range of block `count * es` owned by this CTA, in units of `unit` bytes
struct Part {
  size_t off, bytes;
};
__device__ __forceinline__ Part stg_part(const Ctx &c, size_t total_bytes, size_t unit) {
  const size_t units = (total_bytes + unit - 1) / unit;
  const size_t per = (units + c.nctas - 1) / c.nctas;
  const size_t u0 = static_cast<size_t>(c.cta) * per < units ? static_cast<size_t>(c.cta) * per : units;
  const size_t u1 = u0 + per < units ? u0 + per : units;
  Part p;
  p.off = u0 * unit;
  p.bytes = (u1 * unit < total_bytes ? u1 * unit : total_bytes) - p.off;
  if (u0 >= units) p.bytes = 0;
  return p;
}

// ---------------------------------------------------------------- LL bodies
// One generic LL exchange: `blk` bytes per block; my part of every block is [part.off, part.off + part.bytes).
//  push:    for every destination q: lines of (same_src ? src : src + q * blk)
//  consume: REDUCE patterns fold the lines of all sources in communicator-rank order (bit-identical on every rank);
//           placement patterns store the lines of source q at dst + q * blk (or dst for bcast / scatter)
// `fwd`: two-shot all-reduce, first half: the reduced line is stored to dst + me_blk_off AND pushed to every peer as the
// next message (second half of the all-reduce) without leaving the registers.
__device__ __noinline__ void ll_exchange(const Ctx &c, StgPattern pat, const char *src, char *dst, size_t blk, bool fwd,
                                         uint32_t *s_v, uint32_t *s_e, uint32_t *s_v2) {
  const WorkItem &it = c.it;
  const StgRoles ro{pat, c.P(), c.r(), it.desc.root_src_dst};
  const StgGeom g = stg_geom(c, true);
  const Part part = stg_part(c, blk, 8);
  const size_t lines = (part.bytes + 7) / 8;
  const uint32_t P = ro.P, me = ro.me;
  const bool sum = it.desc.function == static_cast<uint32_t>(reduceFunction::SUM);
  if (lines * 16 > g.slice_bytes) { // the planner sizes messages to fit; never expected
    if (threadIdx.x == 0) atomicOr(c.err, DMA_SIZE_ERROR);
    return;
  }
  // the first line of a same-source push does not depend on the counters: its load overlaps their round trip
  const bool same = ro.same_src();
  uint2 first{0, 0};
  if (same && threadIdx.x < lines) first = ld8_local(src, part.off + static_cast<size_t>(threadIdx.x) * 8, blk);
  stg_open(c, ro, s_v, s_e);
  // ---- push
  for (size_t i = threadIdx.x; i < lines; i += blockDim.x) {
    uint2 v = same ? (i == threadIdx.x ? first : ld8_local(src, part.off + i * 8, blk)) : uint2{0, 0};
    for (uint32_t k = 1; k < P; ++k) {
      const uint32_t q = (me + k) % P; // stagger destinations across ranks
      const uint32_t sv = s_v[q];
      if (!sv) continue;
      if (!same) v = ld8_local(src + static_cast<size_t>(q) * blk, part.off + i * 8, blk);
      ll_store(stg_ptr(c, c.g(q), c.w.rank, sv, g) + i * 16, v.x, v.y, sv);
    }
  }
  // ---- second-half credits of a forwarding exchange are taken now, before anything is consumed
  if (fwd) {
    const StgRoles ro2{SP_ALLGATHER, P, me, 0};
    uint32_t *s_e2 = s_v2 + ACCL_MAX_RANKS;
    stg_open(c, ro2, s_v2, s_e2);
  }
  // ---- consume
  const bool reducing = pat == SP_ALLREDUCE || pat == SP_REDUCE_SCATTER || pat == SP_REDUCE;
  const bool consumer = !(pat == SP_GATHER || pat == SP_REDUCE) || me == ro.root;
  if (consumer) {
    for (size_t i = threadIdx.x; i < lines; i += blockDim.x) {
      const size_t o = part.off + i * 8;
      if (reducing) {
        uint2 acc{0, 0};
        bool ok = true;
        for (uint32_t q = 0; q < P && ok; ++q) {
          uint2 x;
          if (q == me) {
            x = ld8_local(pat == SP_REDUCE_SCATTER ? src + static_cast<size_t>(me) * blk : src, o, blk);
          } else {
            ok = ll_wait(stg_ptr(c, c.w.rank, c.g(q), s_e[q], g) + i * 16, s_e[q], x, c);
          }
          acc = q == 0 ? x : reduce8_dyn(acc, x, it.udtype, sum);
        }
        if (!ok) break;
        if (fwd) {
          // my reduced shard: keep it and forward it to everybody as the all-gather half
          st8_local(dst + static_cast<size_t>(me) * blk, o, blk, acc);
          for (uint32_t k = 1; k < P; ++k) {
            const uint32_t q = (me + k) % P;
            ll_store(stg_ptr(c, c.g(q), c.w.rank, s_v2[q], g) + i * 16, acc.x, acc.y, s_v2[q]);
          }
        } else {
          st8_local(dst, o, blk, acc);
        }
      } else {
        // placement: my own block first (local copy), then whatever arrives
        if (pat == SP_ALLGATHER || pat == SP_GATHER)
          st8_local(dst + static_cast<size_t>(me) * blk, o, blk, ld8_local(src, o, blk));
        else if (pat == SP_ALLTOALL)
          st8_local(dst + static_cast<size_t>(me) * blk, o, blk, ld8_local(src + static_cast<size_t>(me) * blk, o, blk));
        else if (pat == SP_SCATTER && me == ro.root)
          st8_local(dst, o, blk, ld8_local(src + static_cast<size_t>(me) * blk, o, blk));
        for (uint32_t k = 1; k < P; ++k) {
          const uint32_t q = (me + P - k) % P; // the peer that staggered its stores towards me first
          if (!s_e[q]) continue;
          uint2 x;
          if (!ll_wait(stg_ptr(c, c.w.rank, c.g(q), s_e[q], g) + i * 16, s_e[q], x, c)) break;
          char *to = (pat == SP_BCAST || pat == SP_SCATTER) ? dst : dst + static_cast<size_t>(q) * blk;
          st8_local(to, o, blk, x);
        }
      }
    }
  }
  stg_close(c, s_e);
  if (fwd) {
    // ---- second half: the other ranks' reduced shards arrive as a new message each
    uint32_t *s_e2 = s_v2 + ACCL_MAX_RANKS;
    for (size_t i = threadIdx.x; i < lines; i += blockDim.x) {
      const size_t o = part.off + i * 8;
      for (uint32_t k = 1; k < P; ++k) {
        const uint32_t q = (me + P - k) % P;
        uint2 x;
        if (!ll_wait(stg_ptr(c, c.w.rank, c.g(q), s_e2[q], g) + i * 16, s_e2[q], x, c)) break;
        st8_local(dst + static_cast<size_t>(q) * blk, o, blk, x);
      }
    }
    stg_close(c, s_e2);
  }
  __syncthreads();
}

// -------------------------------------------------------------- STAGED bodies
// push my part of one block to the staging slice of every destination (16-byte stores, one read fans out)
__device__ __forceinline__ void stg_push(const Ctx &c, const StgRoles &ro, const char *src, size_t blk, const Part &part,
                                         const StgGeom &g, const uint32_t *s_v, uint32_t *s_tmp) {
  const uint32_t P = ro.P, me = ro.me;
  if (ro.same_src()) {
    __syncthreads();
    if (threadIdx.x == 0) {
      int nd = 0;
      for (uint32_t k = 1; k < P; ++k) {
        const uint32_t q = (me + k) % P;
        if (s_v[q]) c.tab->dst[nd++] = stg_ptr(c, c.g(q), c.w.rank, s_v[q], g);
      }
      c.tab->src[0] = src + part.off;
      *s_tmp = static_cast<uint32_t>(nd);
    }
    __syncthreads();
    const int nd = static_cast<int>(*s_tmp);
    if (nd && part.bytes) copy_dispatch(c.tab, nd, part.bytes, 0, 1);
  } else {
    for (uint32_t k = 1; k < P; ++k) {
      const uint32_t q = (me + k) % P;
      if (!s_v[q] || !part.bytes) continue;
      copy_simple(stg_ptr(c, c.g(q), c.w.rank, s_v[q], g), src + static_cast<size_t>(q) * blk + part.off, part.bytes, 0, 1);
    }
  }
}

// reduce my local part with the staged parts of all expected sources into `out` (communicator-rank order)
__device__ __forceinline__ void stg_reduce(const Ctx &c, const char *local, char *out, const Part &part, const StgGeom &g,
                                           const uint32_t *s_e) {
  const uint32_t P = c.P();
  __syncthreads();
  if (threadIdx.x < P) {
    const uint32_t q = threadIdx.x;
    c.tab->src[q] = q == c.r() ? local + part.off : stg_ptr(c, c.w.rank, c.g(q), s_e[q], g);
  }
  if (threadIdx.x == 0) c.tab->dst[0] = out + part.off;
  __syncthreads();
  if (part.bytes)
    reduce_dispatch(c.tab, static_cast<int>(P), 1, part.bytes / esize_of(c.it.udtype), c.it.udtype, c.it.desc.function, 0, 1, c.err);
}

__device__ __noinline__ void staged_exchange(const Ctx &c, StgPattern pat, const char *src, char *dst, size_t blk, uint32_t *s_v,
                                             uint32_t *s_e, uint32_t *s_tmp) {
  const WorkItem &it = c.it;
  const StgRoles ro{pat, c.P(), c.r(), it.desc.root_src_dst};
  const StgGeom g = stg_geom(c, false);
  const Part part = stg_part(c, blk, 16);
  const uint32_t P = ro.P, me = ro.me;
  if (part.bytes > g.slice_bytes) {
    if (threadIdx.x == 0) atomicOr(c.err, DMA_SIZE_ERROR);
    return;
  }
  stg_open(c, ro, s_v, s_e);
  stg_push(c, ro, src, blk, part, g, s_v, s_tmp);
  stg_signal(c, s_v);
  // my own contribution to placement patterns does not depend on anybody
  if (part.bytes) {
    if ((pat == SP_ALLGATHER || (pat == SP_GATHER && me == ro.root)) && dst + static_cast<size_t>(me) * blk != src)
      copy_simple(dst + static_cast<size_t>(me) * blk + part.off, src + part.off, part.bytes, 0, 1);
    else if (pat == SP_ALLTOALL) copy_simple(dst + static_cast<size_t>(me) * blk + part.off, src + static_cast<size_t>(me) * blk + part.off, part.bytes, 0, 1);
    else if (pat == SP_SCATTER && me == ro.root) copy_simple(dst + part.off, src + static_cast<size_t>(me) * blk + part.off, part.bytes, 0, 1);
  }
  stg_await(c, s_e);
  if (*c.err == 0) {
    switch (pat) {
    case SP_ALLREDUCE: stg_reduce(c, src, dst, part, g, s_e); break;
    case SP_REDUCE_SCATTER: stg_reduce(c, src + static_cast<size_t>(me) * blk, dst, part, g, s_e); break;
    case SP_REDUCE:
      if (me == ro.root) stg_reduce(c, src, dst, part, g, s_e);
      break;
    default:
      for (uint32_t k = 1; k < P; ++k) {
        const uint32_t q = (me + P - k) % P;
        if (!s_e[q] || !part.bytes) continue;
        char *to = (pat == SP_BCAST || pat == SP_SCATTER) ? dst + part.off : dst + static_cast<size_t>(q) * blk + part.off;
        copy_simple(to, stg_ptr(c, c.w.rank, c.g(q), s_e[q], g), part.bytes, 0, 1);
      }
      break;
    }
  }
  stg_close(c, s_e);
  __syncthreads();
}

// ------------------------------------------------------------------ entry
// shard geometry of the two-shot all-reduce: P blocks of `blk` bytes (16-byte multiple) cover count * es
__device__ __forceinline__ size_t ar_shard_bytes(size_t total, uint32_t P) {
  const size_t v = (total + 15) / 16;
  return ((v + P - 1) / P) * 16;
}

__device__ __noinline__ void stg_collective(const Ctx &c, StgPattern pat, uint32_t *s_tmp) {
  const WorkItem &it = c.it;
  __shared__ uint32_t s_v[ACCL_MAX_RANKS], s_e[ACCL_MAX_RANKS], s_v2[2 * ACCL_MAX_RANKS];
  const bool ll = it.algo == ALGO_LL;
  const uint32_t P = c.P(), me = c.r();
  const size_t es = esize_of(it.udtype);
  const size_t blk = static_cast<size_t>(it.desc.count) * es;
  const char *src = c.heap(c.w.rank) + it.desc.addr0();
  char *dst = c.heap(c.w.rank) + (pat == SP_BCAST ? it.desc.addr0() : it.desc.addr2());
  if (pat == SP_ALLREDUCE && !(it.flags & WF_ONESHOT)) {
    // two-shot: reduce-scatter over shards, then all-gather of the reduced shards.  The planner guarantees
    // count * es to be a multiple of 16 * P here, so shards are whole and equal.
    const size_t sh = blk / P;
    if (ll) {
      // src block q goes to rank q; reduced shard is forwarded from registers
      ll_exchange(c, SP_REDUCE_SCATTER, src, dst, sh, true, s_v, s_e, s_v2);
    } else {
      staged_exchange(c, SP_REDUCE_SCATTER, src, dst + static_cast<size_t>(me) * sh, sh, s_v, s_e, s_tmp);
      __syncthreads();
      staged_exchange(c, SP_ALLGATHER, dst + static_cast<size_t>(me) * sh, dst, sh, s_v, s_e, s_tmp);
    }
    return;
  }
  if (ll) ll_exchange(c, pat, src, dst, blk, false, s_v, s_e, s_v2);
  else staged_exchange(c, pat, src, dst, blk, s_v, s_e, s_tmp);
}

} // namespace k
} // namespace cuda
} // namespace accl
