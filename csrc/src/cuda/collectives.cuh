// Collective and primitive bodies executed by `run_work` (see kernels.cuh).
// Two protocol families, selected per call exactly like the reference
// firmware does (eager iff small, or compressed; ccl_offload_control.c:587,
// 808,1878):
//   * eager: data is pushed into the destination's slot ring (the RX
//     buffers), flagged with a per-(channel, src) sequence counter, consumed
//     and acknowledged (credit return).  One network hop, works for every
//     dtype / wire-dtype combination, segmented by slot size.
//   * rendezvous: peers exchange buffer offsets on the sync pads (the
//     address hand-shake), then move data straight between user buffers with
//     multimem (in-switch reduce/broadcast) or peer loads/stores, and meet
//     again on the pads.
#pragma once
#include "kernels.cuh"

namespace accl {
namespace cuda {
namespace k {

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ uint32_t esize(uint32_t dtype) { return esize_of(dtype); }
__device__ __forceinline__ bool is_fp8_dt(uint32_t dtype) {
  return dtype == static_cast<uint32_t>(dataType::float8_e4m3) || dtype == static_cast<uint32_t>(dataType::float8_e5m2);
}
// bytes of n elements in wire representation (block-scaled fp8 appends fp32 scales, 4-byte aligned)
__device__ __forceinline__ uint64_t wire_bytes(uint32_t wdt, uint32_t rl, uint64_t n) {
  uint64_t b = n * esize(wdt);
  if (is_fp8_dt(wdt) && rl) b = ((b + 3) & ~3ull) + 4 * ((n + (1u << rl) - 1) >> rl);
  return b;
}

// Convert n elements src(src_t) -> dst(dst_t) with the whole CTA.  Identical
// types take the vector copy path.  A block-scaled fp8 destination gets one
// scale per 2^rl elements (computed by the 32 lanes of a warp when rl == 5).
__device__ __forceinline__ void cast_copy(void *dst, uint32_t dst_t, const void *src, uint32_t src_t, size_t n, uint32_t rl) {
  if (dst_t == src_t && !(is_fp8_dt(dst_t) && rl)) {
    copy_simple(static_cast<char *>(dst), static_cast<const char *>(src), n * esize(dst_t), 0, 1);
    return;
  }
  const bool dst_scaled = is_fp8_dt(dst_t) && rl, src_scaled = is_fp8_dt(src_t) && rl;
  const float *src_scales = reinterpret_cast<const float *>(static_cast<const char *>(src) + ((n * esize(src_t) + 3) & ~3ull));
  float *dst_scales = reinterpret_cast<float *>(static_cast<char *>(dst) + ((n * esize(dst_t) + 3) & ~3ull));
  if (dst_scaled && rl == 5) {
    const float fmax = dst_t == static_cast<uint32_t>(dataType::float8_e4m3) ? 448.f : 57344.f;
    const size_t nblk = (n + 31) / 32;
    for (size_t b = threadIdx.x / 32; b < nblk; b += blockDim.x / 32) {
      const size_t i = b * 32 + (threadIdx.x & 31);
      float x = 0.f;
      if (i < n) x = static_cast<float>(load_as_double(src, src_t, i)) * (src_scaled ? src_scales[i >> rl] : 1.f);
      float m = fabsf(x);
#pragma unroll
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float scale = m > 0.f ? m / fmax : 1.f;
      if (i < n) store_from_double(dst, dst_t, i, x / scale);
      if ((threadIdx.x & 31) == 0) dst_scales[b] = scale;
    }
    return;
  }
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    double x = load_as_double(src, src_t, i);
    if (src_scaled) x *= src_scales[i >> rl];
    store_from_double(dst, dst_t, i, x);
  }
}

__device__ __forceinline__ double wire_load(const void *p, uint32_t wdt, uint32_t rl, size_t n, size_t i) {
  double x = load_as_double(p, wdt, i);
  if (is_fp8_dt(wdt) && rl)
    x *= reinterpret_cast<const float *>(static_cast<const char *>(p) + ((n * esize(wdt) + 3) & ~3ull))[i >> rl];
  return x;
}

// ------------------------------------------------------------ local calls
__device__ __noinline__ void op_copy(const Ctx &c) {
  const WorkItem &it = c.it;
  const uint32_t cf = it.desc.compression_flags;
  const uint32_t st = (cf & 1) ? it.cdtype : it.udtype, dt = (cf & 4) ? it.cdtype : it.udtype;
  const char *src = c.heap(c.w.rank) + it.desc.addr0();
  char *dst = c.heap(c.w.rank) + it.desc.addr2();
  const size_t n = it.desc.count;
  if (st == dt) {
    copy_simple(dst, src, n * esize(st), c.cta, c.nctas);
  } else {
    const size_t stride = static_cast<size_t>(c.nctas) * blockDim.x;
    for (size_t i = static_cast<size_t>(c.cta) * blockDim.x + threadIdx.x; i < n; i += stride)
      store_from_double(dst, dt, i, load_as_double(src, st, i));
  }
}

__device__ __noinline__ void op_combine(const Ctx &c) {
  const WorkItem &it = c.it;
  const uint32_t cf = it.desc.compression_flags;
  const uint32_t t0 = (cf & 1) ? it.cdtype : it.udtype, t1 = (cf & 2) ? it.cdtype : it.udtype,
                 tr = (cf & 4) ? it.cdtype : it.udtype;
  const char *a = c.heap(c.w.rank) + it.desc.addr0();
  const char *b = c.heap(c.w.rank) + it.desc.addr1();
  char *d = c.heap(c.w.rank) + it.desc.addr2();
  const size_t n = it.desc.count;
  if (t0 == t1 && t1 == tr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      c.tab->src[0] = a;
      c.tab->src[1] = b;
      c.tab->dst[0] = d;
    }
    __syncthreads();
    reduce_dispatch(c.tab, 2, 1, n, t0, it.desc.function, c.cta, c.nctas, c.err);
    return;
  }
  const uint32_t at = it.arith_compressed ? it.cdtype : it.udtype;
  const bool sum = it.desc.function == static_cast<uint32_t>(reduceFunction::SUM);
  const size_t stride = static_cast<size_t>(c.nctas) * blockDim.x;
  for (size_t i = static_cast<size_t>(c.cta) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double x = load_as_double(a, t0, i), y = load_as_double(b, t1, i);
    double r = sum ? x + y : (x > y ? x : y);
    if (at != static_cast<uint32_t>(dataType::float64)) r = static_cast<float>(r);
    store_from_double(d, tr, i, r);
  }
}

// ------------------------------------------------------------ eager slots
struct EgrSeg { // one segment of one block, as seen by this CTA
  size_t elem_off; // offset inside the block
  uint32_t elems;
};

__device__ __forceinline__ void egr_push(const Ctx &c, uint32_t peer_cr, const void *src, uint32_t src_t, uint32_t wire_t,
                                          uint32_t elems, uint32_t tag, uint32_t *s_tmp) {
  const uint32_t ch = static_cast<uint32_t>(c.cta);
  const uint32_t peer = c.g(peer_cr);
  __syncthreads(); // s_tmp may still be read by stragglers of the previous step
  if (threadIdx.x == 0) {
    const uint32_t v = c.me->egr_sent[ch][peer] + 1;
    // credit: the slot I am about to overwrite (message v - depth) must have been consumed
    if (v > c.w.egr_depth) wait_ge(&c.me->egr_ack[ch][peer], v - c.w.egr_depth, c, DEQUEUE_BUFFER_TIMEOUT_ERROR);
    *s_tmp = v;
  }
  __syncthreads();
  const uint32_t v = *s_tmp;
  char *slot = c.heap(peer) + egr_slot_off(c.w, ch, v % c.w.egr_depth, c.w.rank);
  cast_copy(slot, wire_t, src, src_t, elems, c.it.ratio_log);
  __syncthreads();
  if (threadIdx.x == 0) {
    Ctrl *pc = c.ctrl(peer);
    EgrHdr *h = &pc->egr_hdr[ch][v % c.w.egr_depth][c.w.rank];
    st_relaxed_sys(&h->tag, tag);
    st_relaxed_sys(&h->bytes, static_cast<uint32_t>(wire_bytes(wire_t, c.it.ratio_log, elems)));
    st_relaxed_sys(&h->elems, elems);
    st_relaxed_sys(&h->kind, c.it.desc.scenario | (wire_t << 8));
    st_release_sys(&pc->egr_sig[ch][c.w.rank], v);
    c.me->egr_sent[ch][peer] = v;
  }
}

// wait for the next message of `peer`; returns its payload (nullptr on timeout)
__device__ __forceinline__ const char *egr_wait(const Ctx &c, uint32_t peer_cr, uint32_t want_tag, uint32_t want_elems,
                                                 uint32_t wire_t, uint32_t *s_tmp) {
  const uint32_t ch = static_cast<uint32_t>(c.cta);
  const uint32_t peer = c.g(peer_cr);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t e = c.me->egr_expect[ch][peer] + 1;
    bool ok = wait_ge(&c.me->egr_sig[ch][peer], e, c, RECEIVE_TIMEOUT_ERROR);
    if (ok) {
      const EgrHdr *h = &c.me->egr_hdr[ch][e % c.w.egr_depth][peer];
      const uint32_t tag = ld_relaxed_sys(&h->tag);
      if (want_tag != TAG_ANY && tag != TAG_ANY && tag != want_tag) atomicOr(c.err, DMA_TAG_MISMATCH_ERROR);
      if (ld_relaxed_sys(&h->elems) != want_elems) atomicOr(c.err, DMA_NOT_EXPECTED_BTT_ERROR);
      if ((ld_relaxed_sys(&h->kind) >> 8) != wire_t) atomicOr(c.err, COMPRESSION_ERROR);
    }
    c.me->egr_expect[ch][peer] = e;
    *s_tmp = ok ? e : 0xFFFFFFFFu;
  }
  __syncthreads();
  const uint32_t e = *s_tmp;
  if (e == 0xFFFFFFFFu) return nullptr;
  return c.heap(c.w.rank) + egr_slot_off(c.w, ch, e % c.w.egr_depth, peer);
}

// all threads are done with the slot: hand the credit back to the sender
__device__ __forceinline__ void egr_ack(const Ctx &c, uint32_t peer_cr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t ch = static_cast<uint32_t>(c.cta);
    const uint32_t peer = c.g(peer_cr);
    st_relaxed_sys(&c.ctrl(peer)->egr_ack[ch][c.w.rank], c.me->egr_expect[ch][peer]);
  }
}

// segment geometry shared by both ends: the block of `count` elements is cut
// into nctas equal parts (one per channel), each part into slot-sized segments
__device__ __forceinline__ uint32_t egr_seg_elems(const Ctx &c, uint32_t wire_t) {
  const uint32_t rl = c.it.ratio_log;
  if (is_fp8_dt(wire_t) && rl) return c.w.egr_slot_bytes / ((1u << rl) + 4) * (1u << rl);
  return c.w.egr_slot_bytes / esize(wire_t);
}

struct EgrPlan {
  size_t part_off, part_elems;
  uint32_t seg;
};
__device__ __forceinline__ EgrPlan egr_plan(const Ctx &c, size_t count, uint32_t wire_t) {
  EgrPlan p;
  const size_t per = (count + c.nctas - 1) / c.nctas;
  // keep parts 16-byte aligned for the vector paths
  const size_t al = 16;
  const size_t per_al = (per + al - 1) / al * al;
  p.part_off = static_cast<size_t>(c.cta) * per_al;
  p.part_elems = p.part_off >= count ? 0 : (count - p.part_off < per_al ? count - p.part_off : per_al);
  p.seg = egr_seg_elems(c, wire_t);
  return p;
}

// reduce my local block with the slots of all peers into dst (comm-rank order
// so every rank produces bit-identical sums)
__device__ __forceinline__ void egr_reduce_consume(const Ctx &c, const char *local, uint32_t local_t, const char *const *slots,
                                                    uint32_t wire_t, char *dst, uint32_t dst_t, uint32_t n) {
  const uint32_t P = c.P();
  const uint32_t rl = c.it.ratio_log;
  if (local_t == wire_t && wire_t == dst_t && !is_fp8_dt(wire_t)) {
    __syncthreads();
    if (threadIdx.x < P) {
      const uint32_t q = threadIdx.x;
      c.tab->src[q] = q == c.r() ? local : slots[q];
    }
    if (threadIdx.x == 0) c.tab->dst[0] = dst;
    __syncthreads();
    reduce_dispatch(c.tab, static_cast<int>(P), 1, n, wire_t, c.it.desc.function, 0, 1, c.err);
    return;
  }
  const bool sum = c.it.desc.function == static_cast<uint32_t>(reduceFunction::SUM);
  const uint32_t at = c.it.arith_compressed ? c.it.cdtype : c.it.udtype;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0;
    for (uint32_t q = 0; q < P; ++q) {
      const double x = q == c.r() ? load_as_double(local, local_t, i) : wire_load(slots[q], wire_t, rl, n, i);
      if (q == 0) acc = x;
      else acc = sum ? acc + x : (acc > x ? acc : x);
      if (at == static_cast<uint32_t>(dataType::float16)) acc = __half2float(__float2half_rn(static_cast<float>(acc)));
      else if (at != static_cast<uint32_t>(dataType::float64)) acc = static_cast<float>(acc);
    }
    store_from_double(dst, dst_t, i, acc);
  }
}

enum EgrPattern : uint32_t {
  EP_ALLREDUCE, EP_REDUCE_SCATTER, EP_ALLGATHER, EP_BCAST, EP_SCATTER, EP_GATHER, EP_REDUCE, EP_ALLTOALL
};

// One generic eager exchange.  Per segment: push what each peer needs from
// me, wait for what I need from each peer, consume, acknowledge.  Flag
// traffic (credits, headers, arrival counters, acks) is handled by one thread
// per peer in parallel so the latency does not grow with the world size.
__device__ __noinline__ void egr_collective(const Ctx &c, EgrPattern pat, uint32_t *s_tmp, const char **s_slots) {
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r(), root = it.desc.root_src_dst;
  const uint32_t cf = it.desc.compression_flags;
  const uint32_t src_t = (cf & 1) ? it.cdtype : it.udtype, dst_t = (cf & 4) ? it.cdtype : it.udtype;
  const uint32_t wire_t = (cf & 8) ? it.cdtype : it.udtype;
  const size_t count = it.desc.count;
  const char *src = c.heap(c.w.rank) + it.desc.addr0();
  char *dst = c.heap(c.w.rank) + it.desc.addr2();
  if (pat == EP_BCAST) dst = const_cast<char *>(src); // bcast works in place on addr0
  const EgrPlan pl = egr_plan(c, count, wire_t);
  const uint32_t tag = it.desc.tag;
  const size_t se = esize(src_t), de = esize(dst_t);
  const uint32_t ch = static_cast<uint32_t>(c.cta);
  __shared__ uint32_t s_v[ACCL_MAX_RANKS];
  (void)s_tmp;
  const uint32_t t = threadIdx.x;

  // which peers do I push to / expect from?
  auto pushes_to = [&](uint32_t q) -> bool {
    if (q == me) return false;
    switch (pat) {
    case EP_ALLREDUCE: case EP_ALLGATHER: case EP_REDUCE_SCATTER: case EP_ALLTOALL: return true;
    case EP_BCAST: case EP_SCATTER: return me == root;
    case EP_GATHER: case EP_REDUCE: return q == root;
    }
    return false;
  };
  auto expects_from = [&](uint32_t q) -> bool {
    if (q == me) return false;
    switch (pat) {
    case EP_ALLREDUCE: case EP_ALLGATHER: case EP_REDUCE_SCATTER: case EP_ALLTOALL: return true;
    case EP_BCAST: case EP_SCATTER: return me != root && q == root;
    case EP_GATHER: case EP_REDUCE: return me == root;
    }
    return false;
  };
  const bool same_src = pat == EP_ALLREDUCE || pat == EP_ALLGATHER || pat == EP_BCAST || pat == EP_GATHER || pat == EP_REDUCE;

  for (size_t off = 0; off < pl.part_elems; off += pl.seg) {
    const uint32_t n = static_cast<uint32_t>(pl.part_elems - off < pl.seg ? pl.part_elems - off : pl.seg);
    const size_t e0 = pl.part_off + off; // element offset inside a block
    __syncthreads();
    if (t < ACCL_MAX_RANKS) s_slots[t] = nullptr;
    // ---- push phase: credits in parallel
    if (t < P && pushes_to(t)) {
      const uint32_t peer = c.g(t);
      const uint32_t v = c.me->egr_sent[ch][peer] + 1;
      if (v > c.w.egr_depth) wait_ge(&c.me->egr_ack[ch][peer], v - c.w.egr_depth, c, DEQUEUE_BUFFER_TIMEOUT_ERROR);
      s_v[t] = v;
    }
    __syncthreads();
    if (same_src && src_t == wire_t && !is_fp8_dt(wire_t)) {
      // one read of my block, stores fan out to every destination slot
      int nd = 0;
      if (t == 0) {
        for (uint32_t k = 1; k < P; ++k) {
          const uint32_t q = (me + k) % P;
          if (pushes_to(q)) c.tab->dst[nd++] = c.heap(c.g(q)) + egr_slot_off(c.w, ch, s_v[q] % c.w.egr_depth, c.w.rank);
        }
        c.tab->src[0] = src + e0 * se;
        *s_tmp = static_cast<uint32_t>(nd);
      }
      __syncthreads();
      nd = static_cast<int>(*s_tmp);
      if (nd) copy_dispatch(c.tab, nd, static_cast<size_t>(n) * se, 0, 1);
    } else {
      for (uint32_t k = 1; k < P; ++k) {
        const uint32_t q = (me + k) % P; // stagger destinations across ranks
        if (!pushes_to(q)) continue;
        const char *from = same_src ? src + e0 * se : src + (static_cast<size_t>(q) * count + e0) * se;
        char *slot = c.heap(c.g(q)) + egr_slot_off(c.w, ch, s_v[q] % c.w.egr_depth, c.w.rank);
        cast_copy(slot, wire_t, from, src_t, n, it.ratio_log);
      }
    }
    __syncthreads();
    if (t < P && pushes_to(t)) { // headers + arrival flags in parallel
      const uint32_t peer = c.g(t), v = s_v[t];
      Ctrl *pc = c.ctrl(peer);
      EgrHdr *h = &pc->egr_hdr[ch][v % c.w.egr_depth][c.w.rank];
      st_relaxed_sys(&h->tag, tag);
      st_relaxed_sys(&h->bytes, static_cast<uint32_t>(wire_bytes(wire_t, it.ratio_log, n)));
      st_relaxed_sys(&h->elems, n);
      st_relaxed_sys(&h->kind, it.desc.scenario | (wire_t << 8));
      fence_acq_rel_sys(); // the payload was written by other threads of this CTA (ordered by the barrier above)
      st_release_sys(&pc->egr_sig[ch][c.w.rank], v);
      c.me->egr_sent[ch][peer] = v;
    }
    // ---- wait phase: one thread per expected peer
    if (t < P && expects_from(t)) {
      const uint32_t peer = c.g(t);
      const uint32_t e = c.me->egr_expect[ch][peer] + 1;
      if (wait_ge(&c.me->egr_sig[ch][peer], e, c, RECEIVE_TIMEOUT_ERROR)) {
        const EgrHdr *h = &c.me->egr_hdr[ch][e % c.w.egr_depth][peer];
        const uint32_t htag = ld_relaxed_sys(&h->tag);
        if (tag != TAG_ANY && htag != TAG_ANY && htag != tag) atomicOr(c.err, DMA_TAG_MISMATCH_ERROR);
        if (ld_relaxed_sys(&h->elems) != n) atomicOr(c.err, DMA_NOT_EXPECTED_BTT_ERROR);
        if ((ld_relaxed_sys(&h->kind) >> 8) != wire_t) atomicOr(c.err, COMPRESSION_ERROR);
        s_slots[t] = c.heap(c.w.rank) + egr_slot_off(c.w, ch, e % c.w.egr_depth, peer);
      }
      c.me->egr_expect[ch][peer] = e;
    }
    __syncthreads();
    bool ok = *c.err == 0;
    // ---- consume phase
    if (ok) {
      switch (pat) {
      case EP_ALLREDUCE:
        egr_reduce_consume(c, src + e0 * se, src_t, s_slots, wire_t, dst + e0 * de, dst_t, n);
        break;
      case EP_REDUCE_SCATTER:
        egr_reduce_consume(c, src + (static_cast<size_t>(me) * count + e0) * se, src_t, s_slots, wire_t, dst + e0 * de, dst_t, n);
        break;
      case EP_REDUCE:
        if (me == root) egr_reduce_consume(c, src + e0 * se, src_t, s_slots, wire_t, dst + e0 * de, dst_t, n);
        break;
      case EP_ALLGATHER: case EP_GATHER: case EP_ALLTOALL:
        if (pat != EP_GATHER || me == root)
          for (uint32_t q = 0; q < P; ++q) {
            char *to = dst + (static_cast<size_t>(q) * count + e0) * de;
            if (q == me) {
              const char *own = pat == EP_ALLTOALL ? src + (static_cast<size_t>(me) * count + e0) * se : src + e0 * se;
              cast_copy(to, dst_t, own, src_t, n, it.ratio_log);
            } else if (s_slots[q]) {
              cast_copy(to, dst_t, s_slots[q], wire_t, n, it.ratio_log);
            }
          }
        break;
      case EP_BCAST:
        if (me != root && s_slots[root]) cast_copy(dst + e0 * se, src_t, s_slots[root], wire_t, n, it.ratio_log);
        break;
      case EP_SCATTER:
        if (me == root) cast_copy(dst + e0 * de, dst_t, src + (static_cast<size_t>(me) * count + e0) * se, src_t, n, it.ratio_log);
        else if (s_slots[root]) cast_copy(dst + e0 * de, dst_t, s_slots[root], wire_t, n, it.ratio_log);
        break;
      }
    }
    // ---- credits back to the senders
    __syncthreads();
    if (t < P && s_slots[t]) st_relaxed_sys(&c.ctrl(c.g(t))->egr_ack[ch][c.w.rank], c.me->egr_expect[ch][c.g(t)]);
  }
  __syncthreads();
}

// eager point-to-point (channel 0 only)
__device__ __noinline__ void egr_send(const Ctx &c, uint32_t *s_tmp) {
  const WorkItem &it = c.it;
  if (c.cta != 0) return;
  const uint32_t cf = it.desc.compression_flags;
  const uint32_t src_t = (cf & 1) ? it.cdtype : it.udtype, wire_t = (cf & 8) ? it.cdtype : it.udtype;
  const char *src = c.heap(c.w.rank) + it.desc.addr0();
  const uint32_t seg = egr_seg_elems(c, wire_t);
  for (size_t off = 0; off < it.desc.count; off += seg) {
    const uint32_t n = static_cast<uint32_t>(it.desc.count - off < seg ? it.desc.count - off : seg);
    egr_push(c, it.desc.root_src_dst, src + off * esize(src_t), src_t, wire_t, n, it.desc.tag, s_tmp);
    __syncthreads();
  }
}

__device__ __noinline__ void egr_recv(const Ctx &c, uint32_t *s_tmp) {
  const WorkItem &it = c.it;
  if (c.cta != 0) return;
  const uint32_t cf = it.desc.compression_flags;
  const uint32_t dst_t = (cf & 4) ? it.cdtype : it.udtype, wire_t = (cf & 8) ? it.cdtype : it.udtype;
  char *dst = c.heap(c.w.rank) + it.desc.addr2();
  const uint32_t seg = egr_seg_elems(c, wire_t);
  for (size_t off = 0; off < it.desc.count; off += seg) {
    const uint32_t n = static_cast<uint32_t>(it.desc.count - off < seg ? it.desc.count - off : seg);
    const char *sp = egr_wait(c, it.desc.root_src_dst, it.desc.tag, n, wire_t, s_tmp);
    if (!sp) return;
    cast_copy(dst + off * esize(dst_t), dst_t, sp, wire_t, n, it.ratio_log);
    egr_ack(c, it.desc.root_src_dst);
    __syncthreads();
  }
}

// ------------------------------------------------------ rendezvous bodies
__device__ __forceinline__ bool all_equal(const uint64_t *v, uint32_t P) {
  bool eq = true;
  for (uint32_t q = 1; q < P; ++q) eq = eq && v[q] == v[0];
  return eq;
}

__device__ __forceinline__ bool all_aligned16(const uint64_t *v, uint32_t P) {
  bool ok = true;
  for (uint32_t q = 0; q < P; ++q) ok = ok && (v[q] & 15) == 0;
  return ok;
}

// fill the CTA's pointer table: src[k] / dst[k] for k < P, rotated so that my
// own buffers come first (staggers the peers each rank touches first)
__device__ __forceinline__ void fill_table(const Ctx &c, const uint64_t *s_off0, const uint64_t *s_off2, uint64_t src_add,
                                           uint64_t dst_add, bool rotate) {
  __syncthreads();
  if (threadIdx.x < c.P()) {
    const uint32_t k = threadIdx.x;
    const uint32_t q = rotate ? (c.r() + k) % c.P() : k;
    if (s_off0) c.tab->src[k] = c.heap(c.g(q)) + s_off0[q] + src_add;
    if (s_off2) c.tab->dst[k] = c.heap(c.g(q)) + s_off2[q] + dst_add;
  }
  __syncthreads();
}

// allreduce, two-shot.  NVLS: the switch reduces my shard (multimem.ld_reduce)
// and broadcasts it (multimem.st).  P2P: pull my shard from every peer,
// reduce on the SM, store it into every peer's destination.
// `rvb_*` are the wait-free data phases (between the entry and the exit meeting): the direct-launch kernel
// brackets them with chan_sync, the persistent engine runs them as moves on its worker CTAs.
__device__ __noinline__ void rvb_allreduce(const Ctx &c, const uint64_t *s_off0, const uint64_t *s_off2) {
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r();
  const uint32_t dt = it.udtype;
  const size_t count = it.desc.count, es = esize(dt);
  const size_t per_vec = 16 / es;
  const size_t nvec = count / per_vec;
  const size_t shard = (nvec + P - 1) / P;
  const size_t v0 = static_cast<size_t>(me) * shard < nvec ? static_cast<size_t>(me) * shard : nvec;
  const size_t v1 = v0 + shard < nvec ? v0 + shard : nvec;
  const NvOp nop = nvls_op(dt, it.desc.function);
  const bool sym = all_equal(s_off0, P) && all_equal(s_off2, P) && (s_off0[0] & 15) == 0 && (s_off2[0] & 15) == 0;
  const bool nv = (it.flags & WF_USE_MC) && it.algo == ALGO_NVLS && nop != NvOp::none && sym;
  const uint32_t h16 = it.tune.hybrid_16ths;
  if (nv && h16 && h16 < 16 && c.nctas >= 16 && v1 - v0 >= (1u << 16)) {
    // hybrid: the switch serves multimem.ld_reduce below the rate the links carry plain peer traffic at
    // (profiles/SWEEPS.md): h16 / 16 of my shard (and of the channels) go through the peer two-shot body so
    // both limits are used at once
    const int n_p2p = c.nctas * static_cast<int>(h16) / 16 > 0 ? c.nctas * static_cast<int>(h16) / 16 : 1;
    const int n_nv = c.nctas - n_p2p;
    const size_t vm = v1 - (v1 - v0) * h16 / 16;
    if (c.cta < n_nv) {
      nvls_reduce_dispatch<true>(nop, it.tune.nvls_unroll, c.w.mc + s_off0[0], c.w.mc + s_off2[0], v0, vm, c.cta, n_nv);
    } else {
      fill_table(c, s_off0, s_off2, vm * 16, vm * 16, true);
      reduce_dispatch(c.tab, static_cast<int>(P), static_cast<int>(P), (v1 - vm) * per_vec, dt, it.desc.function, c.cta - n_nv,
                      n_p2p, c.err);
    }
  } else if (nv) {
    nvls_reduce_dispatch<true>(nop, it.tune.nvls_unroll, c.w.mc + s_off0[0], c.w.mc + s_off2[0], v0, v1, c.cta, c.nctas);
  } else {
    fill_table(c, s_off0, s_off2, v0 * 16, v0 * 16, true);
    reduce_dispatch(c.tab, static_cast<int>(P), static_cast<int>(P), (v1 - v0) * per_vec, dt, it.desc.function, c.cta,
                    c.nctas, c.err);
  }
  // sub-vector tail: communicator rank 0 handles it element-wise for everybody
  const size_t tail0 = nvec * per_vec;
  if (tail0 < count && me == 0 && c.cta == 0) {
    fill_table(c, s_off0, s_off2, tail0 * es, tail0 * es, false);
    reduce_dispatch(c.tab, static_cast<int>(P), static_cast<int>(P), count - tail0, dt, it.desc.function, 0, 1, c.err);
  }
}

__device__ __noinline__ void rv_allreduce(const Ctx &c, uint64_t *s_off0, uint64_t *s_off2) {
  chan_sync(c, true, c.it.desc.addr0(), c.it.desc.addr2(), s_off0, s_off2);
  if (*c.err == 0) rvb_allreduce(c, s_off0, s_off2);
  chan_sync(c, false, 0, 0, nullptr, nullptr);
}

// every rank pulls everything and reduces locally: no second hop, for
// messages too big for the slots but small enough that latency dominates
__device__ __noinline__ void rvb_allreduce_oneshot(const Ctx &c, const uint64_t *s_off0, const uint64_t *s_off2) {
  const WorkItem &it = c.it;
  const uint32_t P = c.P();
  // In place on ANY rank (source == destination): that rank's result stores would race with the peers' reads of
  // its source.  The exchanged offsets are identical everywhere, so all ranks take the two-shot body together
  // (only a shard's owner writes it, after having read it).
  bool inplace = false;
  for (uint32_t q = 0; q < P; ++q) inplace = inplace || s_off0[q] == s_off2[q];
  if (inplace) {
    rvb_allreduce(c, s_off0, s_off2);
    return;
  }
  const NvOp nop = nvls_op(it.udtype, it.desc.function);
  const size_t es = esize(it.udtype), per_vec = 16 / es, nvec = it.desc.count / per_vec;
  const bool sym = all_equal(s_off0, P) && (s_off0[0] & 15) == 0 && (it.desc.addr2() & 15) == 0;
  size_t done = 0;
  if ((it.flags & WF_USE_MC) && nop != NvOp::none && sym) {
    nvls_reduce_dispatch<false>(nop, it.tune.nvls_unroll, c.w.mc + s_off0[0], c.heap(c.w.rank) + it.desc.addr2(), 0, nvec, c.cta, c.nctas);
    done = nvec * per_vec;
  }
  if (done < it.desc.count) {
    fill_table(c, s_off0, nullptr, done * es, 0, false); // communicator-rank order: identical sums everywhere
    if (threadIdx.x == 0) c.tab->dst[0] = c.heap(c.w.rank) + it.desc.addr2() + done * es;
    __syncthreads();
    reduce_dispatch(c.tab, static_cast<int>(P), 1, it.desc.count - done, it.udtype, it.desc.function, c.cta, c.nctas, c.err);
  }
}
__device__ __noinline__ void rv_allreduce_oneshot(const Ctx &c, uint64_t *s_off0, uint64_t *s_off2) {
  chan_sync(c, true, c.it.desc.addr0(), c.it.desc.addr2(), s_off0, s_off2);
  if (*c.err == 0) rvb_allreduce_oneshot(c, s_off0, s_off2);
  chan_sync(c, false, 0, 0, nullptr, nullptr);
}

__device__ __noinline__ void rvb_reduce_scatter(const Ctx &c, const uint64_t *s_off0, const uint64_t *s_off2) {
  (void)s_off2;
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r();
  const size_t count = it.desc.count, es = esize(it.udtype);
  const size_t blk_bytes = count * es;
  const NvOp nop = nvls_op(it.udtype, it.desc.function);
  char *dst = c.heap(c.w.rank) + it.desc.addr2();
  const bool sym = all_equal(s_off0, P) && ((s_off0[0] + me * blk_bytes) & 15) == 0 && (it.desc.addr2() & 15) == 0;
  size_t done = 0;
  if ((it.flags & WF_USE_MC) && it.algo == ALGO_NVLS && nop != NvOp::none && sym) {
    const size_t nvec = blk_bytes / 16;
    nvls_reduce_dispatch<false>(nop, it.tune.nvls_unroll, c.w.mc + s_off0[0] + me * blk_bytes, dst, 0, nvec, c.cta, c.nctas);
    done = nvec * 16 / es;
  }
  if (done < count) {
    fill_table(c, s_off0, nullptr, me * blk_bytes + done * es, 0, true);
    if (threadIdx.x == 0) c.tab->dst[0] = dst + done * es;
    __syncthreads();
    reduce_dispatch(c.tab, static_cast<int>(P), 1, count - done, it.udtype, it.desc.function, c.cta, c.nctas, c.err);
  }
}
__device__ __noinline__ void rv_reduce_scatter(const Ctx &c, uint64_t *s_off0, uint64_t *s_off2) {
  chan_sync(c, true, c.it.desc.addr0(), c.it.desc.addr2(), s_off0, s_off2);
  if (*c.err == 0) rvb_reduce_scatter(c, s_off0, s_off2);
  chan_sync(c, false, 0, 0, nullptr, nullptr);
}

__device__ __noinline__ void rv_bcast_pipelined(const Ctx &c, const uint64_t *s_off);
__device__ __noinline__ void rv_bcast_flags(const Ctx &c, const uint64_t *s_off);

// push-style data movement shared by allgather / bcast / scatter / gather / alltoall.
// `stepwise_ok`: the caller can run the pipelined broadcasts, which meet the peers between chunks (direct launches
// only: inside the engine every wait must be a parked step, so it takes the single-pass bodies).
__device__ __noinline__ void rvb_move(const Ctx &c, EgrPattern pat, const uint64_t *s_off0, const uint64_t *s_off2, bool stepwise_ok) {
  (void)s_off0;
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r(), root = it.desc.root_src_dst;
  const size_t blk = static_cast<size_t>(it.desc.count) * esize(it.udtype);
  const char *src = c.heap(c.w.rank) + it.desc.addr0();
  const bool mc_ok = (it.flags & WF_USE_MC) && it.algo == ALGO_NVLS && all_equal(s_off2, P) && (blk & 15) == 0 &&
                     (s_off2[0] & 15) == 0 && (it.desc.addr0() & 15) == 0;
  // tune.bcast_flags == 2: never pipeline, the root stores through the switch (or to every peer) in one pass
  bool bcast_pipelined = stepwise_ok && pat == EP_BCAST && P >= 3 && blk >= (128u << 20) && (blk & 15) == 0 && it.tune.bcast_flags != 2;
  for (uint32_t q = 0; q < P; ++q) bcast_pipelined = bcast_pipelined && (s_off2[q] & 15) == 0;
  switch (pat) {
  case EP_ALLGATHER:
    if (mc_ok) {
      nvls_bcast_range(src, c.w.mc + s_off2[0] + me * blk, blk / 16, c.cta, c.nctas);
    } else {
      fill_table(c, nullptr, s_off2, 0, me * blk, true);
      if (threadIdx.x == 0) c.tab->src[0] = src;
      __syncthreads();
      copy_dispatch(c.tab, static_cast<int>(P), blk, c.cta, c.nctas);
    }
    break;
  case EP_BCAST:
    if (bcast_pipelined) {
      if (it.tune.bcast_flags == 1) rv_bcast_flags(c, s_off2);
      else rv_bcast_pipelined(c, s_off2);
    } else if (me == root) {
      if (mc_ok) {
        // the multicast store also rewrites the root's own copy with identical bytes
        nvls_bcast_range(src, c.w.mc + s_off2[0], blk / 16, c.cta, c.nctas);
      } else {
        fill_table(c, nullptr, s_off2, 0, 0, true); // dst[0] is my own buffer: skip it
        if (threadIdx.x == 0) {
          c.tab->src[0] = src;
          for (uint32_t k = 1; k < P; ++k) c.tab->dst[k - 1] = c.tab->dst[k];
        }
        __syncthreads();
        copy_dispatch(c.tab, static_cast<int>(P) - 1, blk, c.cta, c.nctas);
      }
    }
    break;
  case EP_SCATTER:
    if (me == root)
      for (uint32_t k = 0; k < P; ++k) {
        // CTAs visit the destinations in different orders: all P switch ports are fed at once
        // instead of bursting one block at a time into a single port
        const uint32_t q = (me + k + static_cast<uint32_t>(c.cta)) % P;
        copy_simple(c.heap(c.g(q)) + s_off2[q], src + q * blk, blk, c.cta, c.nctas);
      }
    break;
  case EP_GATHER:
    copy_simple(c.heap(c.g(root)) + s_off2[root] + me * blk, src, blk, c.cta, c.nctas);
    break;
  case EP_ALLTOALL:
    for (uint32_t k = 0; k < P; ++k) {
      const uint32_t q = (me + k + static_cast<uint32_t>(c.cta)) % P;
      copy_simple(c.heap(c.g(q)) + s_off2[q] + me * blk, src + q * blk, blk, c.cta, c.nctas);
    }
    break;
  default: break;
  }
}
__device__ __noinline__ void rv_move(const Ctx &c, EgrPattern pat, uint64_t *s_off0, uint64_t *s_off2) {
  // bcast keeps its buffer in addr0 on every rank; the others receive into addr2
  const uint64_t my_dst = pat == EP_BCAST ? c.it.desc.addr0() : c.it.desc.addr2();
  chan_sync(c, true, c.it.desc.addr0(), my_dst, s_off0, s_off2);
  if (*c.err == 0) rvb_move(c, pat, s_off0, s_off2, true);
  chan_sync(c, false, 0, 0, nullptr, nullptr);
}

// Write-only rooted reduce (tune.reduce_push): the message is cut into P-1 slices (one per non-root "worker") and every CTA
// owns a stripe of every slice, processed in chunks of CH vectors.  In step st every rank PUSHES its chunk
// st of slice j into worker j's scratch (stage st & 1, one region per source), and worker j reduces chunk
// st - 1 from its P-1 scratch regions plus its own operand, storing the result straight into the root's
// buffer.  Steps run in lockstep (one chan_sync per step, as in rv_bcast_pipelined), which orders the pushes
// of step st before their reduction in step st + 1 and the reuse of a stage after its reduction.  Every port
// carries N bytes of posted writes in and out; nobody issues remote loads.
// Returns the number of vectors it reduced (0: geometry does not fit, caller falls back).
__device__ __noinline__ size_t rv_reduce_push(const Ctx &c, const uint64_t *s_off0, const uint64_t *s_off2) {
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r(), root = it.desc.root_src_dst, W = P - 1;
  const size_t es = esize(it.udtype);
  const size_t nvec = static_cast<size_t>(it.desc.count) * es / 16;
  // scratch per heap: [stage 2][cta][source rank P][CH vectors]
  size_t CH = c.w.scr_bytes / (static_cast<size_t>(2) * c.nctas * P * 16);
  if (CH > 16384) CH = 16384;
  CH &= ~static_cast<size_t>(31); // whole warps of vectors
  if (CH < 256 || W < 2) return 0;
  const size_t per_slice = (nvec + W - 1) / W;
  const size_t per_cta = (per_slice + c.nctas - 1) / c.nctas;
  const size_t steps = (per_cta + CH - 1) / CH;
  const uint32_t j_me = (me + P - root - 1) % P; // my worker index (meaningless on the root)
  const char *src = c.heap(c.w.rank) + s_off0[me];
  char *root_dst = c.heap(c.g(root)) + s_off2[root];
  auto range = [&](uint32_t j, size_t step, size_t &a, size_t &b) {
    a = j * per_slice + static_cast<size_t>(c.cta) * per_cta + step * CH;
    b = a + CH;
    const size_t lim_cta = j * per_slice + (static_cast<size_t>(c.cta) + 1) * per_cta;
    const size_t lim_slice = (j + 1) * per_slice < nvec ? (j + 1) * per_slice : nvec;
    if (b > lim_cta) b = lim_cta;
    if (b > lim_slice) b = lim_slice;
  };
  auto scr = [&](uint32_t owner_grank, size_t stage, uint32_t src_rank) -> char * {
    return c.heap(owner_grank) + c.w.scr_off + (((stage * c.nctas + c.cta) * P + src_rank) * CH) * 16;
  };
  for (size_t st = 0; st <= steps; ++st) {
    if (st < steps) {
      // deal my chunk of every slice to the worker that owns it (my own slice stays where it is)
      for (uint32_t k = 0; k < W; ++k) {
        const uint32_t j = (k + static_cast<uint32_t>(c.cta)) % W; // CTAs start on different workers
        const uint32_t q = (root + 1 + j) % P;
        if (q == me) continue;
        size_t a, b;
        range(j, st, a, b);
        if (a < b) copy_simple(scr(c.g(q), st & 1, me), src + a * 16, (b - a) * 16, 0, 1);
      }
    }
    if (me != root && st >= 1) {
      size_t a, b;
      range(j_me, st - 1, a, b);
      if (a < b) {
        __syncthreads();
        if (threadIdx.x == 0) {
          c.tab->src[0] = src + a * 16;
          for (uint32_t k = 1; k < P; ++k) c.tab->src[k] = scr(c.w.rank, (st - 1) & 1, (me + k) % P);
          c.tab->dst[0] = root_dst + a * 16;
        }
        __syncthreads();
        reduce_dispatch(c.tab, static_cast<int>(P), 1, (b - a) * (16 / es), it.udtype, it.desc.function, 0, 1, c.err);
      }
    }
    chan_sync(c, false, 0, 0, nullptr, nullptr);
  }
  return nvec;
}

// reduce to root.  Small: the root pulls (or lets the switch reduce) everything.  Large
// (P >= 3): the P-1 other ranks each reduce one slice and store it into the root's buffer,
// so the root's inbound link carries N bytes instead of (P-1) N.
__device__ __noinline__ void rvb_reduce(const Ctx &c, const uint64_t *s_off0, const uint64_t *s_off2, bool stepwise_ok) {
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r(), root = it.desc.root_src_dst;
  const size_t es = esize(it.udtype), count = it.desc.count;
  const NvOp nop = nvls_op(it.udtype, it.desc.function);
  const bool sym = all_equal(s_off0, P) && (s_off0[0] & 15) == 0 && (s_off2[root] & 15) == 0;
  const bool use_mc = (it.flags & WF_USE_MC) && it.algo == ALGO_NVLS && nop != NvOp::none && sym;
  char *root_dst = c.heap(c.g(root)) + s_off2[root];
  const size_t nvec = count * es / 16;
  // tune.reduce_push == 2: the root lets the switch reduce the whole message (one hop, root inbound = N) at every size
  const bool distributed = P >= 3 && count * es >= (1u << 20) && (s_off2[root] & 15) == 0 && !(it.tune.reduce_push == 2 && use_mc);
  size_t done = 0;
  // all ranks take the same decision: it only depends on the call and on the (identical) geometry
  if (stepwise_ok && it.tune.reduce_push == 1 && distributed && count * es >= (8u << 20) && all_aligned16(s_off0, P))
    done = rv_reduce_push(c, s_off0, s_off2) * 16 / es;
  if (done) {
    // handled above; the sub-vector tail (if any) is reduced by the root below
  } else if (distributed) {
    if (me != root) {
      const uint32_t j = (me + P - root - 1) % P; // my index among the P-1 workers
      const size_t per = (nvec + (P - 1) - 1) / (P - 1);
      const size_t v0 = j * per < nvec ? j * per : nvec, v1 = v0 + per < nvec ? v0 + per : nvec;
      // peer pulls: every worker's inbound link carries ~N, like the root's; measured faster than letting the
      // switch reduce the slices (profiles/SWEEPS.md)
      fill_table(c, s_off0, nullptr, v0 * 16, 0, true);
      if (threadIdx.x == 0) c.tab->dst[0] = root_dst + v0 * 16;
      __syncthreads();
      reduce_dispatch(c.tab, static_cast<int>(P), 1, (v1 - v0) * (16 / es), it.udtype, it.desc.function, c.cta, c.nctas, c.err);
    }
    done = nvec * 16 / es; // the sub-vector tail (if any) is reduced by the root below
  } else if (me == root && use_mc) {
    nvls_reduce_dispatch<false>(nop, it.tune.nvls_unroll, c.w.mc + s_off0[0], root_dst, 0, nvec, c.cta, c.nctas);
    done = nvec * 16 / es;
  }
  if (me == root && done < count) {
    fill_table(c, s_off0, nullptr, done * es, 0, false);
    if (threadIdx.x == 0) c.tab->dst[0] = root_dst + done * es;
    __syncthreads();
    reduce_dispatch(c.tab, static_cast<int>(P), 1, count - done, it.udtype, it.desc.function, c.cta, c.nctas, c.err);
  }
}
__device__ __noinline__ void rv_reduce(const Ctx &c, uint64_t *s_off0, uint64_t *s_off2) {
  chan_sync(c, true, c.it.desc.addr0(), c.it.desc.addr2(), s_off0, s_off2);
  if (*c.err == 0) rvb_reduce(c, s_off0, s_off2, true);
  chan_sync(c, false, 0, 0, nullptr, nullptr);
}

// Large broadcast, software-pipelined in two overlapping stages: the root deals slice j of its
// buffer to worker j (root outbound = N in total); worker j forwards what it received to the other
// workers (worker inbound = N, outbound = N (P-2)/(P-1)).  Chunk c is forwarded while chunk c+1 is
// dealt, each CTA running its own stripe of every slice with its own flags.
__device__ __noinline__ void rv_bcast_pipelined(const Ctx &c, const uint64_t *s_off) {
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r(), root = it.desc.root_src_dst;
  const size_t nvec = static_cast<size_t>(it.desc.count) * esize(it.udtype) / 16;
  const uint32_t W = P - 1;                       // workers
  const size_t per_slice = (nvec + W - 1) / W;    // vectors per worker slice
  const size_t per_cta = (per_slice + c.nctas - 1) / c.nctas;
  // per (slice, CTA) step: about 1/8 of the stripe, between 32 KiB and 256 KiB (pipeline depth vs sync cost)
  size_t CH = per_cta / 8;
  CH = CH < 2048 ? 2048 : (CH > 16384 ? 16384 : CH);
  const size_t steps = (per_cta + CH - 1) / CH;
  const uint32_t j_me = (me + P - root - 1) % P;  // my worker index (unused on the root)
  const char *src = c.heap(c.w.rank) + s_off[me];
  auto range = [&](uint32_t j, size_t step, size_t &a, size_t &b) {
    a = j * per_slice + static_cast<size_t>(c.cta) * per_cta + step * CH;
    b = a + CH;
    const size_t lim_cta = j * per_slice + (static_cast<size_t>(c.cta) + 1) * per_cta;
    const size_t lim_slice = (j + 1) * per_slice < nvec ? (j + 1) * per_slice : nvec;
    if (b > lim_cta) b = lim_cta;
    if (b > lim_slice) b = lim_slice;
  };
  for (size_t st = 0; st <= steps; ++st) {
    if (me == root && st < steps) {
      for (uint32_t k = 0; k < W; ++k) {
        const uint32_t j = (k + static_cast<uint32_t>(c.cta)) % W; // CTAs start on different workers
        const uint32_t q = (root + 1 + j) % P;
        size_t a, b;
        range(j, st, a, b);
        if (a < b) copy_simple(c.heap(c.g(q)) + s_off[q] + a * 16, src + a * 16, (b - a) * 16, 0, 1);
      }
    } else if (me != root && st >= 1 && W >= 2) {
      size_t a, b;
      range(j_me, st - 1, a, b);
      if (a < b) {
        __syncthreads();
        if (threadIdx.x == 0) {
          int nd = 0;
          for (uint32_t k = 1; k < W; ++k) { // the other workers, starting after me
            const uint32_t q = (root + 1 + (j_me + k) % W) % P;
            c.tab->dst[nd++] = c.heap(c.g(q)) + s_off[q] + a * 16;
          }
          c.tab->src[0] = src + a * 16; // what the root stored here one step ago
        }
        __syncthreads();
        copy_dispatch(c.tab, static_cast<int>(W) - 1, (b - a) * 16, 0, 1);
      }
    }
    chan_sync(c, false, 0, 0, nullptr, nullptr);
  }
}

// tune.bcast_flags: same deal-and-forward schedule as rv_bcast_pipelined, driven by one-way flags instead of a meeting per
// step: after the stores of a chunk the producer raises a per-(channel, source) counter at the consumer
// (st.release.sys); the consumer waits for the count it needs (ld.acquire.sys).  The root never waits, a
// worker waits for the root's chunk before forwarding it and, at the end, for every other worker to have
// delivered all of its chunks.  Counters are monotonic across calls (step_seen / step_sent), every
// (producer, consumer, step) signals exactly once whether or not the chunk is empty.
__device__ __noinline__ void rv_bcast_flags(const Ctx &c, const uint64_t *s_off) {
  const WorkItem &it = c.it;
  const uint32_t P = c.P(), me = c.r(), root = it.desc.root_src_dst;
  const size_t nvec = static_cast<size_t>(it.desc.count) * esize(it.udtype) / 16;
  const uint32_t W = P - 1, ch = static_cast<uint32_t>(c.cta);
  const size_t per_slice = (nvec + W - 1) / W;
  const size_t per_cta = (per_slice + c.nctas - 1) / c.nctas;
  size_t CH = per_cta / 8;
  CH = CH < 2048 ? 2048 : (CH > 16384 ? 16384 : CH);
  const uint32_t steps = static_cast<uint32_t>((per_cta + CH - 1) / CH);
  const uint32_t j_me = (me + P - root - 1) % P;
  const char *src = c.heap(c.w.rank) + s_off[me];
  auto range = [&](uint32_t j, size_t step, size_t &a, size_t &b) {
    a = j * per_slice + static_cast<size_t>(c.cta) * per_cta + step * CH;
    b = a + CH;
    const size_t lim_cta = j * per_slice + (static_cast<size_t>(c.cta) + 1) * per_cta;
    const size_t lim_slice = (j + 1) * per_slice < nvec ? (j + 1) * per_slice : nvec;
    if (b > lim_cta) b = lim_cta;
    if (b > lim_slice) b = lim_slice;
  };
  // all stores of this CTA to rank q's buffer are done: tell q (one thread, after a CTA barrier)
  auto signal = [&](uint32_t q) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t gq = c.g(q);
      const uint32_t v = c.pads().step_sent[ch][gq] + 1;
      c.pads().step_sent[ch][gq] = v;
      st_release_sys(&c.pads_of(gq).step_sig[ch][c.w.rank], v);
    }
  };
  // wait until rank q has delivered `n` chunks of this call to me
  auto await = [&](uint32_t q, uint32_t n) {
    if (threadIdx.x == 0) wait_ge(&c.pads().step_sig[ch][c.g(q)], c.pads().step_seen[ch][c.g(q)] + n, c, RECEIVE_TIMEOUT_ERROR);
    __syncthreads();
  };
  if (me == root) {
    for (uint32_t st = 0; st < steps; ++st)
      for (uint32_t k = 0; k < W; ++k) {
        const uint32_t j = (k + ch) % W; // CTAs start on different workers
        const uint32_t q = (root + 1 + j) % P;
        size_t a, b;
        range(j, st, a, b);
        if (a < b) copy_simple(c.heap(c.g(q)) + s_off[q] + a * 16, src + a * 16, (b - a) * 16, 0, 1);
        signal(q);
      }
  } else {
    for (uint32_t st = 0; st < steps; ++st) {
      await(root, st + 1);
      size_t a, b;
      range(j_me, st, a, b);
      if (a < b && W >= 2) {
        if (threadIdx.x == 0) {
          int nd = 0;
          for (uint32_t k = 1; k < W; ++k) {
            const uint32_t q = (root + 1 + (j_me + k) % W) % P;
            c.tab->dst[nd++] = c.heap(c.g(q)) + s_off[q] + a * 16;
          }
          c.tab->src[0] = src + a * 16; // what the root stored here
        }
        __syncthreads();
        copy_dispatch(c.tab, static_cast<int>(W) - 1, (b - a) * 16, 0, 1);
      }
      for (uint32_t k = 1; k < W; ++k) signal((root + 1 + (j_me + k) % W) % P);
    }
    // the other workers' slices: complete once each of them has forwarded all of its chunks to me
    for (uint32_t k = 1; k < W; ++k) await((root + 1 + (j_me + k) % W) % P, steps);
    __syncthreads();
    if (threadIdx.x == 0) {
      c.pads().step_seen[ch][c.g(root)] += steps;
      for (uint32_t k = 1; k < W; ++k) c.pads().step_seen[ch][c.g((root + 1 + (j_me + k) % W) % P)] += steps;
    }
    __syncthreads();
  }
}

// rendezvous send / recv: a pair of ranks meets on the pads, the receiver
// announces its buffer, the sender stores straight into it
__device__ __noinline__ void rv_send(const Ctx &c, uint64_t *s_pair) {
  const WorkItem &it = c.it;
  uint32_t *kind = reinterpret_cast<uint32_t *>(s_pair + 1);
  pair_sync(c, it.desc.root_src_dst, it.desc.addr0(), static_cast<uint32_t>(operation::send), s_pair, kind);
  if (*kind != static_cast<uint32_t>(operation::recv)) {
    if (threadIdx.x == 0 && *kind != 0xFFFFFFFFu) atomicOr(c.err, PACK_SEQ_NUMBER_ERROR);
  } else {
    copy_simple(c.heap(c.g(it.desc.root_src_dst)) + *s_pair, c.heap(c.w.rank) + it.desc.addr0(),
                static_cast<size_t>(it.desc.count) * esize(it.udtype), c.cta, c.nctas);
  }
  __shared__ uint64_t s_o;
  __shared__ uint32_t s_k;
  pair_sync(c, it.desc.root_src_dst, 0, static_cast<uint32_t>(operation::send), &s_o, &s_k, false);
}

__device__ __noinline__ void rv_recv(const Ctx &c, uint64_t *s_pair) {
  const WorkItem &it = c.it;
  uint32_t *kind = reinterpret_cast<uint32_t *>(s_pair + 1);
  pair_sync(c, it.desc.root_src_dst, it.desc.addr2(), static_cast<uint32_t>(operation::recv), s_pair, kind);
  if (*kind != static_cast<uint32_t>(operation::send) && *kind != 0xFFFFFFFFu && threadIdx.x == 0)
    atomicOr(c.err, PACK_SEQ_NUMBER_ERROR);
  __shared__ uint64_t s_o;
  __shared__ uint32_t s_k;
  pair_sync(c, it.desc.root_src_dst, 0, static_cast<uint32_t>(operation::recv), &s_o, &s_k, false);
}

} // namespace k
} // namespace cuda
} // namespace accl
