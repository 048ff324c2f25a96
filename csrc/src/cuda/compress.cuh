// Wire-compressed collectives at bandwidth (ALGO_WIRE): operands stay in their own dtype, what crosses NVLink is
// the `compress_dtype` — fp16 / bf16, or fp8 (e4m3 / e5m2) with one fp32 scale per 32 elements.  The cast is fused
// into the store that leaves the GPU and into the load that consumes the peer's data: 16-byte vector accesses,
// fp32 accumulation, no intermediate buffer in the operand dtype and no separate cast kernel.
//
// Reference: the compression lanes of the datapath — any operand / result / wire of a move can be routed through
// hp_compression (kernels/plugins/hp_compression/hp_compression.cpp:30-144, fp32 <-> fp16) by the router
// (kernels/cclo/hls/dma_mover/dma_mover.cpp:30-186); `ETH_COMPRESSED` = compress before the packetizer,
// decompress after the depacketizer.  bf16 and block-scaled fp8 are this library's additions
// (`elem_ratio_log` = log2 of the scale block, arithconfig.hpp).
//
// Two-shot all-reduce over a scratch area of every heap, [half][source rank][channel] regions:
//   1. every rank casts block q of its operand into region [0][me] of rank q          (reduce-scatter, compressed wire)
//   --- meeting ---
//   2. rank q folds its own block (exact) with the P - 1 compressed ones in fp32, writes its shard of the result and
//      casts it into region [1][q] of every peer                                       (all-gather, compressed wire)
//   --- meeting ---
//   3. every rank expands the P - 1 shards it received into its result buffer
// Every CTA owns one stripe of every block and one sync channel, so there is no grid-wide barrier.  Half 0 is
// written before meeting 1 and read before meeting 2, half 1 is written between the meetings and read before the
// next round's / call's meeting 1: two meetings per round, none trailing.  Messages larger than the scratch area
// run in rounds.  reduce_scatter is steps 1-2 (fold only), allgather steps 2-3 (own block forwarded as is).
#pragma once
#include "collectives.cuh"

namespace accl {
namespace cuda {
namespace k {

// ---------------------------------------------------------------- codecs
// A codec moves GROUPS of G elements between fp32 registers and the wire representation of a stripe:
// [payload: G * WB bytes per group][scales: one fp32 per group, block-scaled fp8 only]
struct WireGeom {
  size_t groups;     // groups in the stripe
  size_t scale_off;  // byte offset of the scales inside the stripe's region (0: none)
  size_t bytes;      // region bytes
};

template <typename T16> struct Codec16 { // __half / __nv_bfloat16
  static constexpr int G = 8;
  static __device__ __forceinline__ WireGeom geom(size_t elems) {
    const size_t g = (elems + G - 1) / G;
    return WireGeom{g, 0, g * 16};
  }
  static __device__ __forceinline__ void enc(char *region, const WireGeom &, size_t gi, const float (&v)[G]) {
    st_relaxed_sys16(region + gi * 16, VecOf<T16>::pack(v));
  }
  static __device__ __forceinline__ void dec(const char *region, const WireGeom &, size_t gi, float (&v)[G]) {
    VecOf<T16>::unpack(ld_relaxed_sys16(region + gi * 16), v);
  }
};

template <typename T8> struct F8Max;
template <> struct F8Max<__nv_fp8_e4m3> { static constexpr float value = 448.f; };
template <> struct F8Max<__nv_fp8_e5m2> { static constexpr float value = 57344.f; };

template <typename T8, bool SCALED> struct Codec8 { // __nv_fp8_e4m3 / __nv_fp8_e5m2, one scale per 32 elements
  static constexpr int G = 32;
  static __device__ __forceinline__ WireGeom geom(size_t elems) {
    const size_t g = (elems + G - 1) / G;
    return WireGeom{g, SCALED ? g * 32 : 0, g * 32 + (SCALED ? ((g * 4 + 15) & ~static_cast<size_t>(15)) : 0)};
  }
  static __device__ __forceinline__ void enc(char *region, const WireGeom &g, size_t gi, const float (&v)[G]) {
    float scale = 1.f;
    if (SCALED) {
      float m = 0.f;
#pragma unroll
      for (int i = 0; i < G; ++i) m = fmaxf(m, fabsf(v[i]));
      scale = m > 0.f ? m / F8Max<T8>::value : 1.f;
    }
    const float inv = 1.f / scale;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t x = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) x |= static_cast<uint32_t>(T8(v[4 * i + b] * inv).__x) << (8 * b);
      w[i] = x;
    }
    st_relaxed_sys16(region + gi * 32, Vec16{w[0], w[1], w[2], w[3]});
    st_relaxed_sys16(region + gi * 32 + 16, Vec16{w[4], w[5], w[6], w[7]});
    if (SCALED) st_relaxed_sys(reinterpret_cast<uint32_t *>(region + g.scale_off + gi * 4), __float_as_uint(scale));
  }
  static __device__ __forceinline__ void dec(const char *region, const WireGeom &g, size_t gi, float (&v)[G]) {
    const Vec16 a = ld_relaxed_sys16(region + gi * 32), b = ld_relaxed_sys16(region + gi * 32 + 16);
    const float scale = SCALED ? __uint_as_float(ld_relaxed_sys(reinterpret_cast<const uint32_t *>(region + g.scale_off + gi * 4))) : 1.f;
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        T8 t;
        t.__x = static_cast<unsigned char>(w[i] >> (8 * k));
        v[4 * i + k] = static_cast<float>(t) * scale;
      }
  }
};

// G elements of the operand dtype <-> fp32 registers (local memory, 16-byte accesses when aligned)
template <typename T, int G> __device__ __forceinline__ void load_group(const T *p, size_t avail, bool al, float (&v)[G]) {
  constexpr int PER = 16 / sizeof(T);
  if (al && avail >= static_cast<size_t>(G)) {
#pragma unroll
    for (int j = 0; j < G / PER; ++j) {
      float t[PER];
      VecOf<T>::unpack(ld_stream(reinterpret_cast<const char *>(p) + j * 16), t);
#pragma unroll
      for (int e = 0; e < PER; ++e) v[j * PER + e] = t[e];
    }
  } else {
#pragma unroll
    for (int i = 0; i < G; ++i) v[i] = static_cast<size_t>(i) < avail ? static_cast<float>(Tr<T>::up(p[i])) : 0.f;
  }
}
template <typename T, int G> __device__ __forceinline__ void store_group(T *p, size_t avail, bool al, const float (&v)[G]) {
  constexpr int PER = 16 / sizeof(T);
  if (al && avail >= static_cast<size_t>(G)) {
#pragma unroll
    for (int j = 0; j < G / PER; ++j) {
      float t[PER];
#pragma unroll
      for (int e = 0; e < PER; ++e) t[e] = v[j * PER + e];
      st_stream(reinterpret_cast<char *>(p) + j * 16, VecOf<T>::pack(t));
    }
  } else {
#pragma unroll
    for (int i = 0; i < G; ++i)
      if (static_cast<size_t>(i) < avail) p[i] = Tr<T>::down(v[i]);
  }
}

// ------------------------------------------------------------------ bodies
struct WireLayout {
  size_t blk_elems;    // elements per block (shard) in this round
  size_t stripe_elems; // elements of a block owned by one CTA (multiple of the codec group)
  size_t region;       // bytes of one [half][src][cta] region
  __device__ char *at(const Ctx &c, uint32_t grank, uint32_t half, uint32_t src_cr) const {
    return c.heap(grank) + c.w.scr_off + ((static_cast<size_t>(half) * c.P() + src_cr) * c.nctas + c.cta) * region;
  }
};

// elements of one round so that 2 halves x P sources x nctas regions fit the scratch area
template <typename C> __device__ __forceinline__ WireLayout wire_layout(const Ctx &c, size_t blk_elems_total) {
  WireLayout l;
  const size_t P = c.P(), n = static_cast<size_t>(c.nctas);
  size_t stripe = (blk_elems_total + n - 1) / n;
  stripe = (stripe + C::G - 1) / C::G * C::G;
  // shrink until it fits
  for (;;) {
    const WireGeom g = C::geom(stripe);
    if (2 * P * n * g.bytes <= c.w.scr_bytes || stripe <= static_cast<size_t>(C::G)) {
      l.region = g.bytes;
      break;
    }
    stripe = (stripe / 2 + C::G - 1) / C::G * C::G;
  }
  l.stripe_elems = stripe;
  l.blk_elems = stripe * n;
  return l;
}

// One round of the compressed exchange: elements [base, base + l.blk_elems) of every block.  Block q starts at
// q * stride elements of the buffer and has `blk_len(q)` valid elements (the last shard of an all-reduce is short).
//  do_rs: phase 1 + fold (blocks of src -> my shard), do_ag: the folded / given shard is broadcast compressed and
//  expanded into block q of dst.
template <typename T, typename C, typename Op>
__device__ __forceinline__ void wire_round(const Ctx &c, const WireLayout &l, const T *src, size_t src_stride, T *dst, size_t dst_stride,
                                           size_t blk, size_t total, size_t base, bool do_rs, bool do_ag) {
  constexpr int G = C::G;
  const uint32_t P = c.P(), me = c.r();
  const bool clip = do_rs && do_ag; // all-reduce: blocks are shards of one buffer of `total` elements
  const size_t s0 = base + static_cast<size_t>(c.cta) * l.stripe_elems; // my stripe starts here in every block
  // valid elements of my stripe in block q
  auto nel = [&](uint32_t q) -> size_t {
    size_t bl = blk;
    if (clip) bl = total > static_cast<size_t>(q) * blk ? (total - static_cast<size_t>(q) * blk < blk ? total - static_cast<size_t>(q) * blk : blk) : 0;
    size_t e1 = s0 + l.stripe_elems;
    if (e1 > base + l.blk_elems) e1 = base + l.blk_elems;
    if (e1 > bl) e1 = bl;
    return e1 > s0 ? e1 - s0 : 0;
  };
  const WireGeom g = C::geom(l.stripe_elems);
  const bool al = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (src_stride * sizeof(T)) | (dst_stride * sizeof(T)) |
                    (s0 * sizeof(T))) & 15) == 0;
  if (do_rs) {
    // ---- 1. my stripe of block q -> region [0][me] of rank q
    for (uint32_t k = 1; k < P; ++k) {
      const uint32_t q = (me + k) % P;
      const size_t ne = nel(q), groups = (ne + G - 1) / G;
      char *reg = l.at(c, c.g(q), 0, me);
      const T *from = src + static_cast<size_t>(q) * src_stride + s0;
      for (size_t gi = threadIdx.x; gi < groups; gi += blockDim.x) {
        float v[G];
        load_group<T, G>(from + gi * G, ne - gi * G, al, v);
        C::enc(reg, g, gi, v);
      }
    }
  }
  // the regions written next must be free (all-gather alone starts here: half 1 may still be read by a peer
  // expanding the previous call's shards) / the pushes above must have landed
  chan_sync(c, false, 0, 0, nullptr, nullptr);
  if (*c.err) return;
  // ---- 2. fold (own block exact, peers' from the wire) -> my shard; forward it compressed
  {
    const size_t ne = nel(me), groups = (ne + G - 1) / G;
    const T *own = do_rs ? src + static_cast<size_t>(me) * src_stride + s0 : src + s0;
    T *shard = do_ag ? dst + static_cast<size_t>(me) * dst_stride + s0 : dst + s0;
    for (size_t gi = threadIdx.x; gi < groups; gi += blockDim.x) {
      float acc[G];
      load_group<T, G>(own + gi * G, ne - gi * G, al, acc);
      if (do_rs) {
        // communicator-rank order, own contribution at its own position: the same sum whoever folds
        float sum[G];
        for (uint32_t q = 0; q < P; ++q) {
          float x[G];
          if (q == me) {
#pragma unroll
            for (int i = 0; i < G; ++i) x[i] = acc[i];
          } else {
            C::dec(l.at(c, c.w.rank, 0, q), g, gi, x);
          }
#pragma unroll
          for (int i = 0; i < G; ++i) sum[i] = q == 0 ? x[i] : Op::apply(sum[i], x[i]);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) acc[i] = sum[i];
      }
      if (do_rs || static_cast<const void *>(shard) != static_cast<const void *>(own)) store_group<T, G>(shard + gi * G, ne - gi * G, al, acc);
      if (do_ag)
        for (uint32_t k = 1; k < P; ++k) {
          const uint32_t q = (me + k) % P;
          C::enc(l.at(c, c.g(q), 1, me), g, gi, acc);
        }
    }
  }
  chan_sync(c, false, 0, 0, nullptr, nullptr); // shards have landed / half 0 may be rewritten by the next call
  if (!do_ag || *c.err) return;
  // ---- 3. expand the other ranks' shards
  for (uint32_t k = 1; k < P; ++k) {
    const uint32_t q = (me + P - k) % P;
    const size_t ne = nel(q), groups = (ne + G - 1) / G;
    const char *reg = l.at(c, c.w.rank, 1, q);
    T *to = dst + static_cast<size_t>(q) * dst_stride + s0;
    for (size_t gi = threadIdx.x; gi < groups; gi += blockDim.x) {
      float v[G];
      C::dec(reg, g, gi, v);
      store_group<T, G>(to + gi * G, ne - gi * G, al, v);
    }
  }
}

template <typename T, typename C, typename Op> __device__ __forceinline__ void wire_collective_t(const Ctx &c) {
  const WorkItem &it = c.it;
  const operation op = static_cast<operation>(it.desc.scenario);
  const uint32_t P = c.P();
  const T *src = reinterpret_cast<const T *>(c.heap(c.w.rank) + it.desc.addr0());
  T *dst = reinterpret_cast<T *>(c.heap(c.w.rank) + it.desc.addr2());
  const size_t count = it.desc.count;
  // block = what one rank owns: a shard of the all-reduce, the per-rank block of reduce_scatter / allgather
  const size_t blk = op == operation::allreduce ? (count + P - 1) / P : count;
  const WireLayout l = wire_layout<C>(c, blk);
  for (size_t base = 0; base < blk; base += l.blk_elems) {
    if (op == operation::allreduce) wire_round<T, C, Op>(c, l, src, blk, dst, blk, blk, count, base, true, true);
    else if (op == operation::reduce_scatter) wire_round<T, C, Op>(c, l, src, blk, dst, 0, blk, count, base, true, false);
    else wire_round<T, C, Op>(c, l, src, 0, dst, blk, blk, count, base, false, true);
    if (*c.err) return;
  }
}

template <typename T, typename Op> __device__ __forceinline__ void wire_collective_c(const Ctx &c) {
  const bool scaled = c.it.ratio_log == 5;
  switch (static_cast<dataType>(c.it.cdtype)) {
  case dataType::float16: wire_collective_t<T, Codec16<__half>, Op>(c); break;
  case dataType::bfloat16: wire_collective_t<T, Codec16<__nv_bfloat16>, Op>(c); break;
  case dataType::float8_e4m3:
    if (scaled) wire_collective_t<T, Codec8<__nv_fp8_e4m3, true>, Op>(c);
    else wire_collective_t<T, Codec8<__nv_fp8_e4m3, false>, Op>(c);
    break;
  case dataType::float8_e5m2:
    if (scaled) wire_collective_t<T, Codec8<__nv_fp8_e5m2, true>, Op>(c);
    else wire_collective_t<T, Codec8<__nv_fp8_e5m2, false>, Op>(c);
    break;
  default:
    if (threadIdx.x == 0) atomicOr(c.err, COMPRESSION_ERROR);
    break;
  }
}

__device__ __noinline__ void wire_collective(const Ctx &c) {
  const bool sum = c.it.desc.function == static_cast<uint32_t>(reduceFunction::SUM);
  switch (static_cast<dataType>(c.it.udtype)) {
  case dataType::float32:
    if (sum) wire_collective_c<float, OpSum>(c);
    else wire_collective_c<float, OpMax>(c);
    break;
  case dataType::bfloat16:
    if (sum) wire_collective_c<__nv_bfloat16, OpSum>(c);
    else wire_collective_c<__nv_bfloat16, OpMax>(c);
    break;
  case dataType::float16:
    if (sum) wire_collective_c<__half, OpSum>(c);
    else wire_collective_c<__half, OpMax>(c);
    break;
  default:
    if (threadIdx.x == 0) atomicOr(c.err, COMPRESSION_ERROR);
    break;
  }
}

} // namespace k
} // namespace cuda
} // namespace accl
