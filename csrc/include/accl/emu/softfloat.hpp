// CPU implementations of the narrow float formats and of the cast / reduce
// lanes of the datapath (fp16, bf16, fp8-e4m3, fp8-e5m2; block-scaled fp8).
//
// The reference gets these from Vitis `half`/`ap_fixed` headers
// (test/model/emulator/hls_sim_headers) and the two HLS plugins
// `hp_compression` (kernels/plugins/hp_compression/hp_compression.cpp:30-144)
// and `reduce_ops` (kernels/plugins/reduce_ops/reduce_ops.cpp:31-107).  Here
// they are dependency-free scalar routines with round-to-nearest-even, used
// by the emulator's data mover and by the tests as the numerical reference
// for the CUDA kernels.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>

#include "accl/constants.hpp"

namespace accl {
namespace emu {

inline uint32_t f32_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float bits_f32(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Generic small-float encoder: EXP exponent bits, MAN mantissa bits, IEEE-like
// (inf/nan) unless FN (finite-only "fn" formats such as e4m3: no inf, one NaN,
// max exponent usable for normals).
template <int EXP, int MAN, bool FN> struct SmallFloat {
  static constexpr int BIAS = (1 << (EXP - 1)) - 1;
  static constexpr int BITS = 1 + EXP + MAN;
  static constexpr uint32_t EXP_MASK = (1u << EXP) - 1;
  static constexpr uint32_t MAN_MASK = (1u << MAN) - 1;

  static float max_finite() {
    if (FN) return std::ldexp(1.0f + static_cast<float>(MAN_MASK - 1) / (1 << MAN), static_cast<int>(EXP_MASK) - BIAS);
    return std::ldexp(1.0f + static_cast<float>(MAN_MASK) / (1 << MAN), static_cast<int>(EXP_MASK) - 1 - BIAS);
  }

  static uint32_t encode(float f) {
    const uint32_t sign = (f32_bits(f) >> 31) << (BITS - 1);
    if (std::isnan(f)) return sign | (EXP_MASK << MAN) | (FN ? MAN_MASK : (1u << (MAN - 1)));
    float a = std::fabs(f);
    const float mx = max_finite();
    if (std::isinf(a)) return FN ? (sign | (EXP_MASK << MAN) | (MAN_MASK - 1)) : (sign | (EXP_MASK << MAN));
    if (a == 0.0f) return sign;
    int e;
    float m = std::frexp(a, &e); // a = m * 2^e, m in [0.5,1)
    int exp = e - 1;             // a = (2m) * 2^exp, 2m in [1,2)
    int q;                       // quantisation exponent: value = k * 2^q
    if (exp < 1 - BIAS) q = 1 - BIAS - MAN; // subnormal range
    else q = exp - MAN;
    float scaled = std::ldexp(a, -q);
    float r = std::nearbyint(scaled); // ties-to-even in default rounding mode
    float v = std::ldexp(r, q);
    if (v > mx) {
      if (FN) return sign | (EXP_MASK << MAN) | (MAN_MASK - 1); // saturate
      return sign | (EXP_MASK << MAN);                          // inf
    }
    if (v == 0.0f) return sign;
    m = std::frexp(v, &e);
    exp = e - 1;
    if (exp < 1 - BIAS) { // subnormal
      uint32_t man = static_cast<uint32_t>(std::ldexp(v, BIAS - 1 + MAN));
      return sign | man;
    }
    uint32_t man = static_cast<uint32_t>(std::ldexp(2.0f * m - 1.0f, MAN) + 0.5f) & MAN_MASK;
    return sign | (static_cast<uint32_t>(exp + BIAS) << MAN) | man;
  }

  static float decode(uint32_t u) {
    const bool neg = (u >> (BITS - 1)) & 1;
    const uint32_t e = (u >> MAN) & EXP_MASK;
    const uint32_t m = u & MAN_MASK;
    float v;
    if (e == 0) v = std::ldexp(static_cast<float>(m), 1 - BIAS - MAN);
    else if (e == EXP_MASK && !FN) v = m ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity();
    else if (e == EXP_MASK && FN && m == MAN_MASK) v = std::numeric_limits<float>::quiet_NaN();
    else v = std::ldexp(1.0f + static_cast<float>(m) / (1 << MAN), static_cast<int>(e) - BIAS);
    return neg ? -v : v;
  }
};

using F16 = SmallFloat<5, 10, false>;
using BF16 = SmallFloat<8, 7, false>;
using F8E4M3 = SmallFloat<4, 3, true>;
using F8E5M2 = SmallFloat<5, 2, false>;

inline bool is_float_type(dataType t) {
  return t == dataType::float16 || t == dataType::float32 || t == dataType::float64 || t == dataType::bfloat16 ||
         t == dataType::float8_e4m3 || t == dataType::float8_e5m2;
}
inline bool is_fp8(dataType t) { return t == dataType::float8_e4m3 || t == dataType::float8_e5m2; }

// element i of a typed array -> double (exact for all supported float types
// and for int32; int64 goes through load_i64)
inline double load_elem(const void *p, dataType t, size_t i) {
  const uint8_t *b = static_cast<const uint8_t *>(p);
  switch (t) {
  case dataType::float32: { float f; std::memcpy(&f, b + 4 * i, 4); return f; }
  case dataType::float64: { double d; std::memcpy(&d, b + 8 * i, 8); return d; }
  case dataType::float16: { uint16_t u; std::memcpy(&u, b + 2 * i, 2); return F16::decode(u); }
  case dataType::bfloat16: { uint16_t u; std::memcpy(&u, b + 2 * i, 2); return BF16::decode(u); }
  case dataType::float8_e4m3: return F8E4M3::decode(b[i]);
  case dataType::float8_e5m2: return F8E5M2::decode(b[i]);
  case dataType::int32: { int32_t v; std::memcpy(&v, b + 4 * i, 4); return v; }
  case dataType::int64: { int64_t v; std::memcpy(&v, b + 8 * i, 8); return static_cast<double>(v); }
  case dataType::int8: return static_cast<int8_t>(b[i]);
  default: throw std::invalid_argument("load_elem: unsupported dtype");
  }
}
inline void store_elem(void *p, dataType t, size_t i, double v) {
  uint8_t *b = static_cast<uint8_t *>(p);
  switch (t) {
  case dataType::float32: { float f = static_cast<float>(v); std::memcpy(b + 4 * i, &f, 4); return; }
  case dataType::float64: std::memcpy(b + 8 * i, &v, 8); return;
  case dataType::float16: { uint16_t u = static_cast<uint16_t>(F16::encode(static_cast<float>(v))); std::memcpy(b + 2 * i, &u, 2); return; }
  case dataType::bfloat16: { uint16_t u = static_cast<uint16_t>(BF16::encode(static_cast<float>(v))); std::memcpy(b + 2 * i, &u, 2); return; }
  case dataType::float8_e4m3: b[i] = static_cast<uint8_t>(F8E4M3::encode(static_cast<float>(v))); return;
  case dataType::float8_e5m2: b[i] = static_cast<uint8_t>(F8E5M2::encode(static_cast<float>(v))); return;
  case dataType::int32: { int32_t x = static_cast<int32_t>(v); std::memcpy(b + 4 * i, &x, 4); return; }
  case dataType::int64: { int64_t x = static_cast<int64_t>(v); std::memcpy(b + 8 * i, &x, 8); return; }
  case dataType::int8: b[i] = static_cast<uint8_t>(static_cast<int8_t>(v)); return;
  default: throw std::invalid_argument("store_elem: unsupported dtype");
  }
}

// Bytes occupied by n elements of representation `t`.  Block-scaled fp8
// carries one fp32 scale per 2^ratio_log elements after the payload.
inline size_t repr_bytes(dataType t, size_t n, uint32_t ratio_log) {
  size_t b = n * dtype_bytes(t);
  if (is_fp8(t) && ratio_log > 0) {
    const size_t blk = size_t(1) << ratio_log;
    b += 4 * ((n + blk - 1) / blk);
  }
  return b;
}

// Cast lane: n elements from (src, src_t) to (dst, dst_t).  When the
// destination is block-scaled fp8 each block is scaled so its absolute
// maximum maps to the format's largest finite value; when the source is, the
// stored scales are applied back.
inline void convert_buffer(const void *src, dataType src_t, void *dst, dataType dst_t, size_t n, uint32_t ratio_log) {
  if (src_t == dst_t) {
    std::memmove(dst, src, repr_bytes(src_t, n, ratio_log));
    return;
  }
  if (src_t == dataType::int64 || dst_t == dataType::int64) {
    if (src_t != dst_t) throw std::invalid_argument("convert_buffer: no cast lane for int64");
  }
  const bool src_scaled = is_fp8(src_t) && ratio_log > 0;
  const bool dst_scaled = is_fp8(dst_t) && ratio_log > 0;
  const size_t blk = ratio_log > 0 ? (size_t(1) << ratio_log) : n;
  const uint8_t *src_scales = static_cast<const uint8_t *>(src) + n * dtype_bytes(src_t);
  uint8_t *dst_scales = static_cast<uint8_t *>(dst) + n * dtype_bytes(dst_t);
  const float dst_max = dst_t == dataType::float8_e4m3 ? F8E4M3::max_finite() : F8E5M2::max_finite();
  for (size_t b0 = 0, bi = 0; b0 < n; b0 += blk, ++bi) {
    const size_t b1 = std::min(n, b0 + blk);
    float in_scale = 1.0f;
    if (src_scaled) std::memcpy(&in_scale, src_scales + 4 * bi, 4);
    float out_scale = 1.0f;
    if (dst_scaled) {
      double amax = 0;
      for (size_t i = b0; i < b1; ++i) amax = std::max(amax, std::fabs(load_elem(src, src_t, i) * in_scale));
      out_scale = amax > 0 ? static_cast<float>(amax / dst_max) : 1.0f;
      std::memcpy(dst_scales + 4 * bi, &out_scale, 4);
    }
    for (size_t i = b0; i < b1; ++i) store_elem(dst, dst_t, i, load_elem(src, src_t, i) * in_scale / out_scale);
  }
}

// Arithmetic lane: dst[i] = fn(a[i], b[i]) on n elements of type t (never a
// block-scaled representation: fp8 wires are decompressed first).
inline void reduce_buffer(const void *a, const void *b, void *dst, dataType t, size_t n, reduceFunction fn) {
  if (t == dataType::int64) {
    for (size_t i = 0; i < n; ++i) {
      int64_t x, y;
      std::memcpy(&x, static_cast<const uint8_t *>(a) + 8 * i, 8);
      std::memcpy(&y, static_cast<const uint8_t *>(b) + 8 * i, 8);
      int64_t r = fn == reduceFunction::SUM ? static_cast<int64_t>(static_cast<uint64_t>(x) + static_cast<uint64_t>(y)) : std::max(x, y);
      std::memcpy(static_cast<uint8_t *>(dst) + 8 * i, &r, 8);
    }
    return;
  }
  if (t == dataType::int32) {
    for (size_t i = 0; i < n; ++i) {
      int32_t x, y;
      std::memcpy(&x, static_cast<const uint8_t *>(a) + 4 * i, 4);
      std::memcpy(&y, static_cast<const uint8_t *>(b) + 4 * i, 4);
      int32_t r = fn == reduceFunction::SUM ? static_cast<int32_t>(static_cast<uint32_t>(x) + static_cast<uint32_t>(y)) : std::max(x, y);
      std::memcpy(static_cast<uint8_t *>(dst) + 4 * i, &r, 4);
    }
    return;
  }
  for (size_t i = 0; i < n; ++i) {
    double x = load_elem(a, t, i), y = load_elem(b, t, i);
    // float32 sums must round like a float add, not a double add
    double r;
    if (fn == reduceFunction::SUM) r = t == dataType::float64 ? x + y : static_cast<double>(static_cast<float>(x) + static_cast<float>(y));
    else r = std::max(x, y);
    store_elem(dst, t, i, r);
  }
}

} // namespace emu
} // namespace accl
