#include "accl/arithconfig.hpp"

namespace accl {

ArithConfig::ArithConfig(dataType u, dataType c, uint32_t ratio_log, bool arith_compressed)
    : uncompressed_dtype(u), compressed_dtype(c), uncompressed_elem_bytes(dtype_bytes(u)),
      compressed_elem_bytes(dtype_bytes(c)), elem_ratio_log(ratio_log), compressor_lane(cast_lane_id(u, c)),
      decompressor_lane(cast_lane_id(c, u)), arith_is_compressed(arith_compressed) {
  const dataType arith_t = arith_compressed ? c : u;
  arith_fn = {arith_fn_id(reduceFunction::SUM, arith_t), arith_fn_id(reduceFunction::MAX, arith_t)};
}

const arithConfigMap &default_arith_config() {
  static const arithConfigMap m = [] {
    arithConfigMap t;
    auto add = [&](dataType u, dataType c, uint32_t ratio_log, bool arith_c) {
      t.emplace(arithConfigKey{u, c}, ArithConfig(u, c, ratio_log, arith_c));
    };
    using D = dataType;
    add(D::float16, D::float16, 0, false);
    add(D::float32, D::float16, 0, true); // reduce in fp16, as the reference's default table does
    add(D::float32, D::float32, 0, false);
    add(D::float64, D::float64, 0, false);
    add(D::int32, D::int32, 0, false);
    add(D::int64, D::int64, 0, false);
    // B200 additions
    add(D::bfloat16, D::bfloat16, 0, false);
    add(D::float32, D::bfloat16, 0, false); // bf16 on the wire, fp32 accumulate
    // block-scaled fp8 wire formats: one fp32 scale per 32 elements
    add(D::float32, D::float8_e4m3, 5, false);
    add(D::float32, D::float8_e5m2, 5, false);
    add(D::bfloat16, D::float8_e4m3, 5, false);
    add(D::float16, D::float8_e4m3, 5, false);
    return t;
  }();
  return m;
}

void serialize_arithconfig(const ArithConfig &c, uint32_t out[ARITHCFG_WORDS]) {
  // dtype codes ride in the upper half of the byte-width words so the engine
  // can pick typed kernels without a second table
  out[0] = c.uncompressed_elem_bytes | (static_cast<uint32_t>(c.uncompressed_dtype) << 16);
  out[1] = c.compressed_elem_bytes | (static_cast<uint32_t>(c.compressed_dtype) << 16);
  out[2] = c.elem_ratio_log;
  out[3] = c.compressor_lane;
  out[4] = c.decompressor_lane;
  out[5] = c.arith_is_compressed ? 1u : 0u;
  out[6] = c.arith_fn.size() > 0 ? c.arith_fn[0] : 0;
  out[7] = c.arith_fn.size() > 1 ? c.arith_fn[1] : 0;
}

} // namespace accl
