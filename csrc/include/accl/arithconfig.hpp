// Datapath ("arithmetic") configurations: for every (buffer dtype, wire dtype)
// pair the engine may be asked to handle, how wide the elements are on each
// side, which cast kernels convert between them, whether the reduction runs
// on the compressed or the uncompressed representation, and which reduce
// functor ids implement SUM and MAX.
//
// Same role and field meaning as the reference's ArithConfig and
// DEFAULT_ARITH_CONFIG (driver/xrt/include/accl/arithconfig.hpp:32-119); the
// B200 table adds bfloat16 and block-scaled fp8 wire formats
// (`elem_ratio_log` > 0 == one scale per 2^elem_ratio_log elements, as the
// reference's docs anticipate: driver/xrt/docs/getting_started/compression.rst).
#pragma once
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

#include "accl/constants.hpp"

namespace accl {

// reduce functor ids: function index * 16 + dtype code. Shared by the
// emulator's arithmetic unit and the CUDA reduce templates.
constexpr uint32_t arith_fn_id(reduceFunction f, dataType t) {
  return static_cast<uint32_t>(f) * 16u + static_cast<uint32_t>(t);
}
// cast lane ids: (from dtype << 4) | to dtype
constexpr uint32_t cast_lane_id(dataType from, dataType to) {
  return (static_cast<uint32_t>(from) << 4) | static_cast<uint32_t>(to);
}

struct ArithConfig {
  dataType uncompressed_dtype = dataType::none;
  dataType compressed_dtype = dataType::none;
  uint32_t uncompressed_elem_bytes = 0;
  uint32_t compressed_elem_bytes = 0;
  // log2(elements per scale block); 0 = plain per-element cast
  uint32_t elem_ratio_log = 0;
  uint32_t compressor_lane = 0;   // cast_lane_id(uncompressed -> compressed)
  uint32_t decompressor_lane = 0; // cast_lane_id(compressed -> uncompressed)
  bool arith_is_compressed = false; // reduce in the compressed representation
  std::vector<uint32_t> arith_fn; // [SUM id, MAX id]

  // filled in when the table is written to the device (index into the
  // device-resident arith table; the reference stores a byte address)
  addr_t exchmem_addr = 0;

  ArithConfig() = default;
  ArithConfig(dataType u, dataType c, uint32_t ratio_log, bool arith_compressed);
};

using arithConfigKey = std::pair<dataType, dataType>; // {uncompressed, compressed}
using arithConfigMap = std::map<arithConfigKey, ArithConfig>;

// f16/f16, f32/f16 (arith on compressed, like the reference), f32/f32,
// f64/f64, i32/i32, i64/i64 + bf16/bf16, f32/bf16, f32/fp8 (block-scaled),
// bf16/fp8, f16/fp8.
const arithConfigMap &default_arith_config();

// flat serialisation used by both backends: 8 words per entry
constexpr int ARITHCFG_WORDS = 8;
void serialize_arithconfig(const ArithConfig &c, uint32_t out[ARITHCFG_WORDS]);

} // namespace accl
