// Device-API plugins: compute kernels that talk to the collective machinery
// themselves (the reference's kernels/plugins/vadd_put pattern, plus the
// north-star tcgen05 GEMM whose epilogue feeds a reduce-scatter).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "accl/cuda/devtypes.hpp"

namespace accl {
namespace cuda {

class CudaDevice;

struct GemmRsArgs {
  const void *a;     // [M, K] bf16 row-major (this rank's K-slice of the activations)
  const void *w;     // [N, K] bf16 row-major (this rank's K-slice of the weight, nn.Linear layout)
  uint64_t out_off;  // heap offset of this rank's [M / P, N] output shard (bf16, or fp32 with out_f32)
  uint32_t m, n, k;
  uint32_t epoch;    // launch counter (same on every rank), starts at 1
  int variant = 0;   // 0: automatic, 1: one CTA per 128 x 256 tile, 2: CTA pair (cta_group::2) per 256 x 256 tile
  bool out_f32 = false; // shard is fp32: the P partial products are added without intermediate rounding
};

// C = A * W^T computed tile by tile on the 5th-gen tensor cores (TMA ->
// smem -> tcgen05.mma -> TMEM); every finished accumulator tile is converted
// to bf16, staged in shared memory and added straight into its owner rank's
// output shard over NVLink by the TMA unit (cp.reduce.async.bulk.tensor .add on
// the peer-mapped heap): reduce-scatter along M with no intermediate buffer
// and no separate collective.  Returns after enqueueing.
cudaError_t launch_gemm_rs(CudaDevice &dev, const GemmRsArgs &args, cudaStream_t stream);

// vector add whose result feeds an all-reduce issued by the kernel itself through the device API (engine must be
// enabled): out = allreduce_sum(x + y).  The vector is produced in chunks of `chunk_elems` (0: 1 Mi elements); each
// finished chunk is handed to the engine at once (all_reduce_async) while the kernel computes the next one.
cudaError_t launch_vadd_allreduce(CudaDevice &dev, uint64_t x_off, uint64_t y_off, uint64_t tmp_off, uint64_t out_off,
                                  uint32_t count, uint32_t chunk_elems, uint32_t comm_adr, uint32_t dpcfg_adr, uint32_t *status_dev,
                                  cudaStream_t stream);

// the reference's vadd_put: src + 1 is pushed tile by tile into stream `stream_id` of rank `dst_rank` while computing
cudaError_t launch_vadd_put(CudaDevice &dev, uint64_t src_off, uint32_t count, uint32_t dst_rank, uint32_t stream_id,
                            uint32_t *status_dev, cudaStream_t stream);
// consumer of a stream: `count` fp32 of stream `stream_id` -> heap offset dst_off
cudaError_t launch_stream_pull(CudaDevice &dev, uint64_t dst_off, uint32_t count, uint32_t stream_id, uint32_t *status_dev,
                               cudaStream_t stream);

// user kernel in the stream path: pulls `count` fp32 from this rank's stream FIFO, optionally adds one,
// pushes them back (reference kernels/plugins/loopback).  scratch_off: heap scratch of count * 4 bytes.
cudaError_t launch_loopback(CudaDevice &dev, uint64_t scratch_off, uint32_t count, bool add_one, uint32_t *status_dev,
                            cudaStream_t stream);

} // namespace cuda
} // namespace accl
