// GEMM -> reduce-scatter plugin (BASELINE config #5, the north-star fused path).
//
// Each rank holds a K-slice of a row-parallel linear layer: A_r [M, K_r] and
// W_r [N, K_r] (bf16, K contiguous).  The full product is C = sum_r A_r W_r^T,
// reduce-scattered along M: rank o ends up with rows [o*M/P, (o+1)*M/P).
//
// One persistent, warp-specialised sm_100a kernel per rank, in two flavours selected at run time
// (GemmRsArgs::variant, default: the CTA pair when the shape allows it):
//   * single CTA per 128 x 256 tile (tcgen05.mma cta_group::1, 4-stage ring), and
//   * CTA PAIR per 256 x 256 tile (cluster of 2, tcgen05.mma cta_group::2 with M = 256: each CTA loads 128 rows of A
//     and 128 rows of W per stage — 32 KB instead of 48 KB, so 6 stages fit and W is fetched once per 256 rows),
// each with a bf16 or an fp32 reduction of the partial tiles (the TMA unit adds in the shard's element type: fp32
// shards accumulate the P partial products without intermediate rounding).
// Roles of the 8 warps of a CTA:
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of 128x64 (A) and
//               256x64 (W) bf16 tiles, 128B-swizzled, into a 4-stage smem ring,
//               signalling mbarriers with complete_tx
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma
//               (cta_group::1, kind::f16, M=128 N=256 K=16) from smem
//               descriptors into a TMEM accumulator (2 x 256 columns, double
//               buffered); tcgen05.commit releases smem stages and publishes
//               finished accumulators
//   warp 2      allocates / frees TMEM
//   warps 4-7   epilogue: tcgen05.ld the fp32 accumulator (32 lanes x 32
//               columns per instruction), convert to bf16, stage 32x64 boxes in
//               128B-swizzled smem and hand them to the TMA unit as
//               cp.reduce.async.bulk.tensor ... .add (SASS UTMAREDG.2D.ADD) whose
//               tensor map points at the OWNER rank's output shard in the
//               peer-mapped symmetric heap — the tile goes into the collective
//               as it leaves the tensor core, coalesced by hardware, with no
//               SM store instructions on the NVLink path
// Tiles are visited owner-rotated (peers' rows first, own rows last) so the
// NVLink traffic overlaps the remaining math.  Shards are zeroed in-kernel and
// two flag barriers over the sync pads (channel MAX_CH-1) bracket the adds.
//
// The reference has no GEMM; its mechanism for compute-initiated communication
// is kernels/plugins/vadd_put/vadd_put.cpp (data.push + stream_put from inside
// the compute kernel).  Baseline to beat: cuBLAS GEMM + NCCL reduce_scatter.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "accl/cuda/cudadevice.hpp"
#include "accl/cuda/driver_api.hpp"
#include "accl/cuda/plugins.hpp"
#include "accl/device/api.cuh"
#include "kernels.cuh"

namespace accl {
namespace cuda {

namespace g {
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4, ACC_STAGES = 2;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_COLS = 64;                        // columns per TMA reduce box (128 B of bf16: one swizzle row)
constexpr int EPI_BUF_BYTES = 32 * EPI_COLS * 2;    // 32 rows x 64 cols bf16 = 4 KB
constexpr int EPI_BYTES = 4 /*warps*/ * 2 /*double buffer*/ * EPI_BUF_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (sm_100 format, version 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);      // start address >> 4
  d |= static_cast<uint64_t>(0) << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
  d |= static_cast<uint64_t>(1) << 46;                          // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                          // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N x M
__device__ __forceinline__ uint32_t make_idesc(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// smem tile -> global memory of the tensor map's rank, element-wise ADD performed by the TMA unit
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, const void *smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(uint32_t lo_f32, uint32_t hi_f32) {
  __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(lo_f32), __uint_as_float(hi_f32));
  return *reinterpret_cast<uint32_t *>(&v);
}
} // namespace g

struct alignas(64) OutMaps {
  CUtensorMap m[ACCL_MAX_RANKS]; // owner rank o: its [M/P, N] shard through my peer mapping
};

struct GemmRsParams {
  DevWorld w;
  uint64_t out_off;
  uint32_t m, n, k;
  uint32_t epoch;
  uint32_t timeout_us;
  unsigned int *grid_flags; // zeroed before every launch: [0] zero-phase arrivals, [1] release, [2] finished CTAs, [3] error, [8..] peer offsets
};

namespace g {
// ---- CTA-pair (cta_group::2) flavours of the primitives
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same object in the pair's leader (cluster rank 0)
__device__ __forceinline__ uint32_t leader_addr(const void *p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
  return r;
}
// issued by both CTAs; the transaction bytes are accounted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}

template <bool PAIR> struct Cfg {
  static constexpr int STAGES = PAIR ? 6 : 4;
  static constexpr int BN_LOAD = PAIR ? BN / 2 : BN;          // W rows this CTA loads per stage
  static constexpr int B_STAGE = BN_LOAD * BK * 2;
  static constexpr int STAGE = A_STAGE_BYTES + B_STAGE;
  static constexpr int SMEM = STAGES * STAGE + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TILE_M = PAIR ? 2 * BM : BM;           // rows of one output tile
};
} // namespace g

// Tile order: groups of GROUP_M tile rows are swept column by column, so the ~(SMs / CTAs per tile) tiles in flight
// form a near-square GROUP_M x ~9 patch whose A rows and W rows (K-long, 4 MB each at K = 8192) stay in the 126 MB L2.
// A plain row-major sweep keeps all of W in flight at once: measured 2.2 GB of DRAM reads for an 8192^3 GEMM whose
// operands are 256 MB (ncu, profiles/ncu_gemm_rs_pair.md) and a tensor pipe waiting on L2 misses.
template <int GROUP_M>
__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t tiles_m, uint32_t tiles_n, uint32_t &mo, uint32_t &nb) {
  const uint32_t width = GROUP_M * tiles_n;
  const uint32_t group = t / width, first_m = group * GROUP_M;
  const uint32_t gsize = tiles_m - first_m < static_cast<uint32_t>(GROUP_M) ? tiles_m - first_m : static_cast<uint32_t>(GROUP_M);
  const uint32_t in = t % width;
  mo = first_m + in % gsize;
  nb = in / gsize;
}

// One epilogue chunk: 32 accumulator rows (this warp's TMEM lanes) x COLS columns -> staged in 128B-swizzled
// shared memory -> handed to the collective (device API: reduce_scatter_emit_tile).
template <bool F32>
__device__ __forceinline__ void epilogue_chunk(uint32_t taddr, uint8_t *buf, int lane) {
  using namespace g;
  uint32_t r[32];
  if (F32) {
    // 32 fp32 columns = one 128-byte row per lane
    tmem_ld_32x32(taddr, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t chunk = static_cast<uint32_t>(j) ^ (lane & 7u); // XOR swizzle of the 16-byte chunk with the row
      *reinterpret_cast<uint4 *>(buf + lane * 128 + chunk * 16) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tmem_ld_32x32(taddr + h * 32, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t chunk = static_cast<uint32_t>(h * 4 + j) ^ (lane & 7u);
        uint4 v;
        v.x = pack_bf16x2(r[8 * j + 0], r[8 * j + 1]);
        v.y = pack_bf16x2(r[8 * j + 2], r[8 * j + 3]);
        v.z = pack_bf16x2(r[8 * j + 4], r[8 * j + 5]);
        v.w = pack_bf16x2(r[8 * j + 6], r[8 * j + 7]);
        *reinterpret_cast<uint4 *>(buf + lane * 128 + chunk * 16) = v;
      }
    }
  }
}

template <bool PAIR, bool F32>
__global__ void __launch_bounds__(g::THREADS, 1)
k_plugin_gemm_rs(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ OutMaps out_maps, GemmRsParams p) {
  using namespace g;
  using C = Cfg<PAIR>;
  constexpr int STAGES = C::STAGES;
  constexpr int COLS = F32 ? 32 : 64; // columns per emitted box: one 128-byte swizzle row
  constexpr int GROUP_M = PAIR ? 8 : 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t *smem_epi = smem + STAGES * C::STAGE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * C::STAGE + EPI_BYTES);
  uint64_t *full = bars, *empty = bars + STAGES, *tmem_full = bars + 2 * STAGES, *tmem_empty = bars + 2 * STAGES + ACC_STAGES;
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 2 * ACC_STAGES);

  __shared__ uint32_t s_err;
  __shared__ k::PtrTable s_tab;
  __shared__ WorkItem s_item; // identity communicator for the sync pads
  __shared__ uint64_t s_off0[ACCL_MAX_RANKS], s_off2[ACCL_MAX_RANKS];

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t cr = PAIR ? cluster_ctarank() : 0; // 0 = leader of the pair
  const uint32_t unit = PAIR ? blockIdx.x / 2 : blockIdx.x, num_units = PAIR ? gridDim.x / 2 : gridDim.x;
  const uint32_t P = p.w.world, me = p.w.rank;
  const uint32_t tiles_m = p.m / C::TILE_M, tiles_n = p.n / BN, num_tiles = tiles_m * tiles_n;
  const uint32_t k_blocks = p.k / BK;
  const uint32_t rows_per_rank = p.m / P, tiles_m_per_rank = tiles_m / P;
  char *my_heap = p.w.window + static_cast<uint64_t>(me) * p.w.heap_bytes;

  // ---------------- one-time setup
  if (threadIdx.x == 0) {
    s_err = 0;
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);  // the producer's arrive.expect_tx (pair: the leader's; the peer's copy is never used)
      mbar_init(&empty[i], 1); // tcgen05.commit (pair: the leader's multicast commit)
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PAIR ? 8 : 4); // one arrival per epilogue warp (pair: of both CTAs, on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    s_item.comm_size = P;
    s_item.comm_rank = me;
    for (uint32_t r = 0; r < static_cast<uint32_t>(ACCL_MAX_RANKS); ++r) s_item.members[r] = static_cast<uint8_t>(r < P ? r : 0);
    s_item.desc.scenario = 0x47; // 'G': both ends must be in the same plugin
    s_item.timeout_us = p.timeout_us;
    s_item.bank = 0;
    s_item.comm_sig = 0x47454Du;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync(); // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // ---------------- zero my output shard and meet the peers — WITHOUT holding back the math: the TMA producer
  // (warp 0) and the MMA issuer (warp 1) go straight to the main loop; warps 2-7 zero the shard (the reduce-adds
  // accumulate into it), warp 3 of CTA 0 then exchanges the shard offsets with the peers on the sync pads and
  // releases the epilogues of the whole grid.  The first tile's accumulator is ready long after (K / 64 x 512 cycles).
  k::Ctx ctx{p.w, s_item, MAX_CH - 1, 1, reinterpret_cast<Ctrl *>(my_heap), &s_err, static_cast<uint64_t>(p.timeout_us) * 1000ull, &s_tab};
  if (warp >= 2) {
    dev::Vec16 z{0, 0, 0, 0};
    char *shard = my_heap + p.out_off;
    const size_t nvec = static_cast<size_t>(rows_per_rank) * p.n * (F32 ? 4 : 2) / 16;
    const size_t zt = static_cast<size_t>(blockIdx.x) * 192 + (threadIdx.x - 64), zstride = static_cast<size_t>(gridDim.x) * 192;
    for (size_t i = zt; i < nvec; i += zstride) dev::st_stream(shard + i * 16, z);
    __threadfence();
    asm volatile("bar.sync 1, 192;" ::: "memory"); // warps 2-7 only
    if (threadIdx.x == 64) atomicAdd(&p.grid_flags[0], 1u);
    if (blockIdx.x == 0 && warp == 3) {
      // the whole grid has zeroed -> meet every peer (sync pads, channel MAX_CH-1), one lane per peer
      if (lane == 0)
        while (atomicAdd(&p.grid_flags[0], 0u) < gridDim.x) dev::nanosleep(100);
      __syncwarp();
      __threadfence();
      uint64_t peer_off = p.out_off;
      if (static_cast<uint32_t>(lane) < P && static_cast<uint32_t>(lane) != me) {
        PadBank &mine = ctx.pads();
        PadBank &theirs = ctx.pads_of(lane);
        const uint32_t ch = MAX_CH - 1;
        const uint32_t v = mine.sent[ch][lane] + 1;
        mine.sent[ch][lane] = v;
        SyncRec *rr = &theirs.rec[ch][me];
        dev::st_relaxed_sys(&rr->off0, p.out_off);
        dev::st_relaxed_sys(&rr->off2, p.out_off);
        dev::st_relaxed_sys(&rr->kind, ctx.kind_word());
        dev::st_release_sys(&theirs.sig[ch][me], v);
        const uint32_t e = mine.expect[ch][lane] + 1;
        mine.expect[ch][lane] = e;
        if (k::wait_ge(&mine.sig[ch][lane], e, ctx, RECEIVE_TIMEOUT_ERROR)) {
          peer_off = dev::ld_relaxed_sys(&mine.rec[ch][lane].off0);
          if (dev::ld_relaxed_sys(&mine.rec[ch][lane].kind) != ctx.kind_word()) atomicOr(&s_err, PACK_SEQ_NUMBER_ERROR);
        }
      }
      if (static_cast<uint32_t>(lane) < P) reinterpret_cast<volatile uint64_t *>(p.grid_flags + 8)[lane] = peer_off;
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicExch(&p.grid_flags[1], 1u); // release the epilogues of every CTA
    }
  }

  // ---------------- warp-specialised main loop
  if (warp == 0) {
    // ===== TMA producer (pair: both CTAs, each its 128 rows of A and its 128 rows of W)
    uint32_t stage = 0, phase = 0;
    for (uint32_t t = unit; t < num_tiles; t += num_units) {
      uint32_t mo, nb;
      tile_coords<GROUP_M>(t, tiles_m, tiles_n, mo, nb);
      const uint32_t mb = (mo + (me + 1) * tiles_m_per_rank) % tiles_m; // peers' rows first, mine last
      for (uint32_t kb = 0; kb < k_blocks; ++kb) {
        if (lane == 0) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (PAIR) {
            if (cr == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE); // both CTAs' loads land on the leader's barrier
            tma_load_2d_pair(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full[stage], static_cast<int>(kb * BK),
                             static_cast<int>(mb * C::TILE_M + cr * BM));
            tma_load_2d_pair(smem_b + stage * C::B_STAGE, &tmap_b, &full[stage], static_cast<int>(kb * BK),
                             static_cast<int>(nb * BN + cr * C::BN_LOAD));
          } else {
            mbar_expect_tx(&full[stage], C::STAGE);
            tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full[stage], static_cast<int>(kb * BK), static_cast<int>(mb * BM));
            tma_load_2d(smem_b + stage * C::B_STAGE, &tmap_b, &full[stage], static_cast<int>(kb * BK), static_cast<int>(nb * BN));
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && cr == 0) {
    // ===== MMA issuer (pair: the leader only)
    const uint32_t idesc = make_idesc(C::TILE_M, BN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (uint32_t t = unit; t < num_tiles; t += num_units) {
      if (lane == 0) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1); // epilogue(s) have drained this accumulator
        tc_fence_after();
      }
      __syncwarp();
      for (uint32_t kb = 0; kb < k_blocks; ++kb) {
        if (lane == 0) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * C::B_STAGE));
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) { // +32 bytes (>>4 = 2) per K=16 step inside the swizzle atom
            if (PAIR) umma_f16_pair(tmem_base + acc * BN, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
            else umma_f16(tmem_base + acc * BN, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) ? 1u : 0u);
          }
          if (PAIR) {
            umma_commit_pair(&empty[stage]); // smem stage reusable (in both CTAs) once these MMAs retire
            if (kb == k_blocks - 1) umma_commit_pair(&tmem_full[acc]);
          } else {
            umma_commit(&empty[stage]);
            if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> shard element type -> swizzled smem -> emitted into the collective
    const uint32_t ew = warp - 4; // == warp % 4: this warp owns TMEM lanes [32*ew, 32*ew+32)
    if (lane == 0) {
      uint32_t spins = 0;
      while (*reinterpret_cast<volatile unsigned int *>(&p.grid_flags[1]) == 0)
        if (++spins > 64) dev::nanosleep(200);
    }
    __syncwarp();
    __threadfence();
    // shards must sit at the same heap offset on every rank (the output tensor maps assume it)
    if (lane < P && reinterpret_cast<const volatile uint64_t *>(p.grid_flags + 8)[lane] != p.out_off) atomicOr(&p.grid_flags[3], static_cast<unsigned>(DMA_MISMATCH_ERROR));
    uint8_t *my_epi = smem_epi + ew * 2 * EPI_BUF_BYTES;
    uint32_t acc = 0, acc_phase = 0, ebuf = 0;
    for (uint32_t t = unit; t < num_tiles; t += num_units) {
      uint32_t mo, nb;
      tile_coords<GROUP_M>(t, tiles_m, tiles_n, mo, nb);
      const uint32_t mb = (mo + (me + 1) * tiles_m_per_rank) % tiles_m;
      const uint32_t row0 = mb * C::TILE_M + cr * BM + ew * 32; // first of this warp's 32 rows
      const uint32_t owner = row0 / rows_per_rank;               // a warp's rows never straddle owners
      const int local_row0 = static_cast<int>(row0 - owner * rows_per_rank);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / COLS; ++c) {
        uint8_t *buf = my_epi + ebuf * EPI_BUF_BYTES;
        // the TMA unit must be done READING this buffer (issued two chunks ago)
        if (lane == 0) device::emit_wait_read<1>();
        __syncwarp();
        epilogue_chunk<F32>(tmem_base + ((ew * 32u) << 16) + acc * BN + c * COLS, buf, lane);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the TMA unit
        __syncwarp();
        if (lane == 0) device::reduce_scatter_emit_tile(&out_maps.m[owner], buf, static_cast<int>(nb * BN + c * COLS), local_row0);
        ebuf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tmem_empty[acc]);
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) device::emit_flush(); // all my reductions have completed
    __syncwarp();
    __threadfence_system(); // my adds are performed before this CTA reports completion
  }

  // ---------------- teardown: free TMEM; the last CTA of the grid meets the peers again
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync(); // nobody frees TMEM or exits while the peer may still signal its barriers
  if (warp == 2) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(&p.grid_flags[2], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    // every CTA of this rank has pushed its tiles: when all ranks meet here, every shard is complete
    k::chan_sync(ctx, false, 0, 0, nullptr, nullptr);
    if (threadIdx.x == 0 && s_err) atomicOr(&p.grid_flags[3], s_err);
  }
}

// ----------------------------------------------------------------------- host
namespace {
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled unavailable");
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

CUtensorMap make_map(const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols, bool f32 = false) {
  CUtensorMap m;
  const uint64_t es = f32 ? 4 : 2;
  const cuuint64_t dims[2] = {cols, rows};           // innermost first
  const cuuint64_t strides[1] = {cols * es};         // bytes between rows
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed: " + cu_error_string(r));
  return m;
}

struct PluginState {
  unsigned int *flags = nullptr; // 64 words
};
std::mutex g_state_m;
std::map<CudaDevice *, PluginState> g_state;

template <bool PAIR, bool F32> void set_smem_attr() {
  cudaFuncSetAttribute(k_plugin_gemm_rs<PAIR, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, g::Cfg<PAIR>::SMEM);
}
} // namespace

void preload_gemm_rs_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_plugin_gemm_rs<false, false>);
  cudaFuncGetAttributes(&a, k_plugin_gemm_rs<false, true>);
  cudaFuncGetAttributes(&a, k_plugin_gemm_rs<true, false>);
  cudaFuncGetAttributes(&a, k_plugin_gemm_rs<true, true>);
  set_smem_attr<false, false>();
  set_smem_attr<false, true>();
  set_smem_attr<true, false>();
  set_smem_attr<true, true>();
}

template <bool PAIR, bool F32>
static cudaError_t launch_variant(const CUtensorMap &ta, const CUtensorMap &tb, const OutMaps &om, const GemmRsParams &p, uint32_t tiles,
                                  int sms, cudaStream_t stream) {
  using namespace g;
  set_smem_attr<PAIR, F32>();
  cudaLaunchConfig_t lc = {};
  lc.blockDim = dim3(THREADS);
  lc.dynamicSmemBytes = Cfg<PAIR>::SMEM;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  if (PAIR) {
    uint32_t pairs = std::min<uint32_t>(static_cast<uint32_t>(sms) / 2, tiles);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    lc.gridDim = dim3(2 * pairs);
    // the kernel spins on grid-wide flags: every cluster must be resident at once
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, k_plugin_gemm_rs<PAIR, F32>, &lc) == cudaSuccess && max_clusters > 0)
      pairs = std::min<uint32_t>(pairs, static_cast<uint32_t>(max_clusters));
    lc.gridDim = dim3(2 * pairs);
  } else {
    lc.gridDim = dim3(std::min<uint32_t>(static_cast<uint32_t>(sms), tiles));
  }
  return cudaLaunchKernelEx(&lc, k_plugin_gemm_rs<PAIR, F32>, ta, tb, om, p);
}

cudaError_t launch_gemm_rs(CudaDevice &dev, const GemmRsArgs &a, cudaStream_t stream) {
  using namespace g;
  const uint32_t P = dev.world().world;
  if (a.m % (BM * P) || a.n % BN || a.k % BK)
    throw std::invalid_argument("gemm_rs: need M % (128*world) == 0, N % 256 == 0, K % 64 == 0");
  ACCL_CUDART(cudaSetDevice(dev.device()));
  PluginState *st;
  {
    std::lock_guard<std::mutex> lk(g_state_m);
    st = &g_state[&dev];
    if (!st->flags) {
      ACCL_CUDART(cudaMalloc(&st->flags, 64 * sizeof(unsigned int)));
      ACCL_CUDART(cudaMemset(st->flags, 0, 64 * sizeof(unsigned int)));
    }
  }
  // CTA-pair kernel: 256-row tiles must not straddle shard owners.  variant: 0 = automatic, 1 = single CTA, 2 = pair;
  // ACCL_GEMM_VARIANT overrides the automatic choice.
  static const int env_variant = [] {
    const char *e = std::getenv("ACCL_GEMM_VARIANT");
    return e ? std::atoi(e) : 0;
  }();
  int variant = a.variant ? a.variant : env_variant;
  const bool pair_ok = a.m % (2 * BM * P) == 0;
  if (variant == 2 && !pair_ok) throw std::invalid_argument("gemm_rs: the CTA-pair kernel needs M % (256*world) == 0");
  const bool pair = variant == 2 || (variant == 0 && pair_ok);
  const bool f32 = a.out_f32;
  const CUtensorMap ta = make_map(a.a, a.m, a.k, BM, BK);
  const CUtensorMap tb = make_map(a.w, a.n, a.k, pair ? BN / 2 : BN, BK); // each CTA of a pair loads half of the N tile
  OutMaps om;
  std::memset(&om, 0, sizeof(om));
  for (uint32_t o = 0; o < P; ++o) // symmetric allocation: the shard sits at out_off in every heap (verified in-kernel)
    om.m[o] = make_map(dev.world().window + static_cast<uint64_t>(o) * dev.world().heap_bytes + a.out_off, a.m / P, a.n, 32,
                       f32 ? 32 : EPI_COLS, f32);
  GemmRsParams p;
  p.w = dev.world();
  p.out_off = a.out_off;
  p.m = a.m;
  p.n = a.n;
  p.k = a.k;
  p.epoch = a.epoch;
  p.timeout_us = dev.timeout_us();
  p.grid_flags = st->flags;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev.device());
  ACCL_CUDART(cudaMemsetAsync(st->flags, 0, 64 * sizeof(unsigned int), stream));
  const uint32_t tiles = (a.m / (pair ? 2 * BM : BM)) * (a.n / BN);
  if (pair) return f32 ? launch_variant<true, true>(ta, tb, om, p, tiles, sms, stream) : launch_variant<true, false>(ta, tb, om, p, tiles, sms, stream);
  return f32 ? launch_variant<false, true>(ta, tb, om, p, tiles, sms, stream) : launch_variant<false, false>(ta, tb, om, p, tiles, sms, stream);
}

} // namespace cuda
} // namespace accl
