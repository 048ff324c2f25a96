// accl_probe: bring-up probe for a B200 NVSwitch node.
//
// Launch one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK (and
// MASTER_ADDR / MASTER_PORT) set, e.g.
//   for r in 0 1; do RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r ./accl_probe & done; wait
// It builds the symmetric heap (VMM + fd passing + NVLS multicast), checks
// peer stores/flags, multimem load-reduce/store, stream memory operations and
// a persistent polling kernel, and prints raw link bandwidths for different
// CTA counts.  The reference's equivalent is the hardware bring-up half of
// accl_network_utils (configure_* + a first nop call).
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "accl/bootstrap.hpp"
#include "accl/common.hpp"
#include "accl/cuda/driver_api.hpp"
#include "accl/cuda/symheap.hpp"
#include "accl/device/primitives.cuh"

using namespace accl;
using namespace accl::dev;

#define CK(x) ACCL_CUDART(x)

__global__ void k_fill(float *p, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// every rank: write pattern into next rank's heap, then raise its flag
__global__ void k_p2p_put(float *peer_buf, uint32_t *peer_flag, size_t n, float v, uint32_t seq) {
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) peer_buf[i] = v;
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(peer_flag, seq);
}
__global__ void k_p2p_wait_check(const float *buf, const uint32_t *flag, size_t n, float expect, uint32_t seq, int *bad) {
  if (threadIdx.x == 0)
    while (ld_acquire_sys(flag) != seq) nanosleep(64);
  __syncthreads();
  for (size_t i = threadIdx.x; i < n; i += blockDim.x)
    if (reinterpret_cast<const volatile float *>(buf)[i] != expect) atomicAdd(bad, 1);
}

__global__ void k_barrier_mc(uint32_t *mc_flag, const uint32_t *my_flag, uint32_t target) {
  if (threadIdx.x == 0) {
    multimem_red_release_add(mc_flag, 1);
    while (ld_acquire_sys(my_flag) < target) nanosleep(32);
  }
}
__global__ void k_barrier_p2p(uint32_t **peer_flags, const uint32_t *my_flag, int world, uint32_t target) {
  if ((int)threadIdx.x < world) red_release_sys_add(peer_flags[threadIdx.x], 1);
  if (threadIdx.x == 0)
    while (ld_acquire_sys(my_flag) < target) nanosleep(32);
}

// bandwidth kernels, 16 B per thread per step, unroll 4
template <int MODE>
__global__ void __launch_bounds__(512) k_bw(const char *src, char *dst, size_t bytes) {
  const size_t nvec = bytes / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    Vec16 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const char *s = src + (i + u * stride) * 16;
      if (MODE == 0) v[u] = ld_stream(s);                  // plain / peer read
      if (MODE == 1) v[u] = multimem_ld_reduce_add_f32(s); // in-switch reduce
      if (MODE == 2) v[u] = ld_stream(s);                  // local read, multicast store
      if (MODE == 3) v[u] = multimem_ld_reduce_add_bf16(s);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      char *d = dst + (i + u * stride) * 16;
      if (MODE == 2) multimem_st16(d, v[u]);
      else st_stream(d, v[u]);
    }
  }
  for (; i < nvec; i += stride) {
    Vec16 v = (MODE == 1) ? multimem_ld_reduce_add_f32(src + i * 16)
              : (MODE == 3) ? multimem_ld_reduce_add_bf16(src + i * 16) : ld_stream(src + i * 16);
    if (MODE == 2) multimem_st16(dst + i * 16, v);
    else st_stream(dst + i * 16, v);
  }
}

// persistent poller: waits for doorbell values, echoes them to `ack`
__global__ void k_poller(const uint64_t *doorbell, uint64_t *ack, uint64_t last) {
  uint64_t seen = 0;
  while (seen < last) {
    uint64_t v = ld_acquire_sys(doorbell);
    if (v > seen) {
      seen = v;
      st_release_sys(ack, v);
    } else {
      nanosleep(20);
    }
  }
}

static float time_ms(cudaStream_t s, int iters, const std::function<void()> &f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(a, s));
  for (int i = 0; i < iters; ++i) f();
  CK(cudaEventRecord(b, s));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / iters;
}

int main(int argc, char **argv) {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  auto oob = TcpOob::from_env();
  const int rank = oob->rank(), world = oob->size();
  int dev = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : rank;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (getenv("ACCL_PROBE_SAME_GPU")) dev = 0;
  dev %= ndev;
  CK(cudaSetDevice(dev));
  auto topo = cuda::probe_topology(dev);
  printf("[r%d] %s\n", rank, topo.describe().c_str());
  for (int p = 0; p < ndev && rank == 0; ++p) {
    int can = 0;
    if (p != dev) cudaDeviceCanAccessPeer(&can, dev, p);
    printf("[r0] canAccessPeer(%d->%d)=%d\n", dev, p, can);
  }
  size_t heap_bytes = (argc > 1 ? strtoull(argv[1], nullptr, 0) : 1024ull) << 20;
  cuda::SymHeap heap(*oob, dev, heap_bytes, true);
  printf("[r%d] heap bytes=%zu mc=%p note='%s'\n", rank, heap.bytes(), (void *)heap.mc_base(), heap.multicast_note().c_str());
  auto &drv = cuda::DriverApi::get();
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));

  // layout: [0,4K) flags, data from 1 MiB
  const size_t DATA = 1 << 20;
  uint32_t *my_flags = reinterpret_cast<uint32_t *>(heap.local());
  const int next = (rank + 1) % world, prev = (rank + world - 1) % world;
  int *bad;
  CK(cudaMallocManaged(&bad, sizeof(int)));
  *bad = 0;

  // 1. peer put + flag
  {
    const size_t n = 1 << 16;
    float *peer_buf = reinterpret_cast<float *>(heap.base(next) + DATA);
    uint32_t *peer_flag = reinterpret_cast<uint32_t *>(heap.base(next)) + 0;
    k_p2p_put<<<1, 256, 0, s>>>(peer_buf, peer_flag, n, 100.f + rank, 7);
    k_p2p_wait_check<<<1, 256, 0, s>>>(reinterpret_cast<float *>(heap.local() + DATA), my_flags + 0, n, 100.f + prev, 7, bad);
    CK(cudaStreamSynchronize(s));
    printf("[r%d] p2p put+flag: %s (bad=%d)\n", rank, *bad ? "FAIL" : "ok", *bad);
  }
  oob->barrier();

  // 2. barriers: p2p and multicast
  uint32_t **d_peer_flags;
  {
    std::vector<uint32_t *> pf;
    for (int r = 0; r < world; ++r) pf.push_back(reinterpret_cast<uint32_t *>(heap.base(r)) + 16);
    CK(cudaMalloc(&d_peer_flags, sizeof(uint32_t *) * world));
    CK(cudaMemcpy(d_peer_flags, pf.data(), sizeof(uint32_t *) * world, cudaMemcpyHostToDevice));
  }
  uint32_t bar_p2p = 0, bar_mc = 0;
  auto barrier_p2p = [&] { bar_p2p += world; k_barrier_p2p<<<1, 32, 0, s>>>(d_peer_flags, my_flags + 16, world, bar_p2p); };
  auto barrier_mc = [&] { bar_mc += world; k_barrier_mc<<<1, 32, 0, s>>>(reinterpret_cast<uint32_t *>(heap.mc_base()) + 32, my_flags + 32, bar_mc); };
  {
    float ms = time_ms(s, 200, barrier_p2p);
    printf("[r%d] p2p-flag barrier kernel: %.2f us per launch\n", rank, ms * 1e3);
    if (heap.has_multicast()) {
      ms = time_ms(s, 200, barrier_mc);
      printf("[r%d] multimem.red barrier kernel: %.2f us per launch\n", rank, ms * 1e3);
    }
  }
  auto gbar = [&] { if (heap.has_multicast()) barrier_mc(); else barrier_p2p(); };

  // 3. multimem correctness
  const size_t NB = std::min<size_t>(heap.bytes() / 4, 256u << 20); // bytes per test buffer
  char *A = heap.local() + DATA, *B = heap.local() + DATA + NB;
  if (heap.has_multicast()) {
    const size_t n = 1 << 20;
    k_fill<<<148, 512, 0, s>>>(reinterpret_cast<float *>(A), n, 1.f + rank);
    gbar();
    k_bw<1><<<64, 512, 0, s>>>(heap.mc_base() + DATA, B, n * 4);
    CK(cudaStreamSynchronize(s));
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), B, n * 4, cudaMemcpyDeviceToHost));
    float expect = world * (world + 1) / 2.f;
    size_t nb = 0;
    for (float x : h) nb += x != expect;
    printf("[r%d] multimem.ld_reduce f32: %s (expect %.1f got %.1f, bad=%zu)\n", rank, nb ? "FAIL" : "ok", expect, h[12345], nb);
    gbar();
    // multimem.st: rank 0 broadcasts
    if (rank == 0) {
      k_fill<<<148, 512, 0, s>>>(reinterpret_cast<float *>(B), n, 42.f);
      k_bw<2><<<64, 512, 0, s>>>(B, heap.mc_base() + DATA, n * 4);
    }
    gbar();
    CK(cudaStreamSynchronize(s));
    CK(cudaMemcpy(h.data(), A, n * 4, cudaMemcpyDeviceToHost));
    nb = 0;
    for (float x : h) nb += x != 42.f;
    printf("[r%d] multimem.st bcast: %s (bad=%zu)\n", rank, nb ? "FAIL" : "ok", nb);
  }
  oob->barrier();

  // 4. bandwidth sweeps
  for (int ctas : {8, 16, 32, 64, 148, 296}) {
    gbar();
    float ms = time_ms(s, 10, [&] { k_bw<0><<<ctas, 512, 0, s>>>(heap.base(next) + DATA, B, NB); });
    double pull = NB / ms * 1e-6;
    gbar();
    ms = time_ms(s, 10, [&] { k_bw<0><<<ctas, 512, 0, s>>>(A, heap.base(next) + DATA + NB, NB); });
    double push = NB / ms * 1e-6;
    double red = 0, redbf = 0, mcst = 0, loc = 0;
    gbar();
    ms = time_ms(s, 10, [&] { k_bw<0><<<ctas, 512, 0, s>>>(A, B, NB); });
    loc = NB / ms * 1e-6;
    if (heap.has_multicast()) {
      gbar();
      ms = time_ms(s, 10, [&] { k_bw<1><<<ctas, 512, 0, s>>>(heap.mc_base() + DATA, B, NB); });
      red = NB / ms * 1e-6;
      gbar();
      ms = time_ms(s, 10, [&] { k_bw<3><<<ctas, 512, 0, s>>>(heap.mc_base() + DATA, B, NB); });
      redbf = NB / ms * 1e-6;
      gbar();
      // every rank broadcasts its own 1/world slice (allgather traffic pattern)
      size_t slice = NB / world / 16 * 16;
      ms = time_ms(s, 10, [&] { k_bw<2><<<ctas, 512, 0, s>>>(B + rank * slice, heap.mc_base() + DATA + rank * slice, slice); });
      mcst = slice * world / ms * 1e-6;
    }
    gbar();
    CK(cudaStreamSynchronize(s));
    printf("[r%d] ctas=%3d  GB/s: local-copy %.0f | peer-pull %.0f | peer-push %.0f | mc.ld_reduce(f32) %.0f (bf16 %.0f) out | mc.st allgather-equivalent %.0f\n",
           rank, ctas, loc, pull, push, red, redbf, mcst);
  }
  oob->barrier();

  // 5. stream memory operations + persistent poller (doorbell in device memory)
  {
    uint64_t *door = reinterpret_cast<uint64_t *>(heap.local() + 4096);
    uint64_t *ack_h = nullptr;
    CK(cudaHostAlloc(&ack_h, 64, cudaHostAllocMapped));
    *ack_h = 0;
    uint64_t *ack_d = nullptr;
    CK(cudaHostGetDevicePointer(&ack_d, ack_h, 0));
    CK(cudaMemset(door, 0, 8));
    cudaStream_t eng, usr;
    CK(cudaStreamCreateWithFlags(&eng, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&usr, cudaStreamNonBlocking));
    const int N = 200;
    k_poller<<<1, 1, 0, eng>>>(door, ack_d, N);
    Timer t;
    double tot = 0;
    bool ok = true;
    for (int i = 1; i <= N; ++i) {
      t.start();
      CUresult r = drv.cuStreamWriteValue64(usr, reinterpret_cast<CUdeviceptr>(door), i, 0);
      if (r != CUDA_SUCCESS) { printf("[r%d] cuStreamWriteValue64 FAILED %s\n", rank, cuda::cu_error_string(r).c_str()); ok = false; break; }
      while (*reinterpret_cast<volatile uint64_t *>(ack_h) < (uint64_t)i) {}
      t.end();
      if (i > 20) tot += t.elapsed_ns();
    }
    if (ok) {
      CK(cudaStreamSynchronize(eng));
      printf("[r%d] host -> cuStreamWriteValue64 -> persistent kernel -> host-mapped ack round trip: %.2f us\n", rank, tot / (N - 20) * 1e-3);
      // stream-ordered completion wait
      CK(cudaMemset(door, 0, 8));
      uint64_t *done = reinterpret_cast<uint64_t *>(heap.local() + 8192);
      CK(cudaMemset(done, 0, 8));
      k_poller<<<1, 1, 0, eng>>>(door, done, N);
      t.start();
      for (int i = 1; i <= N; ++i) {
        drv.cuStreamWriteValue64(usr, reinterpret_cast<CUdeviceptr>(door), i, 0);
        drv.cuStreamWaitValue64(usr, reinterpret_cast<CUdeviceptr>(done), i, CU_STREAM_WAIT_VALUE_GEQ);
      }
      CK(cudaStreamSynchronize(usr));
      t.end();
      printf("[r%d] stream-ordered doorbell+wait pairs: %.2f us each\n", rank, t.elapsed_ns() * 1e-3 / N);
      CK(cudaStreamSynchronize(eng));
    }
  }
  oob->barrier();
  printf("[r%d] probe done\n", rank);
  return 0;
}
