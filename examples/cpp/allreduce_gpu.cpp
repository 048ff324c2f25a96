// C++ consumer on GPUs: one rank per visible GPU (threads of this process), device-resident operands in
// the symmetric heap, stream-ordered asynchronous all-reduce, engine-measured duration.
//
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -DACCL_WITH_CUDA -Icsrc/include \
//        examples/cpp/allreduce_gpu.cpp build/lib/libaccl.a -lpthread -ldl -lrt -o allreduce_gpu
//
// One process per GPU works the same way with accl::TcpOob::from_env() and the CudaDevice constructor that
// takes an Oob (see csrc/src/cuda/bind_cuda.cpp `make_cuda_rank`).
#include <cuda_runtime.h>

#include <cstdio>
#include <thread>
#include <vector>

#include "accl/accl.hpp"
#include "accl/bootstrap.hpp"
#include "accl/cuda/cudadevice.hpp"

int main() {
  using namespace accl;
  int ngpu = 0;
  if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu == 0) {
    std::printf("no GPU visible\n");
    return 0;
  }
  std::vector<int> devices;
  for (int i = 0; i < ngpu; ++i) devices.push_back(i);
  cuda::CudaConfig cfg;
  cfg.heap_bytes = 1ull << 30;
  cfg.max_ctas = 64;
  auto devs = cuda::make_local_world(devices, cfg);
  const int W = static_cast<int>(devs.size());
  std::vector<std::unique_ptr<ACCL>> world;
  for (auto &d : devs) world.emplace_back(new ACCL(std::move(d)));
  const std::vector<rank_t> ranks = generate_ranks(true, W, 5500, 64 << 10);
  const unsigned n = 16u << 20; // 64 MiB of fp32 per rank
  std::vector<int> ok(static_cast<size_t>(W), 0);
  std::vector<std::thread> threads;
  for (int r = 0; r < W; ++r)
    threads.emplace_back([&, r] {
      cudaSetDevice(devices[static_cast<size_t>(r)]);
      ACCL &accl = *world[static_cast<size_t>(r)];
      accl.initialize(ranks, r, 4, 64 << 10, 64 << 10, 1u << 30);
      auto src = accl.create_buffer<float>(n, dataType::float32);
      auto dst = accl.create_buffer<float>(n, dataType::float32);
      std::vector<float> h(n, static_cast<float>(r + 1));
      cudaMemcpy(src->device_ptr(), h.data(), n * sizeof(float), cudaMemcpyHostToDevice);
      cudaStream_t stream;
      cudaStreamCreate(&stream);
      accl.set_stream(stream);
      // operands already on the device (from_fpga / to_fpga), asynchronous, ordered on `stream`
      ACCLRequest *req = accl.allreduce(*src, *dst, n, reduceFunction::SUM, GLOBAL_COMM, true, true, dataType::none, true);
      accl.wait(req);
      std::printf("rank %d: %u floats all-reduced in %.1f us (engine-measured)\n", r, n, accl.get_duration(req) / 1e3);
      accl.free_request(req);
      float first = 0;
      cudaMemcpyAsync(&first, dst->device_ptr(), sizeof(float), cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      ok[static_cast<size_t>(r)] = first == W * (W + 1) / 2.0f;
      accl.free_request(accl.barrier());
      cudaStreamDestroy(stream);
    });
  for (auto &t : threads) t.join();
  int bad = 0;
  for (int v : ok) bad += !v;
  std::printf("%s\n", bad ? "FAILED" : "all ranks ok");
  return bad;
}
