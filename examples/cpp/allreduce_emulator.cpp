// Minimal C++ consumer of the library: four ranks (threads) on the CPU emulator, one all-reduce and a
// neighbour exchange through the public API.  Build against the static library:
//
//   python -m accl_b200.utils.build            # produces build/lib/libaccl.a
//   g++ -std=c++17 -O2 -Icsrc/include examples/cpp/allreduce_emulator.cpp build/lib/libaccl.a \
//       /usr/local/cuda/lib64/libcudart_static.a -lpthread -ldl -lrt -o allreduce_emulator
//   (omit libcudart_static.a for a --cpu-only build)
//
// The same code drives GPUs by constructing the ACCL objects from accl::cuda::CudaDevice instead of
// accl::emu::EmuDevice (see csrc/src/cuda/bind_cuda.cpp: make_local_world / TcpOob).
#include <cstdio>
#include <thread>
#include <vector>

#include "accl/accl.hpp"
#include "accl/bootstrap.hpp"
#include "accl/emu/emudevice.hpp"

int main() {
  using namespace accl;
  const int W = 4;
  const unsigned n = 1 << 16;
  auto devices = emu::make_inproc_world(W, 64u << 20);
  std::vector<std::unique_ptr<ACCL>> world;
  for (auto &d : devices) world.emplace_back(new ACCL(std::move(d)));
  const std::vector<rank_t> ranks = generate_ranks(/*local=*/true, W, 5500, 1024);
  std::vector<int> ok(W, 0);
  std::vector<std::thread> threads;
  for (int r = 0; r < W; ++r)
    threads.emplace_back([&, r] {
      ACCL &accl = *world[r];
      accl.initialize(ranks, r, /*n_egr_rx_bufs=*/16, /*egr_rx_buf_size=*/1024, /*max_egr_size=*/1024, /*max_rndzv_size=*/1 << 20);
      auto src = accl.create_buffer<float>(n, dataType::float32);
      auto dst = accl.create_buffer<float>(n, dataType::float32);
      for (unsigned i = 0; i < n; ++i) (*src)[i] = static_cast<float>(r + 1);
      ACCLRequest *req = accl.allreduce(*src, *dst, n, reduceFunction::SUM);   // blocking; host-resident operands
      std::printf("rank %d: allreduce took %llu ns in the engine\n", r, static_cast<unsigned long long>(accl.get_duration(req)));
      accl.free_request(req);
      bool good = (*dst)[0] == W * (W + 1) / 2.0f && (*dst)[n - 1] == W * (W + 1) / 2.0f;
      // ring exchange: asynchronous send, blocking receive
      ACCLRequest *s = accl.send(*src, 256, (r + 1) % W, /*tag=*/7, GLOBAL_COMM, false, dataType::none, /*run_async=*/true);
      accl.free_request(accl.recv(*dst, 256, (r + W - 1) % W, 7));
      accl.wait(s);
      accl.free_request(s);
      good = good && (*dst)[0] == static_cast<float>((r + W - 1) % W + 1);
      accl.free_request(accl.barrier());
      ok[r] = good;
    });
  for (auto &t : threads) t.join();
  int bad = 0;
  for (int r = 0; r < W; ++r) bad += !ok[r];
  std::printf("%s\n", bad ? "FAILED" : "all ranks ok");
  return bad;
}
