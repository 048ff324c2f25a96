// emu_selftest: C++-only smoke test of the host API on the CPU emulator
// (ranks as threads).  Mirrors BASELINE config #1: send/recv + allreduce fp32.
// Also the unit used for ASAN/TSAN builds of the emulator:
//   g++ -fsanitize=address,undefined -g -Icsrc/include csrc/tools/emu_selftest.cpp csrc/src/host/*.cpp csrc/src/emu/*.cpp -lpthread
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

#include "accl/accl.hpp"
#include "accl/emu/emudevice.hpp"

using namespace accl;

int main(int argc, char **argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 2;
  auto devs = emu::make_inproc_world(W, 64u << 20);
  std::vector<std::unique_ptr<ACCL>> accls;
  for (auto &d : devs) accls.emplace_back(new ACCL(std::move(d)));
  std::vector<rank_t> ranks;
  for (int i = 0; i < W; ++i) ranks.emplace_back("127.0.0.1", 5500 + i, i, 1024);
  std::vector<int> fails(W, 0);
  std::vector<std::thread> ts;
  for (int r = 0; r < W; ++r)
    ts.emplace_back([&, r] {
      try {
        ACCL &a = *accls[r];
        a.initialize(ranks, r, 16, 1024, 1024, 32768);
        for (unsigned n : {16u, 1000u, 20000u}) {
          auto s = a.create_buffer<float>(n, dataType::float32);
          auto d = a.create_buffer<float>(n, dataType::float32);
          for (unsigned i = 0; i < n; ++i) (*s)[i] = float(i % 13) + r;
          const int nxt = (r + 1) % W, prv = (r + W - 1) % W;
          if (W > 1) {
            // async send + blocking recv: legal for eager and for rendezvous (a blocking
            // rendezvous send before the peer posts its recv would wait forever, as in MPI)
            ACCLRequest *sreq = a.send(*s, n, nxt, 5, GLOBAL_COMM, false, dataType::none, true);
            a.free_request(a.recv(*d, n, prv, 5));
            a.wait(sreq);
            a.free_request(sreq);
            for (unsigned i = 0; i < n; ++i)
              if ((*d)[i] != float(i % 13) + prv) { fails[r]++; break; }
          }
          a.free_request(a.allreduce(*s, *d, n, reduceFunction::SUM));
          for (unsigned i = 0; i < n; ++i) {
            float e = 0;
            for (int q = 0; q < W; ++q) e += float(i % 13) + q;
            if (std::fabs((*d)[i] - e) > 1e-4) { fails[r]++; break; }
          }
        }
        a.free_request(a.barrier());
      } catch (const std::exception &e) {
        fprintf(stderr, "rank %d: %s\n", r, e.what());
        fails[r] = 100;
      }
    });
  for (auto &t : ts) t.join();
  int bad = 0;
  for (int f : fails) bad += f;
  printf("emu_selftest world=%d: %s\n", W, bad ? "FAIL" : "ok");
  return bad != 0;
}
