// emu_bench: sweep benchmark through the C++ API on the CPU emulator, engine-measured durations
// (get_duration), one CSV row per (collective, size): the counterpart of the reference's
// ACCLSweepBenchmark (test/host/xrt/src/bench.cpp:25-61: count = 2^4..2^19 fp32 over sendrecv, bcast,
// scatter, gather, allgather, reduce, reduce_scatter, allreduce; CSV Test,Param,Cycles).
//
//   emu_bench [world=4] [min_log2=4] [max_log2=19] [iters=5] > emu_bench.csv
//
// The GPU sweeps live in bench/sweep.py (they need NCCL side by side); this one characterises the
// emulator itself (protocol switch-over, per-call overhead of the step machines).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "accl/accl.hpp"
#include "accl/emu/emudevice.hpp"

using namespace accl;

int main(int argc, char **argv) {
  const int W = argc > 1 ? std::atoi(argv[1]) : 4;
  const int lo = argc > 2 ? std::atoi(argv[2]) : 4, hi = argc > 3 ? std::atoi(argv[3]) : 19;
  const int iters = argc > 4 ? std::atoi(argv[4]) : 5;
  auto devs = emu::make_inproc_world(W, 256u << 20);
  std::vector<std::unique_ptr<ACCL>> accls;
  for (auto &d : devs) accls.emplace_back(new ACCL(std::move(d)));
  std::vector<rank_t> ranks;
  for (int i = 0; i < W; ++i) ranks.emplace_back("127.0.0.1", 5500 + i, i, 1024);
  const char *ops[] = {"sendrecv", "bcast", "scatter", "gather", "allgather", "reduce", "reduce_scatter", "allreduce"};
  std::mutex out_m;
  std::vector<std::string> rows;
  std::vector<std::thread> ts;
  for (int r = 0; r < W; ++r)
    ts.emplace_back([&, r] {
      ACCL &a = *accls[r];
      // the reference's default test geometry: 16 x 1 KiB rx buffers, 1 KiB eager limit, 32 KiB rendezvous segments
      a.initialize(ranks, r, 16, 1024, 1024, 32 * 1024);
      const unsigned maxn = 1u << hi;
      auto s = a.create_buffer<float>(static_cast<size_t>(maxn) * W, dataType::float32);
      auto d = a.create_buffer<float>(static_cast<size_t>(maxn) * W, dataType::float32);
      for (size_t i = 0; i < static_cast<size_t>(maxn) * W; ++i) (*s)[i] = static_cast<float>(r);
      s->sync_to_device();
      for (const char *op : ops)
        for (int lg = lo; lg <= hi; ++lg) {
          const unsigned n = 1u << lg;
          std::vector<uint64_t> ns;
          for (int it = 0; it < iters; ++it) {
            a.free_request(a.barrier());
            ACCLRequest *q = nullptr;
            const std::string o = op;
            // device-resident operands (from_fpga / to_fpga): time the engine, not the host mirrors
            if (o == "sendrecv") {
              const unsigned nxt = static_cast<unsigned>((r + 1) % W), prv = static_cast<unsigned>((r + W - 1) % W);
              ACCLRequest *sq = a.send(*s, n, nxt, 1, GLOBAL_COMM, true, dataType::none, true);
              q = a.recv(*d, n, prv, 1, GLOBAL_COMM, true);
              a.wait(sq);
              a.free_request(sq);
            } else if (o == "bcast") q = a.bcast(*s, n, 0, GLOBAL_COMM, true, true);
            else if (o == "scatter") q = a.scatter(*s, *d, n, 0, GLOBAL_COMM, true, true);
            else if (o == "gather") q = a.gather(*s, *d, n, 0, GLOBAL_COMM, true, true);
            else if (o == "allgather") q = a.allgather(*s, *d, n, GLOBAL_COMM, true, true);
            else if (o == "reduce") q = a.reduce(*s, *d, n, 0, reduceFunction::SUM, GLOBAL_COMM, true, true);
            else if (o == "reduce_scatter") q = a.reduce_scatter(*s, *d, n, reduceFunction::SUM, GLOBAL_COMM, true, true);
            else q = a.allreduce(*s, *d, n, reduceFunction::SUM, GLOBAL_COMM, true, true);
            ns.push_back(a.get_duration(q));
            a.free_request(q);
          }
          std::sort(ns.begin(), ns.end());
          if (r == 0) {
            std::lock_guard<std::mutex> g(out_m);
            rows.push_back(std::string(op) + "," + std::to_string(n) + "," + std::to_string(static_cast<unsigned long long>(n) * 4) +
                           "," + std::to_string(ns[ns.size() / 2]) + "," + std::to_string(ns.front()));
          }
        }
      a.free_request(a.barrier());
    });
  for (auto &t : ts) t.join();
  std::printf("Test,Count,Bytes,MedianNs,MinNs\n");
  for (auto &row : rows) std::printf("%s\n", row.c_str());
  return 0;
}
