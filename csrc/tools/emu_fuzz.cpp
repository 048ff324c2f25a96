// emu_fuzz: random call programs through the C++ API on the CPU emulator (ranks as threads), every result
// checked against a host-side reference.  Meant to be run under the sanitizers, where random schedules and
// geometries shake out races and lifetime bugs that a fixed test list cannot:
//
//   python -m accl_b200.utils.build --sanitize thread --tool emu_fuzz && build/bin/emu_fuzz_thread 300 7
//   emu_fuzz [programs=200] [seed=1]
//
// The Python property tests (tests/test_emulator_property.py) cover the same space through the bindings.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "accl/accl.hpp"
#include "accl/bootstrap.hpp"
#include "accl/emu/emudevice.hpp"

using namespace accl;

namespace {
enum Op { SENDRECV, BCAST, SCATTER, GATHER, ALLGATHER, REDUCE, ALLREDUCE, REDUCE_SCATTER, ALLTOALL, BARRIER, N_OPS };
const char *op_name[] = {"sendrecv", "bcast", "scatter", "gather", "allgather", "reduce", "allreduce", "reduce_scatter", "alltoall", "barrier"};

struct Step {
  Op op;
  unsigned count, root;
  reduceFunction func;
  unsigned salt;
  bool msg = false;       // true: a point-to-point message src -> dst (asynchronous send, blocking receive)
  unsigned src = 0, dst = 0, tag = 0;
};

// small integers: every sum is exact in fp32, results compare with ==
std::vector<float> data(unsigned n, int rank, unsigned salt) {
  std::mt19937 g(977u * salt + 31u * static_cast<unsigned>(rank) + 5u);
  std::uniform_int_distribution<int> d(-64, 63);
  std::vector<float> v(n);
  for (auto &x : v) x = static_cast<float>(d(g));
  return v;
}
std::vector<float> reduced(int world, unsigned n, reduceFunction f, unsigned salt) {
  std::vector<float> out = data(n, 0, salt);
  for (int r = 1; r < world; ++r) {
    auto x = data(n, r, salt);
    for (unsigned i = 0; i < n; ++i) out[i] = f == reduceFunction::SUM ? out[i] + x[i] : std::max(out[i], x[i]);
  }
  return out;
}
#define CHECK(cond)                                                                                      \
  do {                                                                                                   \
    if (!(cond)) throw std::runtime_error(std::string(op_name[s.op]) + ": " #cond " (line " + std::to_string(__LINE__) + ")"); \
  } while (0)

void run_step(ACCL &a, int r, int w, const Step &s) {
  const unsigned n = s.count, W = static_cast<unsigned>(w), root = s.root % W;
  auto buf = [&](unsigned len) { return a.create_buffer<float>(len, dataType::float32); };
  auto fill = [&](Buffer<float> &b, const std::vector<float> &v) { std::memcpy(b.buffer(), v.data(), v.size() * 4); };
  switch (s.op) {
  case SENDRECV: {
    auto src = buf(n), dst = buf(n);
    fill(*src, data(n, r, s.salt));
    ACCLRequest *q = a.send(*src, n, static_cast<unsigned>((r + 1) % w), s.salt & 0xFF, GLOBAL_COMM, false, dataType::none, true);
    a.free_request(a.recv(*dst, n, static_cast<unsigned>((r + w - 1) % w), s.salt & 0xFF));
    a.wait(q);
    a.free_request(q);
    auto e = data(n, (r + w - 1) % w, s.salt);
    for (unsigned i = 0; i < n; ++i) CHECK((*dst)[i] == e[i]);
    break;
  }
  case BCAST: {
    auto b = buf(n);
    fill(*b, data(n, r, s.salt));
    a.free_request(a.bcast(*b, n, root));
    auto e = data(n, static_cast<int>(root), s.salt);
    for (unsigned i = 0; i < n; ++i) CHECK((*b)[i] == e[i]);
    break;
  }
  case SCATTER: {
    auto src = buf(n * W), dst = buf(n);
    fill(*src, data(n * W, r, s.salt));
    a.free_request(a.scatter(*src, *dst, n, root));
    auto e = data(n * W, static_cast<int>(root), s.salt);
    for (unsigned i = 0; i < n; ++i) CHECK((*dst)[i] == e[static_cast<unsigned>(r) * n + i]);
    break;
  }
  case GATHER: {
    auto src = buf(n), dst = buf(n * W);
    fill(*src, data(n, r, s.salt));
    a.free_request(a.gather(*src, *dst, n, root));
    if (static_cast<unsigned>(r) == root)
      for (unsigned q = 0; q < W; ++q) {
        auto e = data(n, static_cast<int>(q), s.salt);
        for (unsigned i = 0; i < n; ++i) CHECK((*dst)[q * n + i] == e[i]);
      }
    break;
  }
  case ALLGATHER: {
    auto src = buf(n), dst = buf(n * W);
    fill(*src, data(n, r, s.salt));
    a.free_request(a.allgather(*src, *dst, n));
    for (unsigned q = 0; q < W; ++q) {
      auto e = data(n, static_cast<int>(q), s.salt);
      for (unsigned i = 0; i < n; ++i) CHECK((*dst)[q * n + i] == e[i]);
    }
    break;
  }
  case REDUCE: {
    auto src = buf(n), dst = buf(n);
    fill(*src, data(n, r, s.salt));
    a.free_request(a.reduce(*src, *dst, n, root, s.func));
    if (static_cast<unsigned>(r) == root) {
      auto e = reduced(w, n, s.func, s.salt);
      for (unsigned i = 0; i < n; ++i) CHECK((*dst)[i] == e[i]);
    }
    break;
  }
  case ALLREDUCE: {
    auto src = buf(n), dst = buf(n);
    fill(*src, data(n, r, s.salt));
    a.free_request(a.allreduce(*src, *dst, n, s.func));
    auto e = reduced(w, n, s.func, s.salt);
    for (unsigned i = 0; i < n; ++i) CHECK((*dst)[i] == e[i]);
    break;
  }
  case REDUCE_SCATTER: {
    auto src = buf(n * W), dst = buf(n);
    fill(*src, data(n * W, r, s.salt));
    a.free_request(a.reduce_scatter(*src, *dst, n, s.func));
    auto e = reduced(w, n * W, s.func, s.salt);
    for (unsigned i = 0; i < n; ++i) CHECK((*dst)[i] == e[static_cast<unsigned>(r) * n + i]);
    break;
  }
  case ALLTOALL: {
    auto src = buf(n * W), dst = buf(n * W);
    fill(*src, data(n * W, r, s.salt));
    a.free_request(a.alltoall(*src, *dst, n));
    for (unsigned q = 0; q < W; ++q) {
      auto e = data(n * W, static_cast<int>(q), s.salt);
      for (unsigned i = 0; i < n; ++i) CHECK((*dst)[q * n + i] == e[static_cast<unsigned>(r) * n + i]);
    }
    break;
  }
  default: a.free_request(a.barrier()); break;
  }
}
} // namespace

int main(int argc, char **argv) {
  const int programs = argc > 1 ? std::atoi(argv[1]) : 200;
  const unsigned seed = argc > 2 ? static_cast<unsigned>(std::atoi(argv[2])) : 1u;
  std::mt19937 rng(seed);
  auto pick = [&](unsigned lo, unsigned hi) { return std::uniform_int_distribution<unsigned>(lo, hi)(rng); };
  int failed = 0;
  for (int p = 0; p < programs; ++p) {
    const int W = static_cast<int>(pick(2, 5));
    const unsigned bufs[] = {64, 128, 256, 1024, 4096}, nbs[] = {8, 16, 32}, egrs[] = {64, 256, 1024, 4096}, rvs[] = {256, 1024, 32768, 1u << 20};
    const unsigned buf = bufs[pick(0, 4)], nb = nbs[pick(0, 2)];
    unsigned egr = std::max(std::min(egrs[pick(0, 3)], buf * nb / 2), buf);
    const unsigned rv = std::max(rvs[pick(0, 3)], 2 * egr);
    const bool one_hop = pick(0, 1) != 0; // reference-style rings / trees, or the one-hop schedules of the B200 backend
    std::vector<Step> steps(pick(1, 5));
    for (auto &s : steps) {
      s = Step{static_cast<Op>(pick(0, N_OPS - 1)), pick(1, 3000), pick(0, 5), pick(0, 1) ? reduceFunction::SUM : reduceFunction::MAX, pick(0, 1000)};
      if (pick(0, 9) < 3) { // mix in parked / eager point-to-point traffic between the collectives
        s.msg = true;
        s.src = pick(0, static_cast<unsigned>(W) - 1);
        s.dst = pick(0, static_cast<unsigned>(W) - 1);
        s.tag = pick(0, 300);
      }
    }
    auto devs = emu::make_inproc_world(W, 64u << 20);
    std::vector<std::unique_ptr<ACCL>> accls;
    for (auto &d : devs) accls.emplace_back(new ACCL(std::move(d)));
    const auto ranks = generate_ranks(true, W, 5500, buf);
    std::vector<std::string> errs(static_cast<size_t>(W));
    std::vector<std::thread> ts;
    for (int r = 0; r < W; ++r)
      ts.emplace_back([&, r] {
        try {
          accls[static_cast<size_t>(r)]->initialize(ranks, r, static_cast<int>(nb), buf, egr, rv);
          accls[static_cast<size_t>(r)]->free_request(accls[static_cast<size_t>(r)]->set_timeout(30000000)); // peers may be seconds late under a sanitizer
          ACCL &a = *accls[static_cast<size_t>(r)];
          a.set_one_hop_schedules(one_hop);
          std::vector<ACCLRequest *> pending;
          std::vector<std::unique_ptr<Buffer<float>>> keep;
          for (const Step &s : steps) {
            if (!s.msg) {
              run_step(a, r, W, s);
              continue;
            }
            if (s.src == s.dst) continue;
            if (static_cast<unsigned>(r) == s.src) {
              auto b = a.create_buffer<float>(s.count, dataType::float32);
              auto v = data(s.count, r, s.salt);
              std::memcpy(b->buffer(), v.data(), v.size() * 4);
              pending.push_back(a.send(*b, s.count, s.dst, s.tag, GLOBAL_COMM, false, dataType::none, true));
              keep.push_back(std::move(b));
            } else if (static_cast<unsigned>(r) == s.dst) {
              auto b = a.create_buffer<float>(s.count, dataType::float32);
              a.free_request(a.recv(*b, s.count, s.src, s.tag));
              auto e = data(s.count, static_cast<int>(s.src), s.salt);
              for (unsigned i = 0; i < s.count; ++i)
                if ((*b)[i] != e[i]) throw std::runtime_error("message " + std::to_string(s.src) + "->" + std::to_string(s.dst) + ": data mismatch");
            }
          }
          for (ACCLRequest *q : pending) {
            a.wait(q);
            if (a.get_retcode(q) != 0) throw std::runtime_error("asynchronous send failed");
            a.free_request(q);
          }
          a.free_request(a.barrier());
        } catch (const std::exception &e) {
          errs[static_cast<size_t>(r)] = e.what();
        }
      });
    for (auto &t : ts) t.join();
    bool bad = false;
    for (int r = 0; r < W; ++r)
      if (!errs[static_cast<size_t>(r)].empty()) {
        if (!bad) {
          std::printf("FAIL program %d (seed %u): world %d, rx %u x %u, eager <= %u, rendezvous seg %u:", p, seed, W, nb, buf, egr, rv);
          for (const Step &s : steps) {
            if (s.msg) std::printf(" msg(%u->%u, %u, tag %u)", s.src, s.dst, s.count, s.tag);
            else std::printf(" %s(%u, root %u, %s)", op_name[s.op], s.count, s.root, s.func == reduceFunction::SUM ? "sum" : "max");
          }
          std::printf("\n");
        }
        std::printf("  rank %d: %s\n", r, errs[static_cast<size_t>(r)].c_str());
        bad = true;
      }
    failed += bad;
  }
  std::printf("emu_fuzz: %d program(s), %d failed (seed %u)\n", programs, failed, seed);
  return failed != 0;
}
