// pybind11 glue for the CUDA backend (kept out of pyaccl.cpp so that the
// emulator-only build has no CUDA dependency).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "accl/accl.hpp"
#include "accl/cuda/cudadevice.hpp"
#include "accl/cuda/driver_api.hpp"
#include "accl/cuda/plan.hpp"
#include "accl/cuda/plugins.hpp"

namespace py = pybind11;

namespace accl {
namespace cuda {

static CudaConfig make_cfg(int device, size_t heap_mb, bool multicast, int max_ctas, bool engine, int nvls_min_ranks,
                           size_t oneshot_kb, int nvls_ops = -1) {
  CudaConfig c;
  c.device = device;
  c.heap_bytes = heap_mb << 20;
  c.multicast = multicast;
  c.max_ctas = max_ctas;
  c.engine = engine;
  c.nvls_min_ranks = nvls_min_ranks;
  c.oneshot_max_bytes = oneshot_kb << 10;
  if (nvls_ops >= 0) c.nvls_ops = static_cast<uint32_t>(nvls_ops);
  return c;
}

void bind_cuda(py::module_ &m) {
  m.def("cuda_driver_available", [] { return DriverApi::available(); });
  // The call planner (plan.hpp) as a pure function, for unit tests on machines without a GPU: which protocol /
  // algorithm / channel count a call of `count` elements of `dtype` gets on a communicator of `world` ranks.
  m.def("cuda_plan", [](operation op, uint32_t count, dataType dtype, uint32_t world, uint32_t max_eager_bytes, uint32_t max_ctas,
                        bool has_mc, uint32_t nvls_min_ranks, uint64_t oneshot_max_bytes, bool compressed) {
    std::vector<uint32_t> exch(exchmem::SIZE_WORDS, 0);
    exch[exchmem::MAX_EAGER_SIZE / 4] = max_eager_bytes;
    PlanCfg cfg{};
    cfg.max_ctas = max_ctas;
    cfg.nvls_min_ranks = nvls_min_ranks;
    cfg.has_mc = has_mc ? 1 : 0;
    cfg.heap_world = world;
    cfg.oneshot_max_bytes = oneshot_max_bytes;
    cfg.nvls_ops = NVLS_OPS_DEFAULT;
    WorkItem w{};
    w.desc.scenario = static_cast<uint32_t>(op);
    w.desc.count = count;
    w.desc.compression_flags = compressed ? 1u : 0u;
    w.comm_size = world;
    w.udtype = static_cast<uint32_t>(dtype);
    plan_call(exch.data(), cfg, w);
    static const char *names[] = {"auto", "local", "eager", "nvls", "p2p", "p2p_oneshot"};
    py::dict d;
    d["algo"] = w.algo < 6 ? names[w.algo] : "?";
    d["n_ctas"] = w.n_ctas;
    d["use_mc"] = (w.flags & WF_USE_MC) != 0;
    return d;
  }, py::arg("op"), py::arg("count"), py::arg("dtype"), py::arg("world"), py::arg("max_eager_bytes") = 65536,
        py::arg("max_ctas") = 128, py::arg("has_mc") = true, py::arg("nvls_min_ranks") = 3,
        py::arg("oneshot_max_bytes") = 2u << 20, py::arg("compressed") = false);
  m.def("cuda_debug_state", [](ACCL &a) {
    auto *d = dynamic_cast<CudaDevice *>(a.device());
    if (!d) throw std::runtime_error("not a CUDA backend");
    return d->debug_state();
  });
  // out_shard[M/P, N] (heap buffer, bf16) = reduce_scatter_M( A[M,K] @ W[N,K]^T ), fused on tcgen05 + NVLink
  m.def("gemm_reduce_scatter", [](ACCL &a, uintptr_t a_ptr, uintptr_t w_ptr, BaseBuffer &out, uint32_t M, uint32_t N, uint32_t K,
                                  uintptr_t stream) {
    auto *d = dynamic_cast<CudaDevice *>(a.device());
    if (!d) throw std::runtime_error("not a CUDA backend");
    GemmRsArgs g{reinterpret_cast<const void *>(a_ptr), reinterpret_cast<const void *>(w_ptr), out.address(), M, N, K, 0};
    cudaError_t e = launch_gemm_rs(*d, g, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_reduce_scatter launch: ") + cudaGetErrorString(e));
  }, py::call_guard<py::gil_scoped_release>());
  // out = allreduce_sum(x + y): the kernel computes and then issues the collective itself (device API -> engine)
  m.def("vadd_allreduce", [](ACCL &a, BaseBuffer &x, BaseBuffer &y, BaseBuffer &tmp, BaseBuffer &out, uint32_t count,
                             uintptr_t status_dev_ptr, uintptr_t stream) {
    auto *d = dynamic_cast<CudaDevice *>(a.device());
    if (!d) throw std::runtime_error("not a CUDA backend");
    cudaError_t e = launch_vadd_allreduce(*d, x.address(), y.address(), tmp.address(), out.address(), count,
                                          static_cast<uint32_t>(a.get_communicator_addr(GLOBAL_COMM)),
                                          static_cast<uint32_t>(a.get_arithmetic_config_addr({dataType::float32, dataType::float32})),
                                          reinterpret_cast<uint32_t *>(status_dev_ptr), reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) throw std::runtime_error(std::string("vadd_allreduce launch: ") + cudaGetErrorString(e));
  }, py::call_guard<py::gil_scoped_release>());
  m.def("stream_loopback", [](ACCL &a, BaseBuffer &scratch, uint32_t count, bool add_one, uintptr_t status_dev_ptr,
                              uintptr_t stream) {
    auto *d = dynamic_cast<CudaDevice *>(a.device());
    if (!d) throw std::runtime_error("not a CUDA backend");
    cudaError_t e = launch_loopback(*d, scratch.address(), count, add_one, reinterpret_cast<uint32_t *>(status_dev_ptr),
                                    reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) throw std::runtime_error(std::string("stream_loopback launch: ") + cudaGetErrorString(e));
  }, py::call_guard<py::gil_scoped_release>());
  m.def("cuda_probe", [](int device) { return probe_topology(device).describe(); });
  // N ranks in this process (threads), rank i on devices[i]
  m.def("make_cuda_world", [](std::vector<int> devices, size_t heap_mb, bool multicast, int max_ctas, bool engine,
                              int nvls_min_ranks, size_t oneshot_kb, int nvls_ops) {
    std::vector<std::unique_ptr<ACCL>> out;
    auto devs = make_local_world(devices, make_cfg(0, heap_mb, multicast, max_ctas, engine, nvls_min_ranks, oneshot_kb, nvls_ops));
    for (auto &d : devs) out.emplace_back(new ACCL(std::move(d)));
    return out;
  }, py::arg("devices"), py::arg("heap_mb") = 256, py::arg("multicast") = true, py::arg("max_ctas") = 32,
        py::arg("engine") = false, py::arg("nvls_min_ranks") = 3, py::arg("oneshot_kb") = 2048, py::arg("nvls_ops") = -1,
        py::call_guard<py::gil_scoped_release>());
  // one rank per process; bootstrap over a private TCP rendezvous on addr:port
  m.def("make_cuda_rank", [](int rank, int world, int device, const std::string &addr, int port, size_t heap_mb,
                             bool multicast, int max_ctas, bool engine, int nvls_min_ranks, size_t oneshot_kb, int nvls_ops) {
    auto oob = std::make_shared<TcpOob>(rank, world, addr, port);
    auto dev = std::unique_ptr<CCLO>(
        new CudaDevice(oob, make_cfg(device, heap_mb, multicast, max_ctas, engine, nvls_min_ranks, oneshot_kb, nvls_ops)));
    return std::unique_ptr<ACCL>(new ACCL(std::move(dev)));
  }, py::arg("rank"), py::arg("world_size"), py::arg("device"), py::arg("addr") = "127.0.0.1", py::arg("port") = 29637,
        py::arg("heap_mb") = 1024, py::arg("multicast") = true, py::arg("max_ctas") = 32, py::arg("engine") = false,
        py::arg("nvls_min_ranks") = 3, py::arg("oneshot_kb") = 2048, py::arg("nvls_ops") = -1,
        py::call_guard<py::gil_scoped_release>());
}

} // namespace cuda
} // namespace accl
