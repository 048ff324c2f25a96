#include "accl/cuda/driver_api.hpp"

#include <mutex>

namespace accl {
namespace cuda {

namespace {
std::once_flag g_once;
DriverApi g_api;

template <typename F> bool resolve(const char *name, F &fn, std::string &err) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    err = std::string("cannot resolve driver symbol ") + name + ": " +
          (e != cudaSuccess ? cudaGetErrorString(e) : "not found");
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

void load_once() {
  bool ok = true;
#define ACCL_RESOLVE(name)                                                   \
  if (ok) ok = resolve(#name, g_api.name, g_api.load_error);
  ACCL_DRIVER_SYMBOLS(ACCL_RESOLVE)
#undef ACCL_RESOLVE
  g_api.loaded = ok;
}
} // namespace

DriverApi &DriverApi::get() {
  std::call_once(g_once, load_once);
  if (!g_api.loaded)
    throw std::runtime_error("CUDA driver unavailable: " + g_api.load_error);
  return g_api;
}

bool DriverApi::available() {
  std::call_once(g_once, load_once);
  return g_api.loaded;
}

std::string cu_error_string(CUresult r) {
  const char *s = nullptr;
  if (g_api.loaded && g_api.cuGetErrorString &&
      g_api.cuGetErrorString(r, &s) == CUDA_SUCCESS && s)
    return std::string(s) + " [" + std::to_string(static_cast<int>(r)) + "]";
  return "CUresult " + std::to_string(static_cast<int>(r));
}

} // namespace cuda
} // namespace accl
