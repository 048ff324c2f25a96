// Lazily-resolved CUDA driver API table.
//
// The library never links against libcuda.so: the extension must import on a
// GPU-less build box.  Every driver symbol is resolved at first use through
// cudaGetDriverEntryPoint (served by the statically linked runtime), which
// only dlopens libcuda when a GPU call is actually made.
//
// Role in the reference: the XRT / Coyote shells that `XRTDevice` and
// `CoyoteDevice` sit on (driver/xrt/src/xrtdevice.cpp, coyotedevice.cpp).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

namespace accl {
namespace cuda {

#define ACCL_DRIVER_SYMBOLS(X)         \
  X(cuGetErrorString)                  \
  X(cuDeviceGet)                       \
  X(cuDeviceGetAttribute)              \
  X(cuCtxGetCurrent)                   \
  X(cuMemCreate)                       \
  X(cuMemRelease)                      \
  X(cuMemAddressReserve)               \
  X(cuMemAddressFree)                  \
  X(cuMemMap)                          \
  X(cuMemUnmap)                        \
  X(cuMemSetAccess)                    \
  X(cuMemGetAllocationGranularity)     \
  X(cuMemExportToShareableHandle)      \
  X(cuMemImportFromShareableHandle)    \
  X(cuMulticastCreate)                 \
  X(cuMulticastAddDevice)              \
  X(cuMulticastBindMem)                \
  X(cuMulticastGetGranularity)         \
  X(cuMulticastUnbind)                 \
  X(cuStreamWriteValue64)              \
  X(cuStreamWaitValue64)               \
  X(cuStreamWriteValue32)              \
  X(cuStreamWaitValue32)               \
  X(cuStreamBatchMemOp)

struct DriverApi {
#define ACCL_DECL(name) decltype(&::name) name = nullptr;
  ACCL_DRIVER_SYMBOLS(ACCL_DECL)
#undef ACCL_DECL
  bool loaded = false;
  std::string load_error;

  // Throws std::runtime_error when no driver is present (CPU-only box).
  static DriverApi &get();
  // Non-throwing probe.
  static bool available();
};

std::string cu_error_string(CUresult r);

#define ACCL_CU(call)                                                        \
  do {                                                                       \
    CUresult _r = (call);                                                    \
    if (_r != CUDA_SUCCESS)                                                  \
      throw std::runtime_error(std::string(#call) + " failed: " +            \
                               ::accl::cuda::cu_error_string(_r) + " (" +    \
                               __FILE__ + ":" + std::to_string(__LINE__) +   \
                               ")");                                         \
  } while (0)

#define ACCL_CUDART(call)                                                    \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess)                                                   \
      throw std::runtime_error(std::string(#call) + " failed: " +            \
                               cudaGetErrorString(_e) + " (" + __FILE__ +    \
                               ":" + std::to_string(__LINE__) + ")");        \
  } while (0)

} // namespace cuda
} // namespace accl
