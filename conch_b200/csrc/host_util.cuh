// Host-side helpers shared by the kernel launchers: the driver's tensor-map encoder (looked up once through the runtime,
// so the library links against libcudart only) and per-DEVICE cached facts.  Everything that CUDA scopes to a device or
// context (SM count, "is this sm_100", cudaFuncSetAttribute) is cached per device index, never per process: one process
// may drive several GPUs.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cb200 {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();  // nullptr when the driver entry point is missing

constexpr int CB200_MAX_DEVICES = 64;
struct DeviceInfo {
  int index;     // cudaGetDevice() of the calling thread
  int sm_major;  // compute capability major (10 on B200)
  int sms;       // multiprocessor count
};
// Facts about the calling thread's current device (queried once per device).  Returns nullptr on a CUDA error.
const DeviceInfo* current_device_info();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device): `mask` is a per-instantiation
// static bitmask indexed by device.
template <class Kern>
int set_max_smem_once(Kern kern, int bytes, unsigned long long& mask) {
  const DeviceInfo* di = current_device_info();
  const int dev = di ? di->index : 0;
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(mask & bit)) {
    CB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    mask |= bit;
  }
  return 0;
}

}  // namespace cb200
