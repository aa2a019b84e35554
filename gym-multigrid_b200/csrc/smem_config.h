// smem_config.h -- host helper: opt a kernel in to `bytes` of dynamic shared memory, never lowering a limit that another
// handle (a larger grid, a different mode) raised earlier.  cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel and
// device, shared by every handle of the process.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

namespace mg {

inline cudaError_t raise_smem_limit(const void* fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> limit;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = limit[{fn, dev}];
  if (bytes <= cur) return cudaSuccess;
  if (bytes > 48 * 1024 || cur > 0) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
  }
  cur = bytes;
  return cudaSuccess;
}

}  // namespace mg
