#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*ProbeEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline ProbeEncodeFn probe_encode() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<ProbeEncodeFn>(p);
}
