// Internal helpers shared by the translation units of libvilbert_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/vilbert_b200.h"

namespace vb {
// Records a printf-style message for vb_last_error() and returns `code`.
int set_error(int code, const char* fmt, ...);
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(VB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return VB_OK;
}
int sm_count();
}  // namespace vb
