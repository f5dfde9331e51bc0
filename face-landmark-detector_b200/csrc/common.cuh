// common.cuh — shared host/device helpers for libfld_sm100.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include "../../include/fld.h"

struct fld_handle {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  void* encode_tiled;  // cuTensorMapEncodeTiled entry point (driver API, resolved at create)
};

void fld_set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_fld_launches;

#define FLD_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      fld_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return FLD_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define FLD_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      fld_set_error(__VA_ARGS__);         \
      return FLD_ERR_INVALID;             \
    }                                     \
  } while (0)

// call after every kernel launch: counts the launch and surfaces launch-configuration errors
#define FLD_LAUNCHED()                                                                          \
  do {                                                                                          \
    g_fld_launches.fetch_add(1, std::memory_order_relaxed);                                     \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      fld_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));  \
      return FLD_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

static inline int fld_div_up(int a, int b) { return (a + b - 1) / b; }
static inline int fld_enter(const fld_handle* h) {
  if (!h) { fld_set_error("null handle"); return FLD_ERR_INVALID; }
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) { fld_set_error("cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e)); return FLD_ERR_CUDA; }
  return FLD_OK;
}
