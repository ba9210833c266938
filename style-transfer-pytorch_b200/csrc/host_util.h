// Host-side helpers shared by the .cu translation units: error plumbing and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "stb200.h"

namespace stb {

// Error codes of the C-ABI come from include/stb200.h (STB_OK, STB_ERR_*).
std::string& last_error_string();          // thread-local message buffer (api.cu)
int set_error(int code, const char* fmt, ...);

#define STB_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::stb::set_error(STB_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,   \
                              cudaGetErrorString(_e));                                             \
  } while (0)

#define STB_CHECK(cond, code, ...)                               \
  do {                                                           \
    if (!(cond)) return ::stb::set_error((code), __VA_ARGS__);   \
  } while (0)

#define STB_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// 3-D bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box are innermost-first; strides in bytes for
// dims 1 and 2 (dim 0 is contiguous).
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                      uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2);

// 2-D fp32 row-major [rows][cols] tensor map, SWIZZLE_128B (box_cols * 4 bytes must be 128), zero OOB fill.
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                     uint32_t box_rows);

// 2-D fp16 row-major [rows][cols] tensor map, SWIZZLE_128B (box_cols * 2 bytes must be 128), zero OOB fill.
int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                     uint32_t box_rows);

int num_sms();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember it per (kernel, device) so that contexts
// on several GPUs of one process all get it.
int ensure_dynamic_smem(const void* func, int bytes);

}  // namespace stb
