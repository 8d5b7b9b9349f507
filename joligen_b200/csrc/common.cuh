// Shared host-side helpers for the jg_b200 C-ABI library: error reporting, driver entry points,
// TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/jg_b200.h"

namespace jg {

void set_error(const char* fmt, ...);
const char* get_error();

#define JG_CHECK(cond, code, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      ::jg::set_error(__VA_ARGS__);      \
      return (code);                     \
    }                                    \
  } while (0)

#define JG_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::jg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return JG_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

// Every kernel launch of the library is followed by this check; it also counts the launch (jg_kernel_launches()).
#define JG_LAUNCH_CHECK()                                                                        \
  do {                                                                                           \
    ::jg::count_launch();                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) {                                                                    \
      ::jg::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return JG_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

// Build a bf16 tiled tensor map with 128B swizzle.  dims/box: innermost first.  strides_bytes[i] is
// the byte stride of dimension i+1 (rank-1 entries).  Returns 0 on success.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

int num_sms();
void count_launch();

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace jg
