// Host-side helpers shared by every translation unit of liblavila_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lv {

// error plumbing: thread-local message, integer return codes (0 ok, <0 invalid arg, >0 cudaError_t)
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError() -> code, counts the launch
void count_launch(int n = 1);

int sm_count();                       // SMs of the current device (cached per device)

// cuTensorMapEncodeTiled resolved at run time through cudaGetDriverEntryPoint, so the library links
// against nothing but cudart and can be dlopen'ed on a box without a driver (symbol-export test).
// 2D bf16 tensor: `inner` contiguous elements per row, `rows` rows, `ld` elements between rows.
// Box = box_inner x box_rows, 128-byte swizzle (box_inner * 2 bytes must be <= 128).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t ld,
                      uint32_t box_inner, uint32_t box_rows);

}  // namespace lv

#define LV_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return ::lv::set_error(-1, __VA_ARGS__);      \
  } while (0)
