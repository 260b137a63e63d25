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

// general form: elem_bytes 2 (bf16) or 4 (fp32); swizzle_bytes 128 or 64 (box_inner * elem_bytes must fit the swizzle span)
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t rows, uint64_t ld,
                 uint32_t box_inner, uint32_t box_rows, int swizzle_bytes);

// small-group time attention (attention_time.cu), dispatched from lv_group_attn_fwd / lv_group_attn_bwd (mode 1, T <= 16)
int time_attn_small_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_out, float* lse, int B, int H, int T, int n,
                        cudaStream_t st);
int time_attn_small_bwd(const void* qkv, long long ld_qkv, const void* out, long long ld_out, const float* lse, const void* dout,
                        long long ld_dout, void* dqkv, long long ld_dqkv, float* dcls_kv, int B, int H, int T, int n,
                        cudaStream_t st);

// key-tiled attention for groups of more than 208 keys (attention_big.cu), dispatched from lv_group_attn_fwd / _bwd
int big_group_attn_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_out, float* lse, int mode, int B, int H, int T,
                       int n, cudaStream_t st);
int big_group_attn_bwd(const void* qkv, long long ld_qkv, const void* out, long long ld_out, const float* lse, const void* dout,
                       long long ld_dout, void* dqkv, long long ld_dqkv, float* dcls_kv, int mode, int B, int H, int T, int n,
                       cudaStream_t st);

// key-tiled attention forward on tcgen05 (flash_tc.cu): q / k / v point at head 0 of batch element 0; group g of batch b owns
// query rows b*q_rows + q_grp_row0 + g*q_grp_stride .. +Lq and key rows b*kv_rows + kv_grp_row0 + g*kv_grp_stride .. +Lk
// (+ the extra key / value row b*kv_rows when has_extra); lse (optional) is [row][H].
int flash_attn_fwd_tc(const void* q, long long ld_q, long long q_rows, int q_cols, const void* k, const void* v, long long ld_kv,
                      long long kv_rows, int kv_cols, int kv_head_stride, void* out, long long ld_out, float* lse, int B, int G,
                      int H, int Lq, int Lk, long long q_grp_row0, long long q_grp_stride, long long kv_grp_row0,
                      long long kv_grp_stride, int has_extra, int causal, float scale, cudaStream_t st);
bool flash_tc_enabled();   // LAVILA_B200_FLASH_TC=0 keeps the mma.sync kernels (A/B runs)

}  // namespace lv

#define LV_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return ::lv::set_error(-1, __VA_ARGS__);      \
  } while (0)
