// Library runtime: error reporting, launch accounting, device queries, TMA descriptor encoding.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "../../include/lavila_b200.h"
#include "host_common.h"

namespace lv {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

int sm_count() {
  static int cached[64];
  static std::once_flag once[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::call_once(once[dev], [dev]() {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    // LAVILA_B200_RESERVE_SMS=k: the persistent kernels (one CTA per SM, all of its shared memory) size their grids for n - k
    // SMs, leaving k SMs to concurrent work of other streams -- NCCL's all-reduce CTAs under DistributedDataParallel cannot
    // co-reside with a 227 KB CTA and otherwise only run between our kernels.  Even count (CTA pairs).  Default 0.
    if (const char* e = std::getenv("LAVILA_B200_RESERVE_SMS")) {
      int k = atoi(e);
      k = (k / 2) * 2;
      if (k > 0 && k < n / 2) n -= k;
    }
    cached[dev] = n;
  });
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t rows, uint64_t ld,
                 uint32_t box_inner, uint32_t box_rows, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(-2, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if (elem_bytes != 2 && elem_bytes != 4) return set_error(-1, "make_tmap_2d: element size %d unsupported", elem_bytes);
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * elem_bytes) & 15))
    return set_error(-1, "TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (ld=%llu)",
                     (unsigned long long)ld);
  if (swizzle_bytes != 128 && swizzle_bytes != 64) return set_error(-1, "make_tmap_2d: swizzle %d unsupported", swizzle_bytes);
  if ((uint64_t)box_inner * elem_bytes > (uint64_t)swizzle_bytes)
    return set_error(-1, "make_tmap_2d: box of %u x %d bytes exceeds the %d-byte swizzle span", box_inner, elem_bytes, swizzle_bytes);
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-3, "cuTensorMapEncodeTiled failed (%d) inner=%llu rows=%llu ld=%llu box=%ux%u", (int)r,
                     (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)ld, box_inner, box_rows);
  return 0;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t ld,
                      uint32_t box_inner, uint32_t box_rows) {
  return make_tmap_2d(out, base, 2, inner, rows, ld, box_inner, box_rows, 128);
}

}  // namespace lv

extern "C" int lv_gemm_skinny_splits(int64_t M, int64_t N, int64_t K);

extern "C" int64_t lv_workspace_bytes(int op, int64_t a, int64_t b, int64_t c) {
  if (a <= 0 || b < 0 || c < 0) return -1;
  switch (op) {
    case LV_WS_GEMM_SKINNY: {                       // (M, N, K) -> fp32 [splits][M][N]
      const int s = lv_gemm_skinny_splits(a, b, c);
      return s > 0 ? (int64_t)s * a * b * 4 : -1;
    }
    case LV_WS_ATTN_BWD_DCLS:                       // (B, H) -> fp32 [B][H][2][64], zeroed by the caller
      return b > 0 ? a * b * 2 * 64 * 4 : -1;
    case LV_WS_CLIP_LOSS:                           // (Ng) -> lse_img[Ng] + lse_txt[Ng] + partial[2 Ng] fp32 (+ 4 control words, result[6])
      return a * 4 * 4 + 4 * 4 + 6 * 4;
    case LV_WS_P2P_BLOCK:                           // (Bl, E) -> one rank's symmetric block: 2 slots x (Bl x 2E fp32 + 32 words)
      return b > 0 ? 2 * (a * 2 * b + 32) * 4 : -1;
    default:
      return -1;
  }
}

extern "C" {
int lv_version(void) { return LV_ABI_VERSION; }
const char* lv_last_error(void) { return lv::g_err; }
int64_t lv_launch_count(void) { return lv::g_launches.load(); }
}
