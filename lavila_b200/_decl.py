"""argtypes for the non-GEMM entry points of include/lavila_b200.h."""
import ctypes

P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

SIGNATURES = {
    "lv_layernorm_fwd": [P, L, P, P, F, P, L, P, L, L, I, P],
    "lv_layernorm_bwd": [P, I, L, P, L, P, F, P, L, P, L, P, L, P, L, P, P, L, I, P],
    "lv_group_attn_fwd": [P, L, P, L, P, I, I, I, I, I, I, P],
    "lv_group_attn_bwd": [P, L, P, L, P, P, L, P, L, P, I, I, I, I, I, I, I, P],
    "lv_space_attn_fwd_tc": [P, L, P, L, P, I, I, I, I, P],
    "lv_space_attn_bwd_tc": [P, L, P, L, P, P, L, P, L, P, I, I, I, I, I, P],
    "lv_space_attn_fwd_tc_cls": [P, L, P, L, P, P, I, I, I, I, P],
    "lv_time_attn_fwd_cls": [P, L, P, L, P, P, I, I, I, I, P],
    "lv_time_attn_bwd_cls": [P, L, P, L, P, P, L, P, L, P, P, I, I, I, I, P],
    "lv_space_attn_bwd_tc_cls": [P, L, P, L, P, P, L, P, L, P, P, I, I, I, I, P],
    "lv_flash_attn_fwd": [P, L, L, P, P, L, L, I, P, L, I, I, I, I, I, F, P],
    "lv_flash_attn_fwd_dyn": [P, L, L, P, P, L, L, I, P, L, I, I, I, P, I, F, P],
    "lv_debug_set_buffer": [P],
    "lv_cls_attn_fwd": [P, L, P, L, P, I, I, I, P],
    "lv_cls_attn_bwd": [P, L, P, L, P, L, P, P, L, P, I, I, I, I, P],
    "lv_cls_kv_finalize": [P, P, L, I, I, I, P],
    "lv_cls_query_attn_fwd": [P, P, P, P, I, I, I, P],
    "lv_cls_query_attn_bwd": [P, P, P, P, P, P, P, I, I, I, P],
    "lv_add_rows": [P, I, L, P, I, I, P],
    "lv_cast_f32_bf16": [P, P, L, P],
    "lv_colsum_bf16": [P, L, L, I, P, P],
    "lv_patch_im2col": [P, P, I, I, I, I, I, I, L, P],
    "lv_embed_assemble": [P, P, P, P, P, I, I, I, I, P],
    "lv_embed_assemble_bwd": [P, P, P, P, P, I, I, I, I, P],
    "lv_text_embed": [P, P, P, P, L, I, I, I, P],
    "lv_text_embed_bwd": [P, P, P, P, L, I, I, I, P],
    "lv_argmax_i64": [P, P, I, I, P],
    "lv_gather_rows_f32": [P, P, P, I, I, I, I, P],
    "lv_l2norm_fwd": [P, P, P, I, I, P],
    "lv_l2norm_bwd": [P, P, P, P, I, I, P],
    "lv_clip_loss_fwd": [P, P, P, I, I, P, P, P, P, P, P],
    "lv_clip_loss_bwd": [P, P, P, P, P, P, F, F, I, I, I, I, P, P, P, P],
    "lv_clip_loss_fwd_gather": [P, P, P, I, I, I, ctypes.c_uint32, P, P, P, I, P, P, P, P, P, L, P],
    "lv_clip_loss_gather_max_rows": [I],
    "lv_top_p_filter": [P, L, I, I, F, F, P],
    "lv_clip_transform": [P, I, I, I, I, P, P, P, I, I, P],
    "lv_ssl_clip_loss_fwd": [P, P, P, P, P, I, I, P, P, P, P, P, P],
    "lv_ssl_clip_loss_bwd": [P, P, P, P, P, P, P, P, F, F, I, I, I, I, P, P, P, P],
}


def declare(lib):
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = I
        fn.argtypes = args
