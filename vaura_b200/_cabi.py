"""ctypes binding of include/vaura_b200.h.  The library is mandatory: importing the compute wrappers
without a built ``_lib/libvaura_b200.so`` raises, there is no fallback implementation."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VAURA_B200_LIB: another build of the same library (A/B measurements of compile-time variants, profiles/scripts/)
LIB_PATH = os.environ.get("VAURA_B200_LIB") or os.path.join(HERE, "_lib", "libvaura_b200.so")

EXPORTS = [
    "vaura_version", "vaura_arch", "vaura_last_error", "vaura_launch_count", "vaura_gemv_bf16w", "vaura_linear_bf16",
    "vaura_sampler_create", "vaura_sampler_destroy", "vaura_sampler_cond_project",
    "vaura_sampler_workspace_bytes", "vaura_sampler_generate", "vaura_sampler_last_loop_ms", "vaura_sampler_forward",
    "vaura_sample_logits",
    "vaura_codec_create", "vaura_codec_destroy", "vaura_codec_workspace_bytes", "vaura_codec_decode",
    "vaura_codec_encoder_create", "vaura_codec_encoder_destroy", "vaura_codec_encoder_workspace_bytes", "vaura_codec_encode",
    "vaura_avclip_create", "vaura_avclip_destroy", "vaura_avclip_workspace_bytes", "vaura_avclip_forward",
]

PRECISION_AUTO, PRECISION_FP32ACT, PRECISION_BF16 = 0, 1, 2
KV_F32, KV_BF16 = 0, 1


class SamplerDimsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "num_layers", "d_model", "nhead", "ffn_dim", "vocab", "num_codebooks", "block_size", "cond_dim",
        "cond_in", "cond_tokens", "audio_tokens_per_video_frame")] + [("norm_eps", C.c_float)]


class SamplerWeightsC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "wqkv", "wo", "w13", "w2", "w_heads", "attn_norm", "ffn_norm", "final_norm", "tok_tables", "rope",
        "fc1", "fc2", "empty_video_emb", "wstream")]


class KvCacheC(C.Structure):
    _fields_ = [("pages", C.c_void_p), ("page_table", C.c_void_p), ("num_pages", C.c_int32),
                ("page_size", C.c_int32), ("max_pages_per_seq", C.c_int32), ("dtype", C.c_int32)]


class GenerateParamsC(C.Structure):
    _fields_ = [("batch", C.c_int32), ("use_cfg", C.c_int32), ("timesteps", C.c_int32),
                ("start_offset", C.c_int32), ("end_offset", C.c_int32), ("use_sampling", C.c_int32),
                ("temp", C.c_float), ("top_k", C.c_int32), ("top_p", C.c_float), ("cfg_scale", C.c_float),
                ("seed", C.c_uint64), ("clip_ids", C.c_void_p), ("sequence", C.c_void_p),
                ("cond_rows", C.c_void_p), ("logits_out", C.c_void_p), ("precision", C.c_int32),
                ("stream_id", C.c_uint32)]


class CodecDimsC(C.Structure):
    _fields_ = [("latent_dim", C.c_int32), ("decoder_dim", C.c_int32), ("n_blocks", C.c_int32),
                ("rates", C.c_int32 * 8), ("n_codebooks", C.c_int32), ("codebook_size", C.c_int32)]


class CodecWeightsC(C.Structure):
    _fields_ = [("blob", C.c_void_p), ("offsets", C.POINTER(C.c_int64)), ("n_offsets", C.c_int32)]


class AvclipDimsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("embed_dim", "depth", "num_heads", "mlp_ratio", "img_size", "patch_size",
                                         "in_chans", "frames", "tubelet")]


AvclipWeightsC = CodecWeightsC  # same layout: blob + host offsets table

_lib = None


def load():
    """dlopen the in-tree library and declare signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m vaura_b200.build` (nvcc, sm_100a). "
            "vaura_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.vaura_version.restype = C.c_int
    lib.vaura_arch.restype = C.c_char_p
    lib.vaura_last_error.restype = C.c_char_p
    lib.vaura_launch_count.restype = C.c_ulonglong
    lib.vaura_gemv_bf16w.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.vaura_linear_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p]
    lib.vaura_sampler_create.argtypes = [C.POINTER(SamplerDimsC), C.POINTER(SamplerWeightsC), C.POINTER(C.c_void_p)]
    lib.vaura_sampler_destroy.argtypes = [C.c_void_p]
    lib.vaura_sampler_destroy.restype = None
    lib.vaura_sampler_cond_project.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.vaura_sampler_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.vaura_sampler_workspace_bytes.restype = C.c_size_t
    lib.vaura_sampler_generate.argtypes = [C.c_void_p, C.POINTER(GenerateParamsC), C.POINTER(KvCacheC), C.c_void_p,
                                           C.c_size_t, C.c_void_p]
    lib.vaura_sampler_last_loop_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.vaura_sampler_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                          C.POINTER(KvCacheC), C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vaura_sample_logits.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                        C.c_float, C.c_int32, C.c_float, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    lib.vaura_codec_create.argtypes = [C.POINTER(CodecDimsC), C.POINTER(CodecWeightsC), C.POINTER(C.c_void_p)]
    lib.vaura_codec_destroy.argtypes = [C.c_void_p]
    lib.vaura_codec_destroy.restype = None
    lib.vaura_codec_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.vaura_codec_workspace_bytes.restype = C.c_size_t
    lib.vaura_codec_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]
    lib.vaura_codec_encoder_create.argtypes = [C.POINTER(CodecDimsC), C.c_int32, C.c_int32, C.POINTER(CodecWeightsC),
                                               C.POINTER(C.c_void_p)]
    lib.vaura_codec_encoder_destroy.argtypes = [C.c_void_p]
    lib.vaura_codec_encoder_destroy.restype = None
    lib.vaura_codec_encoder_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.vaura_codec_encoder_workspace_bytes.restype = C.c_size_t
    lib.vaura_codec_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]
    lib.vaura_avclip_create.argtypes = [C.POINTER(AvclipDimsC), C.POINTER(AvclipWeightsC), C.POINTER(C.c_void_p)]
    lib.vaura_avclip_destroy.argtypes = [C.c_void_p]
    lib.vaura_avclip_destroy.restype = None
    lib.vaura_avclip_workspace_bytes.argtypes = [C.c_void_p, C.c_int32]
    lib.vaura_avclip_workspace_bytes.restype = C.c_size_t
    lib.vaura_avclip_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError here = the .so is stale w.r.t. the header
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vaura_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
