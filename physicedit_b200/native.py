"""ctypes binding of libpe_b200.so (the C ABI declared in include/pe_b200.h).

PyTorch is used only for device memory and streams: every function here takes torch CUDA tensors,
checks dtype / contiguity, and passes raw device pointers + sizes + the current stream to the
library.  There is no fallback: if the shared library is missing or the device is not sm_100 the
import of the native path fails loudly (`NativeUnavailable`).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_uint, c_void_p
from typing import Optional, Sequence

import torch

_LIB_PATH = os.environ.get("PE_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libpe_b200.so")

PE_OK = 0
EPI_BIAS = 0
EPI_BIAS_GELU_SIGMOID = 2
EPI_BIAS_GELU_ERF = 3
EPI_GATE_RESIDUAL = 4
EPI_QKV_NORM_ROPE = 5
EPI_BIAS_SILU = 6
EPI_F32 = 7
EPI_ATTN_P, EPI_ATTN_DS = 8, 9
GEMM_FLAG_CTA_PAIR = 1
GEMM_FLAG_TRIM_N = 2
CONV_FLAG_CTA_PAIR = 1
CONV_FLAG_SINGLE_PATCH = 2
ATTN_FLAG_SINGLE_Q_TILE = 1
ATTN_FLAG_P_VIA_SMEM = 2
ATTN_FLAG_SPLIT_ROW_SOFTMAX = 8
ATTN_FLAG_KV64 = 16
ATTN_FLAG_HALF_ROW = 32

EXPORTED_SYMBOLS = [
    "pe_abi_version", "pe_create", "pe_destroy", "pe_last_error", "pe_check_async_error", "pe_sm_count", "pe_workspace",
    "pe_gemm", "pe_attention_fwd", "pe_attention_fwd_routed", "pe_small_attention", "pe_layernorm_modulate", "pe_layernorm_modulate2", "pe_layernorm", "pe_add_rows",
    "pe_rmsnorm", "pe_gemv", "pe_act", "pe_timestep_embedding", "pe_patchify", "pe_unpatchify", "pe_cfg_euler_step",
    "pe_special_gather", "pe_special_blend_scatter",
    "pe_conv2d", "pe_channel_rmsnorm", "pe_upsample2x", "pe_space_to_depth", "pe_nchw_to_nhwc", "pe_nhwc_to_nchw", "pe_transpose",
    "pe_softmax_rows", "pe_softmax_rows_masked", "pe_gemv_swiglu", "pe_decode_attention_fused", "pe_attention_bwd_delta", "pe_gemm_batched", "pe_attention_fwd_lse",
    "pe_gemv_fused", "pe_swiglu", "pe_rope_half", "pe_range_attention", "pe_gather_rows", "pe_argmax", "pe_kv_append", "pe_rope_kv_append", "pe_advance",
]


class NativeUnavailable(RuntimeError):
    pass


class NativeError(RuntimeError):
    pass


class GemmSeg(Structure):
    """Mirror of `pe_gemm_seg` (include/pe_b200.h)."""
    _fields_ = [
        ("a", c_void_p), ("lda", c_int64), ("w", c_void_p), ("bias", c_void_p), ("out", c_void_p), ("ldo", c_int64),
        ("M", c_int32), ("_pad0", c_int32), ("gate", c_void_p), ("out_k", c_void_p), ("out_v", c_void_p),
        ("norm_q_w", c_void_p), ("norm_k_w", c_void_p), ("rope", c_void_p),
        ("q_route", c_void_p * 8), ("k_route", c_void_p * 8), ("v_route", c_void_p * 8), ("route_ranks", c_int32), ("_pad1", c_int32),
    ]


class DecodeReq(Structure):
    """Mirror of `pe_decode_req` (include/pe_b200.h)."""
    _fields_ = [("qkv", c_void_p), ("cache_k", c_void_p), ("cache_v", c_void_p), ("out", c_void_p), ("counters", c_void_p), ("cache_rows", c_int64)]


class GemmBatch(Structure):
    """Mirror of `pe_gemm_batch` (include/pe_b200.h)."""
    _fields_ = [("batch", c_int32), ("vec_per_column", c_int32), ("a_batch_rows", c_int64), ("w_batch_rows", c_int64), ("out_batch_rows", c_int64),
                ("vec", c_void_p), ("vec_batch_stride", c_int64), ("alpha", c_float), ("_pad", c_int32)]


class Conv2dDesc(Structure):
    """Mirror of `pe_conv2d_desc` (include/pe_b200.h)."""
    _fields_ = [
        ("x", c_void_p), ("ldx", c_int64), ("w", c_void_p), ("bias", c_void_p), ("out", c_void_p), ("ldo", c_int64), ("gate", c_void_p),
        ("H", c_int32), ("W", c_int32), ("C", c_int32), ("N", c_int32), ("kh", c_int32), ("kw", c_int32), ("pad", c_int32), ("flags", c_int32),
    ]


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    path = path or _LIB_PATH
    if not os.path.exists(path):
        raise NativeUnavailable(
            f"{path} not found: build it with physicedit_b200/csrc/build.sh (or __graft_entry__.build()); "
            "there is no CPU / PyTorch fallback for the hot path")
    lib = ctypes.CDLL(path)
    lib.pe_last_error.restype = c_char_p
    lib.pe_last_error.argtypes = [c_void_p]
    lib.pe_create.argtypes = [POINTER(c_void_p), c_int]
    lib.pe_destroy.argtypes = [c_void_p]
    lib.pe_sm_count.argtypes = [c_void_p]
    lib.pe_workspace.argtypes = [c_void_p, POINTER(c_void_p), POINTER(ctypes.c_size_t)]
    lib.pe_check_async_error.argtypes = [c_void_p, c_void_p, POINTER(c_uint)]
    lib.pe_gemm.argtypes = [c_void_p, POINTER(GemmSeg), c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.pe_attention_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_float, c_int, c_void_p]
    lib.pe_attention_fwd_routed.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_float, c_int, c_int, POINTER(c_int32),
                                            POINTER(c_void_p), c_int64, c_void_p]
    lib.pe_small_attention.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_int64, c_int64, c_int64, c_float, c_void_p]
    lib.pe_layernorm_modulate.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.pe_layernorm_modulate2.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pe_layernorm.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p]
    lib.pe_add_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]
    lib.pe_rmsnorm.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p]
    lib.pe_act.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]
    lib.pe_gemv.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.pe_timestep_embedding.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    lib.pe_patchify.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.pe_unpatchify.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]
    lib.pe_cfg_euler_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]
    lib.pe_special_gather.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]
    lib.pe_special_blend_scatter.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p]
    lib.pe_conv2d.argtypes = [c_void_p, POINTER(Conv2dDesc), c_int, c_void_p]
    lib.pe_channel_rmsnorm.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p]
    lib.pe_upsample2x.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.pe_space_to_depth.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.pe_nchw_to_nhwc.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    lib.pe_nhwc_to_nchw.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    lib.pe_transpose.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]
    lib.pe_softmax_rows.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p]
    lib.pe_decode_attention_fused.argtypes = [c_void_p, POINTER(DecodeReq), c_int, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_float, c_void_p]
    lib.pe_gemv_swiglu.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_void_p]
    lib.pe_softmax_rows_masked.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p, c_int64, c_int, c_void_p]
    lib.pe_gemm_batched.argtypes = [c_void_p, POINTER(GemmSeg), POINTER(GemmBatch), c_int, c_int, c_int, c_int, c_void_p]
    lib.pe_attention_fwd_lse.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_float, c_int, c_void_p, c_void_p]
    lib.pe_attention_bwd_delta.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p]
    lib.pe_gemv_fused.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p]
    lib.pe_swiglu.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]
    lib.pe_rope_half.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.pe_range_attention.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int64,
                                       c_float, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pe_gather_rows.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]
    lib.pe_argmax.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.pe_kv_append.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]
    lib.pe_advance.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    lib.pe_rope_kv_append.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    return lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _bf16(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise NativeError(f"{name}: expected a CUDA bfloat16 tensor, got {t.dtype} on {t.device}")
    if t.stride(-1) != 1:
        raise NativeError(f"{name}: innermost stride must be 1")
    return t


class Native:
    """One handle per device.  All methods enqueue on torch's current stream and return."""

    _instances: dict = {}

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise NativeUnavailable("no CUDA device: the physicedit_b200 hot path has no CPU fallback")
        self.lib = load_library()
        self.device = int(device)
        h = c_void_p()
        rc = self.lib.pe_create(byref(h), self.device)
        if rc != PE_OK:
            raise NativeUnavailable(f"pe_create(device={device}) failed with {rc} (needs an sm_100 GPU)")
        self.h = h
        self.sm_count = self.lib.pe_sm_count(self.h)
        self.launches = 0      # number of kernels of this library enqueued (bench.py reports it)
        self.prof = None       # {tag: [(start_event, end_event), ...]} when per-launch CUDA-event timing is on
        self.tag = None        # set by the engine before a launch to name it in `prof`

    @classmethod
    def get(cls, device: int = 0) -> "Native":
        dev = int(device)
        if dev not in cls._instances:
            cls._instances[dev] = Native(dev)
        return cls._instances[dev]

    # -------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check(self, rc: int, what: str) -> None:
        if rc != PE_OK:
            raise NativeError(f"{what} failed ({rc}): {self.lib.pe_last_error(self.h).decode()}")
        if self.prof is not None and self._ev0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.device))
            self.prof.setdefault(self.tag or what, []).append((self._ev0, e1))
            self._ev0 = None
        self.tag = None

    _ev0 = None

    def _stream_prof(self) -> int:
        """Stream handle for a launch; when profiling, first records the start event on that stream."""
        st = torch.cuda.current_stream(self.device)
        if self.prof is not None:
            self._ev0 = torch.cuda.Event(enable_timing=True)
            self._ev0.record(st)
        return st.cuda_stream

    def profile_summary(self) -> dict:
        """{tag: (launches, total_ms)} from the recorded event pairs (call after a synchronize)."""
        out = {}
        for tag, evs in (self.prof or {}).items():
            out[tag] = (len(evs), sum(a.elapsed_time(b) for a, b in evs))
        return out

    def workspace_read(self, n_int64: int) -> torch.Tensor:
        """Copy of the first n_int64 words of the handle's diagnostic scratch (trace builds write there)."""
        ptr, nbytes = c_void_p(), ctypes.c_size_t()
        self._check(self.lib.pe_workspace(self.h, byref(ptr), byref(nbytes)), "pe_workspace")
        out = torch.empty(n_int64, dtype=torch.int64)
        torch.cuda.synchronize(self.device)
        ctypes.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(c_void_p(out.data_ptr()), ptr, ctypes.c_size_t(n_int64 * 8), c_int(2))
        return out

    def check_async(self) -> None:
        diag = c_uint(0)
        self._check(self.lib.pe_check_async_error(self.h, self._stream(), byref(diag)), "pe_check_async_error")

    # -------------------------------------------------------------------------------------------
    def gemm(self, segs: Sequence[dict], N: int, K: int, epilogue: int = EPI_BIAS, flags: int = 0) -> None:
        """segs: dicts with a [M,K], w [N,K], bias [N]|None, out [M,*] and the epilogue extras."""
        arr = (GemmSeg * len(segs))()
        for i, s in enumerate(segs):
            a = _bf16(s["a"], "a")
            out = s["out"]
            if epilogue == EPI_F32:
                if out.dtype != torch.float32 or not out.is_cuda or out.stride(-1) != 1:
                    raise NativeError("out: EPI_F32 needs a CUDA float32 tensor with innermost stride 1")
            else:
                _bf16(out, "out")
            w = _bf16(s["w"], "w")
            if not w.is_contiguous():
                raise NativeError("w must be contiguous [N, K]")
            g = arr[i]
            g.a, g.lda, g.w, g.bias = a.data_ptr(), a.stride(0), w.data_ptr(), _ptr(s.get("bias"))
            g.out, g.ldo, g.M = out.data_ptr(), out.stride(0), a.shape[0]
            g.gate, g.out_k, g.out_v = _ptr(s.get("gate")), _ptr(s.get("out_k")), _ptr(s.get("out_v"))
            g.norm_q_w, g.norm_k_w, g.rope = _ptr(s.get("norm_q_w")), _ptr(s.get("norm_k_w")), _ptr(s.get("rope"))
            routes = s.get("routes")                   # head-parallel QKV: (q_ptrs, k_ptrs, v_ptrs), raw device addresses (may be peer memory)
            if routes:
                g.route_ranks = len(routes[0])
                for r, (qp, kp, vp) in enumerate(zip(*routes)):
                    g.q_route[r], g.k_route[r], g.v_route[r] = qp, kp, vp
        self._check(self.lib.pe_gemm(self.h, arr, len(segs), N, K, epilogue, flags, self._stream_prof()), "pe_gemm")
        self.launches += 1

    def linear(self, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int = EPI_BIAS, flags: int = 0) -> torch.Tensor:
        out = torch.empty((x.shape[0], w.shape[0]), dtype=torch.bfloat16, device=x.device)
        self.gemm([dict(a=x, w=w, bias=bias, out=out)], w.shape[0], w.shape[1], epilogue, flags)
        return out

    def attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, H: int, scale: float, flags: int = 0) -> None:
        """q, k, v, o: [S, >= H*128] token-major with a common row stride."""
        for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o")):
            _bf16(t, n)
            if t.stride(0) != q.stride(0):
                raise NativeError("q, k, v, o must share one row stride")
        self._check(self.lib.pe_attention_fwd(self.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), q.shape[0], H,
                                              q.stride(0), scale, flags, self._stream_prof()), "pe_attention_fwd")
        self.launches += 1

    def attention_routed(self, q, k, v, H: int, scale: float, route_end: Sequence[int], o_ptrs: Sequence[int], ldo: int, flags: int = 0) -> None:
        """Joint attention over q / k / v [S, >= H*128] whose output rows are written to o_ptrs[i] + row * ldo + head * 128 for rows below
        route_end[i] (first match); the pointers may be peer-GPU memory (sequence-parallel mode)."""
        for t, n in ((q, "q"), (k, "k"), (v, "v")):
            _bf16(t, n)
            if t.stride(0) != q.stride(0):
                raise NativeError("q, k, v must share one row stride")
        n = len(route_end)
        ends = (c_int32 * n)(*[int(e) for e in route_end])
        ptrs = (c_void_p * n)(*[int(p) for p in o_ptrs])
        self._check(self.lib.pe_attention_fwd_routed(self.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), q.shape[0], H, q.stride(0), scale, flags, n, ends, ptrs,
                                                     ldo, self._stream_prof()), "pe_attention_fwd_routed")
        self.launches += 1

    def small_attention(self, q, k, v, o, B: int, H: int, Sq: int, Skv: int, D: int, scale: float) -> None:
        """q: [B*Sq, ldq], k/v: [B*Skv, ldkv] (same stride), o: [B*Sq, ldo]; head h lives at columns [h*D, (h+1)*D)."""
        for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o")):
            _bf16(t, n)
        if k.stride(0) != v.stride(0):
            raise NativeError("k and v must share one row stride")
        self._check(self.lib.pe_small_attention(self.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, Sq, Skv, D,
                                                q.stride(0), k.stride(0), o.stride(0), scale, self._stream_prof()), "pe_small_attention")
        self.launches += 1

    def layernorm_modulate(self, x, out, shift, one_plus_scale) -> None:
        _bf16(x, "x"); _bf16(out, "out")
        if not (x.is_contiguous() and out.is_contiguous()):
            raise NativeError("layernorm_modulate: x / out must be contiguous")
        self._check(self.lib.pe_layernorm_modulate(self.h, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], shift.data_ptr(),
                                                   one_plus_scale.data_ptr(), self._stream_prof()), "pe_layernorm_modulate")
        self.launches += 1

    def layernorm_modulate2(self, x, out, split_row: int, shift0, ops0, shift1, ops1) -> None:
        _bf16(x, "x"); _bf16(out, "out")
        if not (x.is_contiguous() and out.is_contiguous()):
            raise NativeError("layernorm_modulate2: x / out must be contiguous")
        self._check(self.lib.pe_layernorm_modulate2(self.h, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], split_row, shift0.data_ptr(),
                                                    ops0.data_ptr(), shift1.data_ptr(), ops1.data_ptr(), self._stream_prof()), "pe_layernorm_modulate2")
        self.launches += 1

    def layernorm(self, x, out, w=None, b=None, eps: float = 1e-5) -> None:
        _bf16(x, "x"); _bf16(out, "out")
        self._check(self.lib.pe_layernorm(self.h, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _ptr(w), _ptr(b), eps,
                                          self._stream_prof()), "pe_layernorm")
        self.launches += 1

    def add_rows(self, x, add, period: int, alpha: float = 1.0) -> None:
        _bf16(x, "x"); _bf16(add, "add")
        self._check(self.lib.pe_add_rows(self.h, x.data_ptr(), add.data_ptr(), x.shape[0], x.shape[1], period, alpha, self._stream_prof()),
                    "pe_add_rows")
        self.launches += 1

    def rmsnorm(self, x, out, w, eps: float = 1e-6) -> None:
        _bf16(x, "x"); _bf16(out, "out")
        self._check(self.lib.pe_rmsnorm(self.h, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _ptr(w), eps, self._stream_prof()),
                    "pe_rmsnorm")
        self.launches += 1

    def gemv(self, x, w, bias, y, act_in: int = 0, act_out: int = 0, one_plus_mask=None) -> None:
        """x [batch, K], w [N, K], y [batch, N]."""
        _bf16(x, "x"); _bf16(w, "w"); _bf16(y, "y")
        batch = 1 if x.dim() == 1 else x.shape[0]
        self._check(self.lib.pe_gemv(self.h, x.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), batch, w.shape[0], w.shape[1],
                                     act_in, act_out, _ptr(one_plus_mask), self._stream_prof()), "pe_gemv")
        self.launches += 1

    def act(self, x, y, act: int = 1) -> None:
        """y = bf16(act(x)) element-wise (1 = SiLU)."""
        _bf16(x, "x"); _bf16(y, "y")
        if not (x.is_contiguous() and y.is_contiguous()) or x.numel() != y.numel():
            raise NativeError("act: x / y must be contiguous and of equal size")
        self._check(self.lib.pe_act(self.h, x.data_ptr(), y.data_ptr(), x.numel(), act, self._stream_prof()), "pe_act")
        self.launches += 1

    def timestep_embedding(self, t_in, out, raw: bool = True) -> None:
        _bf16(t_in, "t_in"); _bf16(out, "out")
        self._check(self.lib.pe_timestep_embedding(self.h, t_in.data_ptr(), out.data_ptr(), int(raw), self._stream_prof()), "pe_timestep_embedding")
        self.launches += 1

    def patchify(self, latents, tokens) -> None:
        """latents [16, H8, W8] contiguous -> tokens [(H8/2)*(W8/2), 64] contiguous."""
        _bf16(latents, "latents"); _bf16(tokens, "tokens")
        self._check(self.lib.pe_patchify(self.h, latents.data_ptr(), tokens.data_ptr(), latents.shape[-2], latents.shape[-1],
                                         self._stream_prof()), "pe_patchify")
        self.launches += 1

    def unpatchify(self, tokens, latents) -> None:
        _bf16(latents, "latents"); _bf16(tokens, "tokens")
        self._check(self.lib.pe_unpatchify(self.h, tokens.data_ptr(), tokens.stride(0), latents.data_ptr(), latents.shape[-2],
                                           latents.shape[-1], self._stream_prof()), "pe_unpatchify")
        self.launches += 1

    def cfg_euler_step(self, latents, posi, nega, cfg_scale: float, dsigma: float) -> None:
        _bf16(latents, "latents"); _bf16(posi, "posi")
        self._check(self.lib.pe_cfg_euler_step(self.h, latents.data_ptr(), posi.data_ptr(), _ptr(nega), latents.numel(), cfg_scale,
                                               dsigma, self._stream_prof()), "pe_cfg_euler_step")
        self.launches += 1

    def special_gather(self, prompt_emb, mask_u8, dst, idx) -> None:
        _bf16(prompt_emb, "prompt_emb"); _bf16(dst, "dst")
        self._check(self.lib.pe_special_gather(self.h, prompt_emb.data_ptr(), mask_u8.data_ptr(), prompt_emb.shape[0], prompt_emb.shape[1],
                                               dst.data_ptr(), idx.data_ptr(), dst.shape[0], self._stream_prof()), "pe_special_gather")
        self.launches += 2

    def special_blend_scatter(self, prompt_emb, idx, pred_dino, pred_vae, t_in, t_min: float, t_max: float) -> None:
        _bf16(prompt_emb, "prompt_emb")
        self._check(self.lib.pe_special_blend_scatter(self.h, prompt_emb.data_ptr(), idx.data_ptr(), pred_dino.shape[0], prompt_emb.shape[1],
                                                      pred_dino.data_ptr(), pred_vae.data_ptr(), t_in.data_ptr(), t_min, t_max,
                                                      self._stream_prof()), "pe_special_blend_scatter")
        self.launches += 1

    # ---- QwenImageVAE path (include/pe_b200.h, last section) -----------------------------------------------------------------
    def conv2d(self, x, H: int, W: int, C: int, w, bias, out, N: int, kh: int, kw: int, pad: int, epilogue: int = EPI_BIAS, gate=None,
               flags: int = 0) -> None:
        """x: [H*W, >=C] channels-last map, w: [N, kh*kw*round_up(C,64)], out: [H*W, >=N]; stride 1, same output size."""
        _bf16(x, "x"); _bf16(w, "w"); _bf16(out, "out")
        if x.shape[0] != H * W or out.shape[0] != H * W or not w.is_contiguous():
            raise NativeError("conv2d: x / out must have H*W rows and w must be contiguous")
        cpad = (C + 63) // 64 * 64
        if w.shape[0] != N or w.shape[1] != kh * kw * cpad:
            raise NativeError(f"conv2d: w must be [{N}, {kh * kw * cpad}], got {tuple(w.shape)}")
        d = Conv2dDesc(x=x.data_ptr(), ldx=x.stride(0), w=w.data_ptr(), bias=_ptr(bias), out=out.data_ptr(), ldo=out.stride(0), gate=_ptr(gate),
                       H=H, W=W, C=C, N=N, kh=kh, kw=kw, pad=pad, flags=flags)
        self._check(self.lib.pe_conv2d(self.h, byref(d), epilogue, self._stream_prof()), "pe_conv2d")
        self.launches += 1

    def channel_rmsnorm(self, x, out, C: int, gamma, act: bool) -> None:
        _bf16(x, "x"); _bf16(out, "out"); _bf16(gamma, "gamma")
        self._check(self.lib.pe_channel_rmsnorm(self.h, x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], C, gamma.data_ptr(),
                                                int(act), self._stream_prof()), "pe_channel_rmsnorm")
        self.launches += 1

    def upsample2x(self, x, out, H: int, W: int, C: int) -> None:
        _bf16(x, "x"); _bf16(out, "out")
        if not (x.is_contiguous() and out.is_contiguous()) or x.numel() != H * W * C or out.numel() != 4 * H * W * C:
            raise NativeError("upsample2x: x [H*W, C] and out [4*H*W, C] must be contiguous")
        self._check(self.lib.pe_upsample2x(self.h, x.data_ptr(), out.data_ptr(), H, W, C, self._stream_prof()), "pe_upsample2x")
        self.launches += 1

    def space_to_depth(self, x, out, H: int, W: int, C: int) -> None:
        _bf16(x, "x"); _bf16(out, "out")
        if not (x.is_contiguous() and out.is_contiguous()) or x.numel() != H * W * C or out.numel() != H * W * C:
            raise NativeError("space_to_depth: x [H*W, C] and out [H*W/4, 4C] must be contiguous")
        self._check(self.lib.pe_space_to_depth(self.h, x.data_ptr(), out.data_ptr(), H, W, C, self._stream_prof()), "pe_space_to_depth")
        self.launches += 1

    def nchw_to_nhwc(self, src, dst, C: int, op: int = 0, p0=None, p1=None) -> None:
        """src [C, H, W] contiguous -> dst [H*W, >=C] (columns >= C untouched)."""
        _bf16(src, "src"); _bf16(dst, "dst")
        if not src.is_contiguous() or src.numel() != C * dst.shape[0]:
            raise NativeError("nchw_to_nhwc: src must be contiguous [C, H*W]")
        self._check(self.lib.pe_nchw_to_nhwc(self.h, src.data_ptr(), dst.data_ptr(), dst.stride(0), C, dst.shape[0], op, _ptr(p0), _ptr(p1),
                                             self._stream_prof()), "pe_nchw_to_nhwc")
        self.launches += 1

    def nhwc_to_nchw(self, src, dst, C: int, op: int = 0, p0=None, p1=None) -> None:
        """src [H*W, >=C] -> dst [C, H, W] contiguous (first C channels)."""
        _bf16(src, "src"); _bf16(dst, "dst")
        if not dst.is_contiguous() or dst.numel() != C * src.shape[0]:
            raise NativeError("nhwc_to_nchw: dst must be contiguous [C, H*W]")
        self._check(self.lib.pe_nhwc_to_nchw(self.h, src.data_ptr(), src.stride(0), dst.data_ptr(), C, src.shape[0], op, _ptr(p0), _ptr(p1),
                                             self._stream_prof()), "pe_nhwc_to_nchw")
        self.launches += 1

    def transpose(self, src, dst) -> None:
        _bf16(src, "src"); _bf16(dst, "dst")
        if dst.shape[0] != src.shape[1] or dst.shape[1] != src.shape[0]:
            raise NativeError("transpose: dst must be [C, R]")
        self._check(self.lib.pe_transpose(self.h, src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), src.shape[0], src.shape[1],
                                          self._stream_prof()), "pe_transpose")
        self.launches += 1

    def softmax_rows(self, scores, probs, n: int, scale: float, mask=None) -> None:
        """probs[:, :n] = softmax(scale * scores[:, :n]); probs[:, n:] = 0.  scores fp32 [rows, >=n], probs bf16 [rows, n_pad].
        mask (optional): uint8 [period, >= n], 0 = key hidden; score row r uses mask row r % period."""
        _bf16(probs, "probs")
        if scores.dtype != torch.float32 or not scores.is_cuda or scores.stride(-1) != 1 or scores.shape[0] != probs.shape[0]:
            raise NativeError("softmax_rows: scores must be CUDA float32 [rows, >=n] with as many rows as probs")
        if mask is not None:
            if mask.dtype != torch.uint8 or not mask.is_cuda or mask.stride(-1) != 1 or mask.shape[1] < n:
                raise NativeError("softmax_rows: mask must be CUDA uint8 [period, >= n]")
            self._check(self.lib.pe_softmax_rows_masked(self.h, scores.data_ptr(), scores.stride(0), probs.data_ptr(), probs.stride(0), scores.shape[0],
                                                        n, probs.shape[1], scale, mask.data_ptr(), mask.stride(0), mask.shape[0], self._stream_prof()),
                        "pe_softmax_rows_masked")
            self.launches += 1
            return
        self._check(self.lib.pe_softmax_rows(self.h, scores.data_ptr(), scores.stride(0), probs.data_ptr(), probs.stride(0), scores.shape[0],
                                             n, probs.shape[1], scale, self._stream_prof()), "pe_softmax_rows")
        self.launches += 1

    # ---- training path: batched GEMM (one launch for the per-head products of the attention backward), forward with row statistics
    def gemm_batched(self, a, w, out, batch: int, M: int, N: int, K: int, a_batch_rows: int, w_batch_rows: int, out_batch_rows: int,
                     epilogue: int = EPI_BIAS, vec=None, vec_batch_stride: int = 0, vec_per_column: bool = False, alpha: float = 1.0, flags: int = 0) -> None:
        """`batch` products out_b [M, N] = epilogue(a_b [M, K] w_b[N, K]^T) in one launch.  a / w / out are 2-D views of the flattened operands
        (rows of problem b start at b * *_batch_rows); w must be contiguous with row length K; out bf16, or fp32 for EPI_F32."""
        _bf16(a, "a"); _bf16(w, "w")
        if epilogue == EPI_F32:
            if out.dtype != torch.float32 or not out.is_cuda or out.stride(-1) != 1:
                raise NativeError("out: EPI_F32 needs a CUDA float32 tensor with innermost stride 1")
        else:
            _bf16(out, "out")
        if w.stride(0) != K or a.stride(0) < K:
            raise NativeError("gemm_batched: w rows must be contiguous of length K, a rows at least K long")
        if epilogue in (EPI_ATTN_P, EPI_ATTN_DS) and (vec is None or vec.dtype != torch.float32 or not vec.is_cuda):
            raise NativeError("gemm_batched: the attention-backward epilogues need a CUDA float32 statistic vector")
        seg = (GemmSeg * 1)()
        g = seg[0]
        g.a, g.lda, g.w, g.bias, g.out, g.ldo, g.M = a.data_ptr(), a.stride(0), w.data_ptr(), None, out.data_ptr(), out.stride(0), M
        bt = GemmBatch(batch, int(bool(vec_per_column)), a_batch_rows, w_batch_rows, out_batch_rows, _ptr(vec), vec_batch_stride, alpha, 0)
        self._check(self.lib.pe_gemm_batched(self.h, seg, ctypes.byref(bt), N, K, epilogue, flags, self._stream_prof()), "pe_gemm_batched")
        self.launches += 1

    def attention_lse(self, q, k, v, o, lse, H: int, scale: float, flags: int = 0) -> None:
        """attention() that also writes lse [H, S] (fp32, log2 domain, scale included) for the backward."""
        for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o")):
            _bf16(t, n)
            if t.stride(0) != q.stride(0):
                raise NativeError("q, k, v, o must share one row stride")
        if lse.dtype != torch.float32 or not lse.is_cuda or not lse.is_contiguous() or lse.numel() != H * q.shape[0]:
            raise NativeError("attention_lse: lse must be a contiguous CUDA float32 [H, S]")
        self._check(self.lib.pe_attention_fwd_lse(self.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), q.shape[0], H, q.stride(0), scale, flags,
                                                  lse.data_ptr(), self._stream_prof()), "pe_attention_fwd_lse")
        self.launches += 1

    # ---- training path: attention backward helpers ------------------------------------------------------------------------------
    def attention_bwd_delta(self, d_o, o, delta, H: int) -> None:
        """delta[h, s] = sum_d dO[s, h, d] * O[s, h, d]; dO, O bf16 token-major [S, H * 128], delta fp32 [H, >= S] (row stride free)."""
        _bf16(d_o, "dO"); _bf16(o, "O")
        if delta.dtype != torch.float32 or not delta.is_cuda or delta.stride(-1) != 1 or delta.shape[0] != H or delta.shape[1] < o.shape[0]:
            raise NativeError("attention_bwd_delta: delta must be CUDA float32 [H, >= S]")
        self._check(self.lib.pe_attention_bwd_delta(self.h, d_o.data_ptr(), d_o.stride(0), o.data_ptr(), o.stride(0), o.shape[0], H, delta.data_ptr(),
                                                    delta.stride(0), self._stream_prof()), "pe_attention_bwd_delta")
        self.launches += 1

    # ---- Qwen2.5-VL text-encoder path (include/pe_b200.h, last section) --------------------------------------------------------
    def gemv_fused(self, x, w, bias, y, act_in: int = 0, norm_w=None, eps: float = 1e-6, residual=None) -> None:
        """x [batch, K] (act_in 2: [batch, 2K] gate|up), w [N, K], y [batch, N]; optional RMSNorm prologue and residual epilogue."""
        _bf16(x, "x"); _bf16(w, "w"); _bf16(y, "y")
        batch = 1 if x.dim() == 1 else x.shape[0]
        self._check(self.lib.pe_gemv_fused(self.h, x.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), batch, w.shape[0], w.shape[1], act_in,
                                           _ptr(norm_w), eps, _ptr(residual), self._stream_prof()), "pe_gemv_fused")
        self.launches += 1

    def decode_attention_fused(self, qkv_rows, caches, outs, counters, Hq: int, Hkv: int, D: int, cos, sin, scale: float) -> None:
        """One decode step's rope + KV append + attention for len(qkv_rows) <= 8 requests in one launch.  qkv_rows[i] [(Hq + 2 Hkv) * D], caches[i] =
        (cache_k, cache_v) [cap, ldc], outs[i] [Hq * D], counters[i] int32 [>= 2] = (cache rows before the append, rope row)."""
        n = len(qkv_rows)
        arr = (DecodeReq * n)()
        for i in range(n):
            ck, cv = caches[i]
            _bf16(qkv_rows[i], "qkv"); _bf16(ck, "cache_k"); _bf16(cv, "cache_v"); _bf16(outs[i], "out")
            if ck.stride(0) != cv.stride(0) or ck.stride(0) != caches[0][0].stride(0) or ck.shape != cv.shape:
                raise NativeError("decode_attention_fused: all caches must share one row stride")
            arr[i] = DecodeReq(qkv_rows[i].data_ptr(), ck.data_ptr(), cv.data_ptr(), outs[i].data_ptr(), counters[i].data_ptr(), ck.shape[0])
        self._check(self.lib.pe_decode_attention_fused(self.h, arr, n, Hq, Hkv, D, caches[0][0].stride(0), cos.data_ptr(), sin.data_ptr(), scale,
                                                       self._stream_prof()), "pe_decode_attention_fused")
        self.launches += 1

    def gemv_swiglu(self, x, w, bias, y, norm_w=None, eps: float = 1e-6) -> None:
        """x [batch, K], w [2I, K] = gate rows | up rows, y [batch, I] = silu(gate) * up with the reference's rounding points; optional RMSNorm prologue."""
        _bf16(x, "x"); _bf16(w, "w"); _bf16(y, "y")
        batch = 1 if x.dim() == 1 else x.shape[0]
        self._check(self.lib.pe_gemv_swiglu(self.h, x.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), batch, w.shape[0] // 2, w.shape[1], _ptr(norm_w), eps,
                                            self._stream_prof()), "pe_gemv_swiglu")
        self.launches += 1

    def swiglu(self, x, out, I: int) -> None:
        """x [rows, >= 2I] = gate | up, out [rows, >= I]."""
        _bf16(x, "x"); _bf16(out, "out")
        self._check(self.lib.pe_swiglu(self.h, x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], I, self._stream_prof()), "pe_swiglu")
        self.launches += 1

    def rope_half(self, x, H: int, D: int, cos, sin, row0: int = 0, row_ptr=None, mode: int = 1) -> None:
        """In place on x [T, >= H*D]; cos / sin fp32 [rows, D]."""
        _bf16(x, "x")
        for t in (cos, sin):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.shape[-1] != D:
                raise NativeError("rope_half: cos / sin must be contiguous CUDA float32 [rows, D]")
        if row_ptr is None and row0 + x.shape[0] > cos.shape[0]:
            raise NativeError("rope_half: table too short")
        self._check(self.lib.pe_rope_half(self.h, x.data_ptr(), x.stride(0), x.shape[0], H, D, cos.data_ptr(), sin.data_ptr(), _ptr(row_ptr), row0, mode,
                                          self._stream_prof()), "pe_rope_half")
        self.launches += 1

    def range_attention(self, q, k, v, o, H: int, Hkv: int, D: int, scale: float, kv_lo=None, kv_hi=None, kv_len_ptr=None, Skv: Optional[int] = None) -> None:
        """q [Sq, >= H*D], k / v [Skv, >= Hkv*D] (same stride), o [Sq, >= H*D]; kv_lo / kv_hi int32 [Sq] device tensors or None."""
        for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o")):
            _bf16(t, n)
        if k.stride(0) != v.stride(0):
            raise NativeError("k and v must share one row stride")
        for t in (kv_lo, kv_hi, kv_len_ptr):
            if t is not None and (t.dtype != torch.int32 or not t.is_cuda):
                raise NativeError("range_attention: kv_lo / kv_hi / kv_len_ptr must be CUDA int32 tensors")
        self._check(self.lib.pe_range_attention(self.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), H, Hkv, q.shape[0],
                                                k.shape[0] if Skv is None else Skv, D, q.stride(0), k.stride(0), o.stride(0), scale, _ptr(kv_lo),
                                                _ptr(kv_hi), _ptr(kv_len_ptr), self._stream_prof()), "pe_range_attention")
        self.launches += 1

    def gather_rows(self, table, ids, out) -> None:
        _bf16(table, "table"); _bf16(out, "out")
        if ids.dtype != torch.int64 or not ids.is_cuda or not ids.is_contiguous():
            raise NativeError("gather_rows: ids must be a contiguous CUDA int64 tensor")
        self._check(self.lib.pe_gather_rows(self.h, table.data_ptr(), table.stride(0), ids.data_ptr(), out.data_ptr(), out.stride(0), ids.numel(),
                                            table.shape[1], self._stream_prof()), "pe_gather_rows")
        self.launches += 1

    def argmax(self, x, out, log=None, log_pos=None) -> None:
        _bf16(x, "x")
        self._check(self.lib.pe_argmax(self.h, x.data_ptr(), x.numel(), out.data_ptr(), _ptr(log), _ptr(log_pos), self._stream_prof()), "pe_argmax")
        self.launches += 1

    def kv_append(self, k_new, v_new, cache_k, cache_v, pos) -> None:
        _bf16(k_new, "k_new"); _bf16(v_new, "v_new"); _bf16(cache_k, "cache_k"); _bf16(cache_v, "cache_v")
        self._check(self.lib.pe_kv_append(self.h, k_new.data_ptr(), v_new.data_ptr(), cache_k.data_ptr(), cache_v.data_ptr(), cache_k.stride(0),
                                          k_new.numel(), pos.data_ptr(), self._stream_prof()), "pe_kv_append")
        self.launches += 1

    def rope_kv_append(self, qkv_row, Hq: int, Hkv: int, D: int, cos, sin, cache_k, cache_v, counters) -> None:
        _bf16(qkv_row, "qkv_row"); _bf16(cache_k, "cache_k"); _bf16(cache_v, "cache_v")
        self._check(self.lib.pe_rope_kv_append(self.h, qkv_row.data_ptr(), Hq, Hkv, D, cos.data_ptr(), sin.data_ptr(), cache_k.data_ptr(), cache_v.data_ptr(),
                                               cache_k.stride(0), counters.data_ptr(), self._stream_prof()), "pe_rope_kv_append")
        self.launches += 1

    def advance(self, counters, n: int) -> None:
        self._check(self.lib.pe_advance(self.h, counters.data_ptr(), n, self._stream_prof()), "pe_advance")
        self.launches += 1
