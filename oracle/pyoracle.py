"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the reference's own LogitComparer
(oracle/_ref/libref_logitcomparer.so).  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by blama_b200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_GGML, MODE_F32, MODE_BF16, MODE_GGML_ALT = 0, 1, 2, 3


class TokenData(C.Structure):
    _fields_ = [("token", C.c_int32), ("logit", C.c_float)]


class Metrics(C.Structure):
    _fields_ = [("top1Match", C.c_float), ("distance", C.c_float), ("jsd", C.c_float)]


TD_DTYPE = np.dtype([("token", np.int32), ("logit", np.float32)])


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE])


def _load() -> C.CDLL:
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp, i32, i64, f32, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32
    sig = {
        "orc_model_load": (vp, [C.c_char_p]), "orc_model_free": (None, [vp]),
        "orc_n_vocab": (i32, [vp]), "orc_n_ctx_train": (i32, [vp]), "orc_n_embd": (i32, [vp]), "orc_n_layer": (i32, [vp]),
        "orc_token_bos": (i32, [vp]), "orc_is_eog": (i32, [vp, i32]), "orc_weight_bytes_per_token": (i64, [vp]),
        "orc_ctx_create": (vp, [vp, i32, i32, i32]), "orc_ctx_free": (None, [vp]), "orc_kv_clear": (None, [vp]),
        "orc_kv_shift": (i32, [vp, i32, i32]), "orc_kv_seq_add": (i32, [vp, i32, i32, i32]), "orc_kv_seq_div": (i32, [vp, i32, i32, i32]),
        "orc_next_pos": (i32, [vp]),
        "orc_n_past": (i32, [vp]), "orc_decode": (i32, [vp, vp, i32, i32]),
        "orc_get_logits": (C.POINTER(C.c_float), [vp, i32]), "orc_get_hidden": (C.POINTER(C.c_float), [vp, i32]),
        "orc_topk": (None, [vp, i32, i32, vp]), "orc_gather_sorted": (i32, [vp, i32, vp, i32, vp]),
        "orc_sampler_create": (vp, [u32, f32, f32]), "orc_sampler_create_ex": (vp, [u32, f32, f32, i32, f32, i32]),
        "orc_sampler_free": (None, [vp]), "orc_sampler_reset": (None, [vp]),
        "orc_sampler_sample": (i32, [vp, vp, i32]), "orc_sampler_sample_candidates": (i32, [vp, vp, i32]),
        "orc_session_complete": (i32, [vp, vp, i32, i32, u32, f32, f32, vp, vp]),
        "orc_session_fill_ctx": (i32, [vp, vp, i32, vp, i32, vp, vp, vp, vp]),
        "orc_lc_compare": (Metrics, [vp, i32, vp, i32]), "orc_lc_similarity": (f32, [vp, i32, vp, i32]),
        "orc_lc_score": (f32, [vp, i32]),
        "orc_dequantize": (i32, [i32, vp, i64, vp]), "orc_matvec": (i32, [i32, vp, i64, i64, vp, vp, i32]),
        "orc_quantize_q8_K": (i32, [vp, i64, vp, vp, vp]), "orc_quantize_q8_0": (i32, [vp, i64, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def as_td(pairs) -> np.ndarray:
    """[(id, logit), ...] or structured array -> contiguous TokenData array"""
    if isinstance(pairs, np.ndarray) and pairs.dtype == TD_DTYPE:
        return np.ascontiguousarray(pairs)
    out = np.zeros(len(pairs), dtype=TD_DTYPE)
    for i, (t, l) in enumerate(pairs):
        out[i] = (t, l)
    return out


class Model:
    def __init__(self, path: str):
        self.h = lib().orc_model_load(path.encode())
        if not self.h:
            raise RuntimeError(f"oracle: cannot load {path}")
        self.n_vocab = lib().orc_n_vocab(self.h)
        self.n_embd = lib().orc_n_embd(self.h)
        self.n_layer = lib().orc_n_layer(self.h)
        self.n_ctx_train = lib().orc_n_ctx_train(self.h)
        self.bos = lib().orc_token_bos(self.h)

    def is_eog(self, tok: int) -> bool:
        return bool(lib().orc_is_eog(self.h, int(tok)))

    def weight_bytes_per_token(self) -> int:
        return int(lib().orc_weight_bytes_per_token(self.h))

    def close(self):
        if self.h:
            lib().orc_model_free(self.h)
            self.h = None


class Ctx:
    def __init__(self, model: Model, n_ctx: int = 4096, mode: int = MODE_GGML, n_threads: int = 0):
        self.m = model
        if n_threads <= 0:
            n_threads = os.cpu_count() or 1
        self.n_threads = n_threads
        self.h = lib().orc_ctx_create(model.h, n_ctx, mode, n_threads)

    def clear(self):
        lib().orc_kv_clear(self.h)

    def kv_shift(self, p0: int, p1: int):
        if lib().orc_kv_shift(self.h, p0, p1):
            raise RuntimeError("orc_kv_shift: bad range")

    def kv_seq_add(self, p0: int, p1: int, delta: int):
        if lib().orc_kv_seq_add(self.h, p0, p1, delta):
            raise RuntimeError("orc_kv_seq_add: bad range")

    def kv_seq_div(self, p0: int, p1: int, d: int):
        if lib().orc_kv_seq_div(self.h, p0, p1, d):
            raise RuntimeError("orc_kv_seq_div: bad arguments")

    @property
    def next_pos(self) -> int:
        return lib().orc_next_pos(self.h)

    @property
    def n_past(self) -> int:
        return lib().orc_n_past(self.h)

    def decode(self, tokens: Sequence[int], all_logits: bool = False) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        rc = lib().orc_decode(self.h, _p(t), len(t), 1 if all_logits else 0)
        if rc:
            raise RuntimeError(f"orc_decode failed rc={rc}")
        n = len(t) if all_logits else 1
        V = self.m.n_vocab
        ptr = lib().orc_get_logits(self.h, 0)
        return np.ctypeslib.as_array(ptr, shape=(n, V)).copy()

    def hidden(self, n: int) -> np.ndarray:
        ptr = lib().orc_get_hidden(self.h, 0)
        return np.ctypeslib.as_array(ptr, shape=(n, self.m.n_embd)).copy()

    def complete(self, prompt: Sequence[int], max_tokens: int, seed: int = 0, temp: float = 0.8, top_p: float = 0.95):
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        toks = np.zeros(max_tokens, dtype=np.int32)
        top = np.zeros((max_tokens, 10), dtype=TD_DTYPE)
        n = lib().orc_session_complete(self.h, _p(p), len(p), max_tokens, seed, temp, top_p, _p(toks), _p(top))
        if n < 0:
            raise RuntimeError("orc_session_complete failed")
        return toks[:n].copy(), top[:n].copy()

    def fill_ctx(self, prompt: Sequence[int], resp: Sequence[int], claimed: np.ndarray, n_claimed: Optional[np.ndarray] = None):
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        r = np.ascontiguousarray(resp, dtype=np.int32)
        cl = np.ascontiguousarray(claimed, dtype=np.int32).reshape(len(r), 10)
        nc = np.full(len(r), 10, dtype=np.int32) if n_claimed is None else np.ascontiguousarray(n_claimed, dtype=np.int32)
        out = np.zeros((len(r), 10), dtype=TD_DTYPE)
        out_n = np.zeros(len(r), dtype=np.int32)
        rc = lib().orc_session_fill_ctx(self.h, _p(p), len(p), _p(r), len(r), _p(cl), _p(nc), _p(out), _p(out_n))
        if rc:
            raise RuntimeError("orc_session_fill_ctx failed")
        return out, out_n

    def close(self):
        if self.h:
            lib().orc_ctx_free(self.h)
            self.h = None


def topk(logits: np.ndarray, k: int) -> np.ndarray:
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    out = np.zeros(k, dtype=TD_DTYPE)
    lib().orc_topk(_p(lg), len(lg), k, _p(out))
    return out


def gather_sorted(logits: np.ndarray, ids: Sequence[int]) -> np.ndarray:
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    i = np.ascontiguousarray(ids, dtype=np.int32)
    out = np.zeros(max(1, len(i)), dtype=TD_DTYPE)
    n = lib().orc_gather_sorted(_p(lg), len(lg), _p(i), len(i), _p(out))
    return out[:n]


def lc_compare(a, b) -> Tuple[float, float, float]:
    a, b = as_td(a), as_td(b)
    m = lib().orc_lc_compare(_p(a), len(a), _p(b), len(b))
    return (m.top1Match, m.distance, m.jsd)


def lc_similarity(a, b) -> float:
    a, b = as_td(a), as_td(b)
    return float(lib().orc_lc_similarity(_p(a), len(a), _p(b), len(b)))


def lc_score(metrics: Sequence[Tuple[float, float, float]]) -> float:
    arr = (Metrics * len(metrics))(*[Metrics(*m) for m in metrics])
    return float(lib().orc_lc_score(C.cast(arr, C.c_void_p), len(metrics)))


def dequantize(gtype: int, blocks: np.ndarray, n: int) -> np.ndarray:
    b = np.ascontiguousarray(blocks, dtype=np.uint8)
    out = np.zeros(n, dtype=np.float32)
    rc = lib().orc_dequantize(gtype, _p(b), n, _p(out))
    assert rc == 0
    return out


def matvec(gtype: int, w: np.ndarray, rows: int, k: int, x: np.ndarray, mode: int = MODE_GGML) -> np.ndarray:
    wb = np.ascontiguousarray(w, dtype=np.uint8)
    xx = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros(rows, dtype=np.float32)
    rc = lib().orc_matvec(gtype, _p(wb), rows, k, _p(xx), _p(y), mode)
    assert rc == 0
    return y


def quantize_q8_K(x: np.ndarray):
    xx = np.ascontiguousarray(x, dtype=np.float32)
    k = len(xx)
    qs = np.zeros(k, dtype=np.int8); d = np.zeros(k // 256, dtype=np.float32); bs = np.zeros(k // 16, dtype=np.int16)
    assert lib().orc_quantize_q8_K(_p(xx), k, _p(qs), _p(d), _p(bs)) == 0
    return qs, d, bs


def quantize_q8_0(x: np.ndarray):
    xx = np.ascontiguousarray(x, dtype=np.float32)
    k = len(xx)
    qs = np.zeros(k, dtype=np.int8); d = np.zeros(k // 32, dtype=np.float32)
    assert lib().orc_quantize_q8_0(_p(xx), k, _p(qs), _p(d)) == 0
    return qs, d


class Sampler:
    def __init__(self, seed: int = 0, temp: float = 0.8, top_p: float = 0.95, top_k: int = 40, min_p: float = 0.05, min_keep: int = 0):
        self.h = lib().orc_sampler_create_ex(seed, temp, top_p, top_k, min_p, min_keep)

    def sample(self, logits: np.ndarray) -> int:
        lg = np.ascontiguousarray(logits, dtype=np.float32)
        return int(lib().orc_sampler_sample(self.h, _p(lg), len(lg)))

    def sample_candidates(self, cand) -> int:
        c = as_td(cand)
        return int(lib().orc_sampler_sample_candidates(self.h, _p(c), len(c)))

    def reset(self):
        lib().orc_sampler_reset(self.h)

    def close(self):
        if self.h:
            lib().orc_sampler_free(self.h)
            self.h = None


# ---- the reference's own LogitComparer (oracle/_ref), present only where it was built from /root/reference ----
_ref = None


def ref_lib() -> Optional[C.CDLL]:
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libref_logitcomparer.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        _ref.ref_lc_compare.restype = Metrics
        _ref.ref_lc_compare.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        _ref.ref_lc_similarity.restype = C.c_float
        _ref.ref_lc_similarity.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        _ref.ref_lc_score.restype = C.c_float
        _ref.ref_lc_score.argtypes = [C.c_void_p, C.c_int32]
    return _ref


def ref_compare(a, b) -> Tuple[float, float, float]:
    a, b = as_td(a), as_td(b)
    m = ref_lib().ref_lc_compare(_p(a), len(a), _p(b), len(b))
    return (m.top1Match, m.distance, m.jsd)


def ref_similarity(a, b) -> float:
    a, b = as_td(a), as_td(b)
    return float(ref_lib().ref_lc_similarity(_p(a), len(a), _p(b), len(b)))


def ref_score(metrics) -> float:
    arr = (Metrics * len(metrics))(*[Metrics(*m) for m in metrics])
    return float(ref_lib().ref_lc_score(C.cast(arr, C.c_void_p), len(metrics)))
