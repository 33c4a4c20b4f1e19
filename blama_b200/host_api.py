"""ctypes view of the reference-shaped host API (blama_b200/host/host_capi.cpp): bl::llama::Model / Instance / Session /
LogitComparer / MetricsAggregator / Sampler.  Errors surface as HostError carrying the C++ exception text, which are the
strings the reference's own tests pin (inference/test/t-integration.cpp:137-217)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi

TD_DTYPE = capi.TD_DTYPE
_vp, _i32, _u32, _f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_float

HOST_SYMBOLS = {
    "blh_last_error": (C.c_char_p, []),
    "blh_init": (None, []),
    "blh_model_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "blh_model_free": (None, [_vp]),
    "blh_model_train_ctx": (C.c_int, [_vp]),
    "blh_model_tokenize": (C.c_int, [_vp, C.c_char_p, C.c_int, _vp, C.c_int]),
    "blh_model_token_to_string": (C.c_int, [_vp, _i32, C.c_char_p, C.c_int]),
    "blh_instance_create": (C.c_int, [_vp, _u32, _u32, C.POINTER(_vp)]),
    "blh_instance_free": (None, [_vp]),
    "blh_instance_warmup": (C.c_int, [_vp]),
    "blh_instance_ctx": (_vp, [_vp]),
    "blh_session_start": (C.c_int, [_vp, _u32, _f32, _f32, C.c_int]),
    "blh_session_stop": (None, [_vp]),
    "blh_session_set_initial_prompt": (C.c_int, [_vp, _vp, C.c_int]),
    "blh_session_complete": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "blh_session_stream": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "blh_session_fill_ctx": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "blh_session_verify": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.POINTER(_f32)]),
    "blh_session_get_state": (C.c_int, [_vp]),
    "blh_session_set_state": (C.c_int, [_vp]),
    "blh_lc_compare": (None, [_vp, _i32, _vp, _i32, _vp]),
    "blh_lc_similarity": (_f32, [_vp, _i32, _vp, _i32]),
    "blh_lc_score": (_f32, [_vp, _i32]),
    "blh_sampler_draw": (C.c_int, [_u32, _f32, _f32, _i32, _f32, _i32, _vp, _i32, _i32, _i32, _vp]),
}

_bound = False


def lib() -> C.CDLL:
    global _bound
    l = capi.lib()
    if not _bound:
        for name, (res, args) in HOST_SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        l.blh_init()
        _bound = True
    return l


class HostError(RuntimeError):
    pass


def _check(rc: int):
    if rc:
        raise HostError((lib().blh_last_error() or b"").decode(errors="replace"))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def as_td(pairs) -> np.ndarray:
    if isinstance(pairs, np.ndarray) and pairs.dtype == TD_DTYPE:
        return np.ascontiguousarray(pairs)
    out = np.zeros(len(pairs), dtype=TD_DTYPE)
    for i, (t, l) in enumerate(pairs):
        out[i] = (t, l)
    return out


class Model:
    def __init__(self, path: str, device: int = 0, gpu: bool = True, prefix_bos: bool = False):
        h = _vp()
        _check(lib().blh_model_create(path.encode(), device, int(gpu), int(prefix_bos), C.byref(h)))
        self.h = h

    def train_ctx(self) -> int:
        return lib().blh_model_train_ctx(self.h)

    def tokenize(self, text: str, add_special: bool = True) -> np.ndarray:
        out = np.zeros(len(text) + 4, dtype=np.int32)
        n = lib().blh_model_tokenize(self.h, text.encode(), int(add_special), _p(out), len(out))
        return out[:n].copy()

    def token_to_string(self, tok: int) -> str:
        buf = C.create_string_buffer(256)
        n = lib().blh_model_token_to_string(self.h, int(tok), buf, 256)
        return buf.raw[:min(n, 256)].decode(errors="replace")

    def close(self):
        if self.h:
            lib().blh_model_free(self.h)
            self.h = None


class Instance:
    """Instance + its (single) Session, addressed together like the reference's Instance::startSession."""

    def __init__(self, model: Model, ctx_size: int = 0, batch_size: int = 0):
        h = _vp()
        _check(lib().blh_instance_create(model.h, ctx_size, batch_size, C.byref(h)))
        self.h = h
        self.model = model

    def warmup(self):
        _check(lib().blh_instance_warmup(self.h))

    def raw_ctx(self) -> "capi.Ctx":
        """non-owning capi.Ctx view of this instance's C-ABI context (timers, launch counters)"""
        c = capi.Ctx.__new__(capi.Ctx)
        c.m = None
        c.h = lib().blh_instance_ctx(self.h)
        c.close = lambda: None
        return c

    def start_session(self, seed: int = 0, temperature: float = 0.8, top_p: float = 0.95, sequential_verify: bool = False):
        _check(lib().blh_session_start(self.h, seed, temperature, top_p, int(sequential_verify)))
        return self

    def stop_session(self):
        lib().blh_session_stop(self.h)

    def set_initial_prompt(self, tokens: Sequence[int]):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        _check(lib().blh_session_set_initial_prompt(self.h, _p(t), len(t)))

    def complete(self, max_tokens: int, prompt: Sequence[int] = ()):
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        toks = np.zeros(max(1, max_tokens), dtype=np.int32)
        top = np.zeros((max(1, max_tokens), 10), dtype=TD_DTYPE)
        nl = np.zeros(max(1, max_tokens), dtype=np.int32)
        n = _i32(0)
        _check(lib().blh_session_complete(self.h, _p(p), len(p), max_tokens, _p(toks), _p(top), _p(nl), C.byref(n)))
        return toks[: n.value].copy(), top[: n.value].copy()

    def stream(self, max_tokens: int) -> np.ndarray:
        toks = np.zeros(max(1, max_tokens), dtype=np.int32)
        n = _i32(0)
        _check(lib().blh_session_stream(self.h, max_tokens, _p(toks), C.byref(n)))
        return toks[: n.value].copy()

    def fill_ctx(self, tokens: Sequence[int], claimed: np.ndarray, n_claimed: Optional[np.ndarray] = None):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        n = len(t)
        cl = np.ascontiguousarray(claimed).reshape(n, 10)
        assert cl.dtype == TD_DTYPE
        nc = np.full(n, 10, dtype=np.int32) if n_claimed is None else np.ascontiguousarray(n_claimed, dtype=np.int32)
        out = np.zeros((n, 10), dtype=TD_DTYPE)
        out_n = np.zeros(n, dtype=np.int32)
        _check(lib().blh_session_fill_ctx(self.h, _p(t), n, _p(cl), _p(nc), _p(out), _p(out_n)))
        return out, out_n

    def verify(self, tokens: Sequence[int], claimed: np.ndarray, n_claimed: Optional[np.ndarray] = None) -> float:
        """fillCtx + LogitComparer + MetricsAggregator in C++ (what Server::verify runs); returns the score"""
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        n = len(t)
        cl = np.ascontiguousarray(claimed).reshape(n, 10)
        assert cl.dtype == TD_DTYPE
        nc = np.full(n, 10, dtype=np.int32) if n_claimed is None else np.ascontiguousarray(n_claimed, dtype=np.int32)
        score = _f32(0)
        _check(lib().blh_session_verify(self.h, _p(t), n, _p(cl), _p(nc), C.byref(score)))
        return float(score.value)

    def get_state(self):
        _check(lib().blh_session_get_state(self.h))

    def set_state(self):
        _check(lib().blh_session_set_state(self.h))

    def close(self):
        if self.h:
            lib().blh_instance_free(self.h)
            self.h = None


def lc_compare(a, b) -> Tuple[float, float, float]:
    a, b = as_td(a), as_td(b)
    out = np.zeros(3, dtype=np.float32)
    lib().blh_lc_compare(_p(a), len(a), _p(b), len(b), _p(out))
    return (float(out[0]), float(out[1]), float(out[2]))


def lc_similarity(a, b) -> float:
    a, b = as_td(a), as_td(b)
    return float(lib().blh_lc_similarity(_p(a), len(a), _p(b), len(b)))


def lc_score(metrics) -> float:
    m = np.ascontiguousarray(np.asarray(metrics, dtype=np.float32).reshape(-1, 3))
    return float(lib().blh_lc_score(_p(m), len(m)))


def sampler_draw(cand, n_draws: int, seed: int = 0, temp: float = 0.8, top_p: float = 0.95, top_k: int = 40, min_p: float = 0.05,
                 min_keep: int = 0, is_sorted: bool = True) -> np.ndarray:
    c = as_td(cand)
    out = np.zeros(n_draws, dtype=np.int32)
    _check(lib().blh_sampler_draw(seed, temp, top_p, top_k, min_p, min_keep, _p(c), len(c), int(is_sorted), n_draws, _p(out)))
    return out
