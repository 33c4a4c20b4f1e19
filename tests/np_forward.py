"""Independent numpy float64 forward of the llama / qwen2 graph on weights dequantised by gguf-py.

Used to validate the *structure* of the oracle's forward (op order, rope pairing, GQA head mapping, bias,
tied lm_head) with code that shares nothing with oracle/oracle.cpp.  llama.cpp graph: upstream
src/llama-model.cpp llm_build_llama / llm_build_qwen2 (un-vendored)."""
from __future__ import annotations

import numpy as np


def load_weights(path: str):
    import gguf

    r = gguf.GGUFReader(path)
    W = {}
    for t in r.tensors:
        W[t.name] = np.asarray(gguf.quants.dequantize(t.data, t.tensor_type), dtype=np.float64)
    f = {}
    for k, v in r.fields.items():
        try:
            f[k] = v.contents()
        except Exception:
            pass
    return W, f


def forward(W, f, tokens, kv_f16: bool = True):
    arch = f["general.architecture"]
    g = lambda k: f[f"{arch}.{k}"]
    d, L, nh, nkv = g("embedding_length"), g("block_count"), g("attention.head_count"), g("attention.head_count_kv")
    eps, theta = g("attention.layer_norm_rms_epsilon"), g("rope.freq_base")
    dh = d // nh
    n = len(tokens)
    x = W["token_embd.weight"][np.asarray(tokens)]
    ffac = W.get("rope_freqs.weight")
    inv = theta ** (-np.arange(0, dh, 2) / dh)
    if ffac is not None:
        inv = inv / ffac
    ang = np.arange(n)[:, None] * inv[None, :]
    cos, sin = np.cos(ang), np.sin(ang)
    neox = arch == "qwen2"

    def rms(v, w):
        return v / np.sqrt((v * v).mean(-1, keepdims=True) + eps) * w

    def rope(v, heads):
        v = v.reshape(n, heads, dh).copy()
        if neox:
            a, b = v[..., : dh // 2], v[..., dh // 2:]
            out = np.concatenate([a * cos[:, None] - b * sin[:, None], a * sin[:, None] + b * cos[:, None]], -1)
        else:
            a, b = v[..., 0::2], v[..., 1::2]
            out = np.empty_like(v)
            out[..., 0::2] = a * cos[:, None] - b * sin[:, None]
            out[..., 1::2] = a * sin[:, None] + b * cos[:, None]
        return out

    for l in range(L):
        p = f"blk.{l}."
        h = rms(x, W[p + "attn_norm.weight"])
        q = h @ W[p + "attn_q.weight"].T
        k = h @ W[p + "attn_k.weight"].T
        v = h @ W[p + "attn_v.weight"].T
        if p + "attn_q.bias" in W:
            q, k, v = q + W[p + "attn_q.bias"], k + W[p + "attn_k.bias"], v + W[p + "attn_v.bias"]
        q, k = rope(q, nh), rope(k, nkv)
        v = v.reshape(n, nkv, dh)
        if kv_f16:
            k = k.astype(np.float16).astype(np.float64)
            v = v.astype(np.float16).astype(np.float64)
        out = np.zeros((n, nh, dh))
        mask = np.triu(np.full((n, n), -np.inf), 1)
        for hh in range(nh):
            kk, vv = k[:, hh // (nh // nkv)], v[:, hh // (nh // nkv)]
            s = q[:, hh] @ kk.T / np.sqrt(dh) + mask
            s = np.exp(s - s.max(-1, keepdims=True))
            s /= s.sum(-1, keepdims=True)
            out[:, hh] = s @ vv
        x = x + out.reshape(n, nh * dh) @ W[p + "attn_output.weight"].T
        h = rms(x, W[p + "ffn_norm.weight"])
        gt = h @ W[p + "ffn_gate.weight"].T
        up = h @ W[p + "ffn_up.weight"].T
        x = x + ((gt / (1 + np.exp(-gt))) * up) @ W[p + "ffn_down.weight"].T
    hN = rms(x, W["output_norm.weight"])
    wo = W["output.weight"] if "output.weight" in W else W["token_embd.weight"]
    return hN @ wo.T
