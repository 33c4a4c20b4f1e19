"""Synthetic random-init GGUF v3 writer for the blama hot path.

blama loads its weights with ``llama_model_load_from_file`` (reference
inference/code/llama/Model.cpp:50-53); there is no network here, so every
model the tests / bench use is written by this module: llama / qwen2
architecture metadata plus random quant blocks (Q4_K / Q5_K / Q6_K / Q8_0 /
F32) whose dequantised values are ~N(0, sigma) so the forward pass stays
numerically healthy.  No K-quant *quantiser* is needed for random-init
weights: the blocks themselves are drawn at random (scales, mins and nibbles),
which also exercises every bit-field of the formats.

This is a host-side tool (numpy only); nothing here is on the inference path.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

GGUF_MAGIC = 0x46554747
GGUF_VERSION = 3
ALIGNMENT = 32

# ggml_type ids (ggml.h)
F32, F16, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 8, 12, 13, 14
BLOCK = {F32: (1, 4), F16: (1, 2), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176), Q6_K: (256, 210)}
TYPE_NAME = {F32: "F32", F16: "F16", Q8_0: "Q8_0", Q4_K: "Q4_K", Q5_K: "Q5_K", Q6_K: "Q6_K"}

# gguf metadata value types
T_U8, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64 = range(13)


@dataclasses.dataclass
class ModelShape:
    arch: str            # "llama" | "qwen2"
    name: str
    d_model: int
    n_layer: int
    n_head: int
    n_head_kv: int
    d_head: int
    d_ffn: int
    vocab: int
    ctx_train: int
    rope_theta: float
    rms_eps: float
    ftype: str           # "Q4_K_M" | "Q8_0" | "F32"
    tied: bool = False   # lm_head shares token_embd
    rope_freqs: bool = False
    qkv_bias: bool = False
    is_70b: bool = False


SHAPES: Dict[str, ModelShape] = {
    # tiny shapes for unit tests (d_model % 256 == 0 so K-quants apply)
    "tiny-llama-q4km": ModelShape("llama", "tiny-llama", 256, 2, 4, 2, 64, 512, 1024, 2048, 5e5, 1e-5, "Q4_K_M", rope_freqs=True),
    "tiny-llama-q8": ModelShape("llama", "tiny-llama-q8", 256, 2, 4, 2, 64, 512, 1024, 2048, 5e5, 1e-5, "Q8_0", tied=True, rope_freqs=True),
    "tiny-qwen2-q8": ModelShape("qwen2", "tiny-qwen2", 256, 2, 4, 2, 64, 768, 1536, 2048, 1e6, 1e-6, "Q8_0", qkv_bias=True),
    "tiny-llama-f32": ModelShape("llama", "tiny-llama-f32", 256, 2, 4, 2, 64, 512, 1024, 2048, 5e5, 1e-5, "F32"),
    # d_head = 128 variants (the production head size) at small cost
    "small-llama-q4km": ModelShape("llama", "small-llama", 512, 4, 4, 2, 128, 1536, 4096, 4096, 5e5, 1e-5, "Q4_K_M", rope_freqs=True),
    "small-qwen2-q8": ModelShape("qwen2", "small-qwen2", 512, 3, 4, 2, 128, 1280, 5000, 4096, 1e6, 1e-6, "Q8_0", qkv_bias=True),
    "small-llama70-q4km": ModelShape("llama", "small-llama70", 512, 8, 8, 1, 64, 1024, 2048, 4096, 5e5, 1e-5, "Q4_K_M", rope_freqs=True, is_70b=True),
    # GQA ratios that are not powers of two (Qwen2.5-7B has 7): 3 query heads per KV head, and 7 with NEOX rotary + biases
    "small-llama-gq3": ModelShape("llama", "small-llama-gq3", 768, 2, 6, 2, 128, 1536, 2048, 4096, 5e5, 1e-5, "Q4_K_M", rope_freqs=True),
    "small-qwen2-gq7": ModelShape("qwen2", "small-qwen2-gq7", 1792, 2, 14, 2, 128, 2304, 3000, 4096, 1e6, 1e-6, "Q8_0", qkv_bias=True),
    # FFN rows longer than 16 blocks of 256: the persistent decode kernel quantises the SwiGLU output once and shares it (K-quant and Q8_0 forms)
    "small-llama-wideffn-q4km": ModelShape("llama", "small-llama-wideffn", 512, 2, 4, 2, 128, 4608, 2048, 4096, 5e5, 1e-5, "Q4_K_M", rope_freqs=True),
    "small-qwen2-wideffn-q8": ModelShape("qwen2", "small-qwen2-wideffn", 512, 2, 4, 2, 128, 4608, 2048, 4096, 1e6, 1e-6, "Q8_0", qkv_bias=True),
    # BASELINE.json configs
    "llama-3.2-1b-q8": ModelShape("llama", "Llama-3.2-1B-arch", 2048, 16, 32, 8, 64, 8192, 128256, 131072, 5e5, 1e-5, "Q8_0", tied=True, rope_freqs=True),
    "llama-3.1-8b-q4km": ModelShape("llama", "Llama-3.1-8B-arch", 4096, 32, 32, 8, 128, 14336, 128256, 131072, 5e5, 1e-5, "Q4_K_M", rope_freqs=True),
    "qwen2.5-7b-q8": ModelShape("qwen2", "Qwen2.5-7B-arch", 3584, 28, 28, 4, 128, 18944, 152064, 32768, 1e6, 1e-6, "Q8_0", qkv_bias=True),
    "llama-3.1-70b-q4km": ModelShape("llama", "Llama-3.1-70B-arch", 8192, 80, 64, 8, 128, 28672, 128256, 131072, 5e5, 1e-5, "Q4_K_M", rope_freqs=True, is_70b=True),
}


def use_more_bits(i: int, n: int) -> bool:
    """llama.cpp llama_tensor_get_type helper [upstream-recall, llama-quant.cpp]."""
    return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2


def tensor_type(shape: ModelShape, name: str, layer: int) -> int:
    """Per-tensor ggml type of the file type's mix ("Q4_K_M mix")."""
    if name.endswith("_norm.weight") or name.endswith(".bias") or name == "rope_freqs.weight":
        return F32
    if shape.ftype == "F32":
        return F32
    if shape.ftype == "Q8_0":
        return Q8_0
    assert shape.ftype == "Q4_K_M"
    if name == "output.weight":
        return Q6_K
    if name == "token_embd.weight":
        return Q6_K if shape.tied else Q4_K
    if "attn_v" in name:
        if use_more_bits(layer, shape.n_layer):
            return Q6_K
        return Q5_K if shape.is_70b else Q4_K
    if "ffn_down" in name:
        return Q6_K if use_more_bits(layer, shape.n_layer) else Q4_K
    return Q4_K


# ----------------------------------------------------------------------------------------------
# random quant blocks
# ----------------------------------------------------------------------------------------------

def _pack_k4_scales(sc: np.ndarray, mn: np.ndarray) -> np.ndarray:
    """Pack 8 six-bit scales + 8 six-bit mins into the 12-byte K-quant field
    (inverse of ggml-quants.c get_scale_min_k4)."""
    nb = sc.shape[0]
    out = np.zeros((nb, 12), dtype=np.uint8)
    out[:, 0:4] = (sc[:, 0:4] & 63) | ((sc[:, 4:8] >> 4) << 6)
    out[:, 4:8] = (mn[:, 0:4] & 63) | ((mn[:, 4:8] >> 4) << 6)
    out[:, 8:12] = (sc[:, 4:8] & 0xF) | ((mn[:, 4:8] & 0xF) << 4)
    return out


def random_blocks(rng: np.random.Generator, gtype: int, n_elems: int, sigma: float) -> np.ndarray:
    """Random quant blocks of `gtype` covering n_elems values with dequantised std ~ sigma."""
    bs, nbytes = BLOCK[gtype]
    assert n_elems % bs == 0
    nb = n_elems // bs
    if gtype == F32:
        return (rng.standard_normal(n_elems, dtype=np.float32) * np.float32(sigma)).astype(np.float32).view(np.uint8)
    if gtype == F16:
        return (rng.standard_normal(n_elems, dtype=np.float32) * np.float32(sigma)).astype(np.float16).view(np.uint8)
    raw = rng.integers(0, 256, size=(nb, nbytes), dtype=np.uint8)
    jitter = np.exp(0.25 * rng.standard_normal(nb, dtype=np.float32))
    if gtype == Q8_0:
        d = (sigma / 73.9 * jitter).astype(np.float16)
        raw[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
        return raw.reshape(-1)
    if gtype == Q6_K:
        # w = d * sc * (q - 32); sc int8 uniform (rms 73.9), q uniform 0..63 (std 18.47)
        d = (sigma / (73.9 * 18.47) * jitter).astype(np.float16)
        raw[:, 208:210] = d.view(np.uint8).reshape(nb, 2)
        return raw.reshape(-1)
    # Q4_K / Q5_K: w = d*sc*q - dmin*m ; choose m ~ sc and dmin = centre*d so w is ~symmetric
    sc = rng.integers(8, 64, size=(nb, 8), dtype=np.uint8)
    mn = np.clip(sc.astype(np.int16) + rng.integers(-3, 4, size=(nb, 8)), 0, 63).astype(np.uint8)
    centre, qstd = (7.5, 4.61) if gtype == Q4_K else (15.5, 9.23)
    dval = sigma / (39.0 * qstd) * jitter
    d = dval.astype(np.float16)
    dmin = (dval * centre).astype(np.float16)
    raw[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
    raw[:, 2:4] = dmin.view(np.uint8).reshape(nb, 2)
    raw[:, 4:16] = _pack_k4_scales(sc, mn)
    return raw.reshape(-1)


def llama3_rope_freqs(d_head: int, theta: float, factor: float = 8.0, low: float = 1.0, high: float = 4.0,
                      orig_ctx: int = 8192) -> np.ndarray:
    """rope_freqs.weight as written by llama.cpp's HF converter for Llama-3.1 rope scaling
    [upstream-recall convert_hf_to_gguf.py]; consumed as `freq_factors` by ggml rope."""
    freqs = 1.0 / (theta ** (np.arange(0, d_head, 2, dtype=np.float64) / d_head))
    low_wl, high_wl = orig_ctx / low, orig_ctx / high
    out = []
    for f in freqs:
        wl = 2 * np.pi / f
        if wl < high_wl:
            out.append(1.0)
        elif wl > low_wl:
            out.append(factor)
        else:
            smooth = (orig_ctx / wl - low) / (high - low)
            out.append(1.0 / ((1 - smooth) / factor + smooth))
    return np.asarray(out, dtype=np.float32)


# ----------------------------------------------------------------------------------------------
# GGUF serialisation
# ----------------------------------------------------------------------------------------------

def _s(b: str) -> bytes:
    e = b.encode("utf-8")
    return struct.pack("<Q", len(e)) + e


def _kv(key: str, vtype: int, value) -> bytes:
    out = _s(key) + struct.pack("<I", vtype)
    fmt = {T_U8: "<B", T_I8: "<b", T_U16: "<H", T_I16: "<h", T_U32: "<I", T_I32: "<i", T_F32: "<f", T_BOOL: "<?",
           T_U64: "<Q", T_I64: "<q", T_F64: "<d"}
    if vtype == T_STR:
        return out + _s(value)
    if vtype == T_ARR:
        etype, items = value
        out += struct.pack("<IQ", etype, len(items))
        if etype == T_STR:
            out += b"".join(_s(x) for x in items)
        elif etype == T_I32:
            out += np.asarray(items, dtype="<i4").tobytes()
        elif etype == T_F32:
            out += np.asarray(items, dtype="<f4").tobytes()
        else:
            raise ValueError(etype)
        return out
    return out + struct.pack(fmt[vtype], value)


def bytes_to_unicode() -> List[str]:
    """GPT-2's byte -> printable character table (byte-level BPE alphabet): index = byte value."""
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    out, extra = [""] * 256, 0
    for b in range(256):
        if b in keep:
            out[b] = chr(b)
        else:
            out[b] = chr(256 + extra)
            extra += 1
    return out


_SEED_MERGES: Optional[List[Tuple[str, str]]] = None


def _seed_merges() -> List[Tuple[str, str]]:
    global _SEED_MERGES
    if _SEED_MERGES is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpe_seed_merges.txt")
        with open(path, encoding="utf-8") as f:
            _SEED_MERGES = [tuple(line.rstrip("\n").split(" ")) for line in f if line.strip()]
    return _SEED_MERGES


def synth_vocab(shape: "ModelShape | str", seed: int = 0xB1A4A) -> Tuple[List[str], List[int], List[str]]:
    """(tokens, token types, merges) of a synthetic byte-level BPE vocabulary with shape.vocab entries, laid out like the real
    Llama-3 / Qwen2 files: ids 0..255 the byte alphabet, then one token per merge in rank order (merge r creates id 256 + r), then
    the CONTROL tokens from special_tokens()['bos' ...] upwards.  The first merges are trained on English text
    (tools/gen_bpe_seed_merges.py), the rest are random concatenations of earlier tokens: every merge list is a valid BPE."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    sp = special_tokens(shape)
    base = min(sp.values())
    assert base > 256, "vocabulary too small for a byte-level alphabet"
    b2u = bytes_to_unicode()
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    order = keep + [b for b in range(256) if b not in keep]
    tokens = [b2u[b] for b in order]
    have = set(tokens)
    merges: List[str] = []
    for a, b in _seed_merges():
        if len(tokens) >= base:
            break
        if a in have and b in have and a + b not in have:
            tokens.append(a + b); have.add(a + b); merges.append(f"{a} {b}")
    rng = np.random.default_rng([seed, 0x70CAB])
    short = [t for t in tokens if len(t) <= 6]          # both halves of a random merge: the result stays <= 12 characters
    while len(tokens) < base:
        need = base - len(tokens)
        for ua, ub in rng.random((need + need // 8 + 64, 2)).tolist():
            n = len(short)
            a, b = short[int(ua * n)], short[int(ub * n)]
            new = a + b
            if new in have:
                continue
            tokens.append(new); have.add(new); merges.append(a + " " + b)
            if len(new) <= 6:
                short.append(new)
            if len(tokens) >= base:
                break
    types = [1] * base
    names = {"llama": {"bos": "<|begin_of_text|>", "eos": "<|end_of_text|>", "eot": "<|eot_id|>", "pad": "<|finetune_right_pad_id|>"},
             "qwen2": {"bos": "<|endoftext|>", "eos": "<|im_end|>", "eot": "<|im_end|>", "pad": "<|endoftext|>"}}[shape.arch]
    by_id = {}
    for k in ("pad", "bos", "eot", "eos"):
        by_id[sp[k]] = names[k]
    for i in range(base, shape.vocab):
        tokens.append(by_id.get(i, f"<|reserved_special_token_{i - base}|>"))
        types.append(3)      # CONTROL
    return tokens, types, merges


def special_tokens(shape: ModelShape) -> Dict[str, int]:
    """bos / eos / eot ids placed where the real vocabularies keep them."""
    v = shape.vocab
    if shape.arch == "qwen2":
        base = 151643 if v > 151700 else v - 8
        return {"bos": base, "eos": base + 2, "eot": base + 2, "pad": base}
    base = 128000 if v > 128100 else v - 16
    return {"bos": base, "eos": base + 1, "eot": base + 9, "pad": base + 4}


def plan_tensors(shape: ModelShape) -> List[Tuple[str, Tuple[int, ...], int, float, int]]:
    """[(name, ne (innermost first), ggml type, sigma, layer)] in file order."""
    d, ff, V = shape.d_model, shape.d_ffn, shape.vocab
    dq, dkv = shape.n_head * shape.d_head, shape.n_head_kv * shape.d_head
    sig = 0.02
    # logit std ~ 2: logits = W_out . (normed x), |normed x| ~ sqrt(d)
    sig_out = 2.0 / np.sqrt(d)
    out: List[Tuple[str, Tuple[int, ...], int, float, int]] = []

    def add(name, ne, sigma, layer=-1):
        out.append((name, ne, tensor_type(shape, name, layer), sigma, layer))

    add("token_embd.weight", (d, V), sig_out if shape.tied else 1.0)
    for i in range(shape.n_layer):
        p = f"blk.{i}."
        add(p + "attn_norm.weight", (d,), 0.0, i)
        add(p + "attn_q.weight", (d, dq), 1.0 / np.sqrt(d), i)
        add(p + "attn_k.weight", (d, dkv), 1.0 / np.sqrt(d), i)
        add(p + "attn_v.weight", (d, dkv), 1.0 / np.sqrt(d), i)
        if shape.qkv_bias:
            add(p + "attn_q.bias", (dq,), 0.1, i)
            add(p + "attn_k.bias", (dkv,), 0.1, i)
            add(p + "attn_v.bias", (dkv,), 0.1, i)
        add(p + "attn_output.weight", (dq, d), 0.5 / np.sqrt(dq), i)
        add(p + "ffn_norm.weight", (d,), 0.0, i)
        add(p + "ffn_gate.weight", (d, ff), 1.0 / np.sqrt(d), i)
        add(p + "ffn_up.weight", (d, ff), 1.0 / np.sqrt(d), i)
        add(p + "ffn_down.weight", (ff, d), 0.5 / np.sqrt(ff), i)
    add("output_norm.weight", (d,), 0.0)
    if not shape.tied:
        add("output.weight", (d, V), sig_out)
    if shape.rope_freqs:
        add("rope_freqs.weight", (shape.d_head // 2,), 0.0)
    _ = sig
    return out


def model_bytes(shape: ModelShape) -> int:
    tot = 0
    for _, ne, t, _, _ in plan_tensors(shape):
        bs, nb = BLOCK[t]
        tot += int(np.prod(ne)) // bs * nb
    return tot


def write_gguf(path: str, shape: ModelShape | str, seed: int = 0xB1A4A, chunk_elems: int = 1 << 26, pre: Optional[str] = None) -> Dict[str, int]:
    """Write a random-init GGUF for `shape`; returns the special-token ids.  pre: tokenizer.ggml.pre override (tokenizer tests)."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    a = shape.arch
    sp = special_tokens(shape)
    tokens, ttype, merges = synth_vocab(shape, seed)
    ftype_id = {"F32": 0, "Q8_0": 7, "Q4_K_M": 15}[shape.ftype]
    kvs = [
        _kv("general.architecture", T_STR, a),
        _kv("general.name", T_STR, shape.name + " (random-init, synthetic)"),
        _kv("general.file_type", T_U32, ftype_id),
        _kv("general.alignment", T_U32, ALIGNMENT),
        _kv(f"{a}.context_length", T_U32, shape.ctx_train),
        _kv(f"{a}.embedding_length", T_U32, shape.d_model),
        _kv(f"{a}.block_count", T_U32, shape.n_layer),
        _kv(f"{a}.feed_forward_length", T_U32, shape.d_ffn),
        _kv(f"{a}.attention.head_count", T_U32, shape.n_head),
        _kv(f"{a}.attention.head_count_kv", T_U32, shape.n_head_kv),
        _kv(f"{a}.attention.layer_norm_rms_epsilon", T_F32, shape.rms_eps),
        _kv(f"{a}.rope.freq_base", T_F32, shape.rope_theta),
        _kv(f"{a}.rope.dimension_count", T_U32, shape.d_head),
        _kv(f"{a}.vocab_size", T_U32, shape.vocab),
        _kv("tokenizer.ggml.model", T_STR, "gpt2"),
        _kv("tokenizer.ggml.pre", T_STR, pre or ("llama-bpe" if a == "llama" else "qwen2")),
        _kv("tokenizer.ggml.tokens", T_ARR, (T_STR, tokens)),
        _kv("tokenizer.ggml.token_type", T_ARR, (T_I32, ttype)),
        _kv("tokenizer.ggml.merges", T_ARR, (T_STR, merges)),
        _kv("tokenizer.ggml.bos_token_id", T_U32, sp["bos"]),
        _kv("tokenizer.ggml.eos_token_id", T_U32, sp["eos"]),
        _kv("tokenizer.ggml.eot_token_id", T_U32, sp["eot"]),
        _kv("tokenizer.ggml.padding_token_id", T_U32, sp["pad"]),
        _kv("tokenizer.ggml.add_bos_token", T_BOOL, a == "llama"),
    ]
    plan = plan_tensors(shape)
    infos = b""
    off = 0
    sizes = []
    for name, ne, t, _, _ in plan:
        bs, nbytes = BLOCK[t]
        n = int(np.prod(ne))
        assert ne[0] % bs == 0, (name, ne, t)
        sz = n // bs * nbytes
        infos += _s(name) + struct.pack("<I", len(ne)) + struct.pack(f"<{len(ne)}Q", *ne) + struct.pack("<IQ", t, off)
        sizes.append(sz)
        off += (sz + ALIGNMENT - 1) // ALIGNMENT * ALIGNMENT
    head = struct.pack("<IIQQ", GGUF_MAGIC, GGUF_VERSION, len(plan), len(kvs)) + b"".join(kvs) + infos
    pad = (-len(head)) % ALIGNMENT
    with open(path, "wb") as f:
        f.write(head + b"\0" * pad)
        for idx, ((name, ne, t, sigma, _), sz) in enumerate(zip(plan, sizes)):
            rng = np.random.default_rng([seed, idx])
            n = int(np.prod(ne))
            if name.endswith("_norm.weight"):
                data = (1.0 + 0.02 * rng.standard_normal(n, dtype=np.float32)).astype(np.float32).tobytes()
                f.write(data)
            elif name == "rope_freqs.weight":
                f.write(llama3_rope_freqs(shape.d_head, shape.rope_theta).tobytes())
            else:
                bs, _ = BLOCK[t]
                step = max(bs, chunk_elems // ne[0] * ne[0])
                for s0 in range(0, n, step):
                    f.write(random_blocks(rng, t, min(step, n - s0), sigma).tobytes())
            f.write(b"\0" * ((-sz) % ALIGNMENT))
    return sp


def synth_prompt(shape: ModelShape | str, n: int, seed: int) -> np.ndarray:
    """Seeded uniform token ids excluding the special ids."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    sp = set(special_tokens(shape).values())
    rng = np.random.default_rng([0x5EED, seed])
    out = rng.integers(0, shape.vocab, size=n, dtype=np.int32)
    lo = min(sp)
    out = np.where(np.isin(out, list(sp)), (out + 97) % lo, out).astype(np.int32)
    return out


if __name__ == "__main__":
    import argparse
    import time

    ap = argparse.ArgumentParser()
    ap.add_argument("shape", choices=sorted(SHAPES))
    ap.add_argument("path")
    ap.add_argument("--seed", type=int, default=0xB1A4A)
    args = ap.parse_args()
    t0 = time.time()
    write_gguf(args.path, args.shape, args.seed)
    print(f"wrote {args.path}: {model_bytes(SHAPES[args.shape]) / 1e9:.3f} GB of tensors in {time.time() - t0:.1f}s")
