"""Golden vectors that pin the oracle's transformer forward to an INDEPENDENT implementation: Hugging Face transformers reads the
synthetic GGUF through its own GGUF loader (transformers.modeling_gguf_pytorch_utils.load_gguf_checkpoint: gguf-py de-quantisation,
tensor-name map, reverse permutation of the llama Q / K rows) into its own LlamaForCausalLM / Qwen2ForCausalLM and runs a float32
forward pass.  llama.cpp itself (a pinned submodule of the reference) is not vendored and cannot be built here; Hugging Face's
models are the implementations llama.cpp's graphs are written to reproduce.  What this pins: the layer graph (RMSNorm placement,
GQA, SwiGLU, residuals), the rotary convention (llama: interleaved pairs on permuted rows == rotate-half on the originals; qwen2:
NEOX), the llama-3 frequency factors (rope_freqs.weight divides the base frequencies -- the loader ignores that tensor, so the
script applies it to the rotary module as llama.cpp does), Q/K/V biases, the tied head.  What it does not pin: ggml's quantised
integer dot products (pinned separately: de-quantisation against gguf-py, Q8_K / Q8_0 quantisers and vec_dot restatements by
their own tests).

    python tools/gen_hf_forward_golden.py        -> tests/golden/hf_forward_golden.npz
The oracle's F32 mode differs from the float32 HF pass by the f16 KV cache and the f16 soft-max probabilities llama.cpp uses:
2-6e-3 on logits of standard deviation 2.1 (printed)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blama_b200 import gguf_synth as gs  # noqa: E402

SHAPES = ["small-llama-q4km", "small-qwen2-q8", "tiny-llama-q8", "small-llama70-q4km", "small-llama-gq3"]
BIG_SHAPES = ["llama-3.2-1b-q8"]       # BASELINE configs[0] at full size: only the top-32 of the last positions is kept (128 256 logits per row)
N_TOK, KEEP, TOPK = 24, 4, 32


def hf_logits(path: str, tokens):
    import gguf
    import torch
    from transformers import AutoConfig, AutoModelForCausalLM
    from transformers.modeling_gguf_pytorch_utils import load_gguf_checkpoint

    d, f = os.path.split(path)
    cfg = AutoConfig.from_pretrained(d, gguf_file=f)
    m = AutoModelForCausalLM.from_config(cfg).to(torch.float32).eval()
    sd = load_gguf_checkpoint(path, return_tensors=True, model_to_load=m)["tensors"]
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and set(res.missing_keys) <= {"lm_head.weight"}, res      # tied head: lm_head is the embedding
    if res.missing_keys:
        m.tie_weights()
    rf = [t for t in gguf.GGUFReader(path).tensors if t.name == "rope_freqs.weight"]
    if rf:      # llama-3 frequency factors: theta_i / factor_i (ggml rope with freq_factors), not part of the HF state dict
        ff = torch.tensor(np.array(rf[0].data, dtype=np.float32))
        m.model.rotary_emb.inv_freq = m.model.rotary_emb.inv_freq / ff
        if hasattr(m.model.rotary_emb, "original_inv_freq"):
            m.model.rotary_emb.original_inv_freq = m.model.rotary_emb.inv_freq
    with torch.no_grad():
        return m(torch.tensor([list(tokens)])).logits[0].numpy().astype(np.float32)


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for b in iter(lambda: fh.read(1 << 20), b""):
            h.update(b)
    return h.hexdigest()


def main():
    from oracle import pyoracle as po

    tmp = "/tmp/blama_hf_golden"
    os.makedirs(tmp, exist_ok=True)
    out = {}
    for shape in SHAPES + BIG_SHAPES:
        path = os.path.join(tmp, shape + ".gguf")
        gs.write_gguf(path, shape)
        toks = np.array([int(t) for t in gs.synth_prompt(shape, N_TOK, 1)], dtype=np.int32)
        lg = hf_logits(path, toks)
        om = po.Model(path); oc = po.Ctx(om, 256, po.MODE_F32)
        ref = oc.decode(toks, all_logits=True)
        oc.close(); om.close()
        print(f"{shape}: HF logits std {lg.std():.3f}; oracle F32 mode max |d| {np.abs(lg - ref).max():.2e}, arg-max equal at "
              f"{int((lg.argmax(1) == ref.argmax(1)).sum())}/{N_TOK} positions")
        out[shape + "/tokens"] = toks
        out[shape + "/sha256"] = np.frombuffer(bytes.fromhex(sha256(path)), dtype=np.uint8)
        out[shape + "/argmax"] = lg.argmax(1).astype(np.int32)
        if shape in BIG_SHAPES:
            ids = np.argsort(-lg[-KEEP:], axis=1, kind="stable")[:, :TOPK].astype(np.int32)
            out[shape + "/top_ids"] = ids
            out[shape + "/top_logits"] = np.take_along_axis(lg[-KEEP:], ids, axis=1)
            os.remove(path)
        else:
            out[shape + "/last_logits"] = lg[-KEEP:]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hf_forward_golden.npz"), **out)


if __name__ == "__main__":
    main()
