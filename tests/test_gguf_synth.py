"""The synthetic GGUF writer produces files other GGUF readers accept, with the BASELINE model sizes."""
import numpy as np
import pytest

from blama_b200 import gguf_synth as gs


def test_sizes_match_survey():
    # file sizes implied by shapes x block sizes
    assert abs(gs.model_bytes(gs.SHAPES["llama-3.2-1b-q8"]) / 1e9 - 1.313) < 0.01
    assert abs(gs.model_bytes(gs.SHAPES["llama-3.1-8b-q4km"]) / 1e9 - 4.913) < 0.01
    assert abs(gs.model_bytes(gs.SHAPES["qwen2.5-7b-q8"]) / 1e9 - 8.093) < 0.01
    assert abs(gs.model_bytes(gs.SHAPES["llama-3.1-70b-q4km"]) / 1e9 - 42.51) < 0.05


def test_q4_k_m_type_mix():
    s = gs.SHAPES["llama-3.1-8b-q4km"]
    types = {name: t for name, _, t, _, _ in gs.plan_tensors(s)}
    assert types["output.weight"] == gs.Q6_K and types["token_embd.weight"] == gs.Q4_K
    assert types["blk.0.attn_v.weight"] == gs.Q6_K and types["blk.4.attn_v.weight"] == gs.Q4_K
    assert types["blk.6.ffn_down.weight"] == gs.Q6_K and types["blk.5.ffn_down.weight"] == gs.Q4_K
    assert types["blk.3.attn_q.weight"] == gs.Q4_K and types["blk.3.attn_norm.weight"] == gs.F32
    s70 = gs.SHAPES["llama-3.1-70b-q4km"]
    t70 = {name: t for name, _, t, _, _ in gs.plan_tensors(s70)}
    assert t70["blk.11.attn_v.weight"] == gs.Q5_K


def test_readable_by_gguf_py(gguf_path):
    import gguf

    r = gguf.GGUFReader(gguf_path("tiny-qwen2-q8"))
    names = [t.name for t in r.tensors]
    assert "blk.1.attn_q.bias" in names and "output.weight" in names
    assert r.fields["general.architecture"].contents() == "qwen2"
    assert r.fields["qwen2.attention.head_count_kv"].contents() == 2
    t = next(t for t in r.tensors if t.name == "blk.0.ffn_down.weight")
    w = gguf.quants.dequantize(t.data, t.tensor_type)
    assert w.shape == (256, 768) and 0.005 < float(w.std()) < 0.05


def test_prompts_are_seeded_and_avoid_special_ids():
    a = gs.synth_prompt("llama-3.1-8b-q4km", 512, 1)
    b = gs.synth_prompt("llama-3.1-8b-q4km", 512, 1)
    assert np.array_equal(a, b) and a.min() >= 0 and a.max() < 128256
    assert not set(a.tolist()) & set(gs.special_tokens(gs.SHAPES["llama-3.1-8b-q4km"]).values())
