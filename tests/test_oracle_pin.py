"""CPU: pins the oracle.  The reference holds no golden vectors for this path (SURVEY.md section 8c:
parity unpinned), so the oracle is anchored on the two independent implementations present in this
image: gguf-py's dequantisers (block formats) and transformers' LlamaForCausalLM (forward pass)."""
import numpy as np
import pytest


def test_dequant_is_bit_exact_with_gguf_py():
    from gguf import quants, GGMLQuantizationType as T
    from oracle import gguf_synth as S, llama_oracle as O
    rng = np.random.Generator(np.random.PCG64(1234))
    for t, gt in ((S.Q4_K, T.Q4_K), (S.Q6_K, T.Q6_K), (S.Q8_0, T.Q8_0)):
        b = S.random_blocks(rng, t, 8, 1024)
        assert np.array_equal(O.dequantize(b, t, (8, 1024)), quants.dequantize(b, gt).reshape(8, 1024))


def _unpermute(w, n_head):
    """GGUF (llama.cpp convert) stores q/k rows interleaved per head: [head][hd/2][2]; HF wants
    [head][2][hd/2] (rotate_half convention)."""
    rows, cols = w.shape
    hd = rows // n_head
    return w.reshape(n_head, hd // 2, 2, cols).swapaxes(1, 2).reshape(rows, cols)


@pytest.mark.parametrize("shape_name,rope", [("TINY", None), ("TINY128", None), ("TINY128", "llama3"), ("TINY", "linear")])
def test_forward_matches_transformers(tmp_models, shape_name, rope):
    """rope None: plain RoPE.  "llama3": rope_freqs.weight (Llama-3.1 / 3.2 frequency factors) against transformers'
    rope_type llama3 built from the same parameters.  "linear": {arch}.rope.scaling.type linear against rope_type linear."""
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    from oracle import gguf_synth as S, llama_oracle as O
    shape = getattr(S, shape_name)
    path = str(tmp_models / f"pin_{shape_name}_{rope}_f32.gguf")
    kw, rope_cfg = {}, None
    if rope == "llama3":
        # original context 64 so that the 19 test positions see all three regimes of the factor ramp
        kw["rope_freqs"] = S.llama3_rope_factors(shape.head_dim, shape.rope_base, 8.0, 1.0, 4.0, 64)
        assert len(set(kw["rope_freqs"].tolist())) > 2
        rope_cfg = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                    "original_max_position_embeddings": 64, "rope_theta": shape.rope_base}
    elif rope == "linear":
        kw["rope_scaling"] = ("linear", 4.0)
        rope_cfg = {"rope_type": "linear", "factor": 4.0, "rope_theta": shape.rope_base}
    S.build_model(path, shape, "f32", seed=2024, **kw)
    m = O.load_gguf(path)
    cfg = tr.LlamaConfig(vocab_size=shape.n_vocab, hidden_size=shape.n_embd, intermediate_size=shape.n_ff,
                         num_hidden_layers=shape.n_layer, num_attention_heads=shape.n_head,
                         num_key_value_heads=shape.n_head_kv, head_dim=shape.head_dim, rms_norm_eps=shape.rms_eps,
                         rope_theta=shape.rope_base, max_position_embeddings=shape.n_ctx, tie_word_embeddings=False,
                         attention_bias=False, mlp_bias=False, hidden_act="silu")
    if rope_cfg is not None:
        try:
            cfg = tr.LlamaConfig(**{**cfg.to_dict(), "rope_scaling": rope_cfg, "rope_parameters": rope_cfg})
        except Exception:
            cfg.rope_scaling = rope_cfg
    try:
        cfg._attn_implementation = "eager"
    except Exception:
        pass
    hf = tr.LlamaForCausalLM(cfg).to(torch.float64).eval()
    sd = {"model.embed_tokens.weight": m.w("token_embd.weight"), "model.norm.weight": m.w("output_norm.weight").reshape(-1),
          "lm_head.weight": m.w("output.weight")}
    for il in range(shape.n_layer):
        p, h = f"blk.{il}.", f"model.layers.{il}."
        sd[h + "input_layernorm.weight"] = m.w(p + "attn_norm.weight").reshape(-1)
        sd[h + "post_attention_layernorm.weight"] = m.w(p + "ffn_norm.weight").reshape(-1)
        sd[h + "self_attn.q_proj.weight"] = _unpermute(m.w(p + "attn_q.weight"), shape.n_head)
        sd[h + "self_attn.k_proj.weight"] = _unpermute(m.w(p + "attn_k.weight"), shape.n_head_kv)
        sd[h + "self_attn.v_proj.weight"] = m.w(p + "attn_v.weight")
        sd[h + "self_attn.o_proj.weight"] = m.w(p + "attn_output.weight")
        sd[h + "mlp.gate_proj.weight"] = m.w(p + "ffn_gate.weight")
        sd[h + "mlp.up_proj.weight"] = m.w(p + "ffn_up.weight")
        sd[h + "mlp.down_proj.weight"] = m.w(p + "ffn_down.weight")
    missing, unexpected = hf.load_state_dict({k: torch.tensor(np.ascontiguousarray(v), dtype=torch.float64) for k, v in sd.items()}, strict=False)
    assert not [k for k in missing if "rotary" not in k] and not unexpected
    toks = np.random.Generator(np.random.PCG64(77)).integers(0, shape.n_vocab, size=19)
    with torch.no_grad():
        ref = hf(torch.tensor(toks[None, :])).logits[0].numpy()
    orc = O.LlamaOracle(m, act="exact", kv_f16=False)
    for i, t in enumerate(toks):
        lg = orc.step(int(t))
        assert np.abs(lg - ref[i]).max() <= 1e-5 * max(1.0, np.abs(ref[i]).max()), (i, np.abs(lg - ref[i]).max())
