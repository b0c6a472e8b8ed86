"""GPU parity cases added AFTER this round's GPU budget was spent: they have never run on a B200.  They are marked
xfail(strict=False) so that an unexpected failure cannot mask the verified suite (this file sorts last); an XPASS at the
next run is the signal to move them into tests/test_gpu_decode.py without the marker.

They cover the two combinations shaped like BASELINE.json configs 4 and 5 that the verified suite does not:
a bf16 model through the batched tensor-core prefill + decode, and an all-Q8_0 model through gl_embed."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent: never run on a GPU")]


def _engine(path, **kw):
    from gridllm_b200 import native as N
    return N.Engine(path, **kw)


def test_embed_q8_model_matches_oracle(tiny_q8_gguf):
    """config 5 at test size: all-Q8_0 weights, ragged sequences, prefill -> output_norm -> mean pool -> L2 normalise"""
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny_q8_gguf)
    rng = np.random.Generator(np.random.PCG64(3100))
    seqs = [rng.integers(0, m.n_vocab - 3, size=n) for n in (9, 33, 2)]
    orc = O.LlamaOracle(m, act="i16")
    for mode in (1, 0):
        e = _engine(tiny_q8_gguf, prefill_mode=mode)
        out, st = e.embed(seqs)
        for i, s in enumerate(seqs):
            ref = orc.embed(s)
            assert abs(np.linalg.norm(out[i]) - 1.0) < 1e-5
            assert np.abs(out[i] - ref).max() <= (2e-3 if mode == 1 else 5e-3)
        assert st.prompt_eval_count == 44
        e.close()


def test_bf16_model_prefill_and_decode(tmp_models):
    """config 4 at test size: bf16 weights through the batched tensor-core prefill, then decode steps on its KV pages"""
    from oracle import gguf_synth as S, llama_oracle as O
    path = str(tmp_models / "tiny_bf16.gguf")
    S.build_model(path, S.TINY, "bf16", seed=55)
    m = O.load_gguf(path)
    toks = np.random.Generator(np.random.PCG64(4000)).integers(0, m.n_vocab - 3, size=130)
    orc = O.LlamaOracle(m, act="exact", kv_f16=True)
    for t in toks:
        ref = orc.step(int(t))
    e = _engine(path, prefill_mode=0)
    lb = e.prefill(toks)
    assert np.isfinite(lb).all()
    assert np.abs(lb - ref).max() <= 1e-2 * np.abs(ref).max()
    nxt = int(np.argmax(ref))
    for _ in range(3):
        lg, am, _ = e.decode_step(nxt)
        ref = orc.step(nxt)
        assert np.abs(lg - ref).max() <= 1e-2 * np.abs(ref).max()
        nxt = int(np.argmax(ref))
    e.close()
