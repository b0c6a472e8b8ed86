"""GPU parity of the whole decode path (embedding gather, fused RMSNorm+QKV GEMV+RoPE+KV append,
paged attention, O/gate-up/down GEMVs, lm_head, greedy sampler) against the numpy oracle, through
the C ABI (gl_decode_step / gl_prefill / gl_generate).

Stated tolerances: logits max-abs error <= 2e-3 * max|logit| vs the oracle in the engine's
activation format (act="i16", fp16 KV); <= 1e-2 * max|logit| vs the exact-activation oracle;
logprob |delta| <= 2e-2; token ids equal wherever the oracle's top-1/top-2 margin > 5e-2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(path, **kw):
    from gridllm_b200 import native as N
    return N.Engine(path, **kw)


@pytest.mark.parametrize("fixture", ["tiny_gguf", "tiny128_gguf", "tiny_q8_gguf", "tiny_f16_gguf"])
def test_decode_steps_match_oracle(fixture, request):
    from oracle import llama_oracle as O
    path = request.getfixturevalue(fixture)
    m = O.load_gguf(path)
    e = _engine(path)
    fp = fixture == "tiny_f16_gguf"
    orc = O.LlamaOracle(m, act="exact" if fp else "i16", kv_f16=True)
    orc_exact = O.LlamaOracle(m, act="exact", kv_f16=True)
    rng = np.random.Generator(np.random.PCG64(1000))
    toks = rng.integers(0, m.n_vocab - 3, size=40)      # crosses two 16-token KV pages
    for i, t in enumerate(toks):
        logits, am, lp = e.decode_step(int(t))
        ref = orc.step(int(t))
        ref2 = orc_exact.step(int(t))
        scale = np.abs(ref).max()
        assert np.isfinite(logits).all()
        assert np.abs(logits - ref).max() <= 2e-3 * scale, (fixture, i, np.abs(logits - ref).max(), scale)
        assert np.abs(logits - ref2).max() <= 1e-2 * scale, (fixture, i)
        srt = np.sort(ref)
        if srt[-1] - srt[-2] > 5e-2:
            assert am == int(np.argmax(ref))
        lse = ref.max() + np.log(np.exp(ref - ref.max()).sum())
        assert abs(lp - (ref[am] - lse)) <= 2e-2
    assert e.position() == len(toks)
    e.close()


def test_generate_matches_oracle(tiny_gguf):
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny_gguf)
    e = _engine(tiny_gguf, prefill_mode=1)     # sequential prefill: every position goes through the decode kernels
    orc = O.LlamaOracle(m, act="i16", kv_f16=True)
    for seed in (1000, 1001, 1002):
        prompt = np.random.Generator(np.random.PCG64(seed)).integers(0, m.n_vocab - 3, size=24)
        ref = orc.generate(prompt, 12)
        seen = []
        g = e.generate(prompt, num_predict=12, ignore_eos=True, want_logits=True,
                       on_token=lambda tid, lp, piece: seen.append(tid) and False)
        assert g.stats.prompt_eval_count == 24 and g.stats.eval_count == 12 and g.stats.done_reason == 1
        assert seen == list(g.ids)
        # compare along the oracle's own trajectory while ids agree
        for i in range(12):
            lg = e.last_logits(i)
            assert np.abs(lg - ref["logits"][i]).max() <= 2e-3 * np.abs(ref["logits"][i]).max(), (seed, i)
            assert abs(g.logprobs[i] - ref["logprobs"][i]) <= 2e-2
            if g.ids[i] != ref["ids"][i]:
                assert ref["margins"][i] <= 5e-2, (seed, i, ref["margins"][i])
                break
    e.close()


def test_prefill_then_decode_equals_stepwise(tiny128_gguf):
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny128_gguf)
    e = _engine(tiny128_gguf, prefill_mode=1)
    toks = np.random.Generator(np.random.PCG64(7)).integers(0, m.n_vocab - 3, size=37)
    a = e.prefill(toks)
    e.kv_reset()
    for t in toks:
        b, _, _ = e.decode_step(int(t))
    assert np.abs(a - b).max() <= 1e-2 * np.abs(b).max()
    orc = O.LlamaOracle(m, act="i16")
    for t in toks:
        ref = orc.step(int(t))
    assert np.abs(b - ref).max() <= 2e-3 * np.abs(ref).max()
    e.close()


def test_act8_mode_is_ggml_like(tiny_gguf):
    """act_bits=8 (ggml-style int8 activations) is further from the exact oracle than the default
    15-bit path but tracks the oracle's own q8 mode -- quantifies the Ollama-side divergence."""
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny_gguf)
    e8 = _engine(tiny_gguf, act_bits=8)
    orc8 = O.LlamaOracle(m, act="q8")
    toks = np.random.Generator(np.random.PCG64(11)).integers(0, m.n_vocab - 3, size=8)
    for t in toks:
        lg, _, _ = e8.decode_step(int(t))
        ref = orc8.step(int(t))
    assert np.abs(lg - ref).max() <= 2e-3 * np.abs(ref).max()
    e8.close()


def test_eos_stops_generation(tiny_gguf):
    """stop-id handling on the device: generation ends when a stop id is sampled."""
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny_gguf)
    e = _engine(tiny_gguf)
    prompt = np.random.Generator(np.random.PCG64(1000)).integers(0, m.n_vocab - 3, size=16)
    free = e.generate(prompt, num_predict=10, ignore_eos=True)
    assert len(set(free.ids.tolist())) > 1
    stop_at = 4
    g = e.generate(prompt, num_predict=10, ignore_eos=False, stop_ids=[int(free.ids[stop_at])])
    first = list(free.ids).index(int(free.ids[stop_at]))
    assert list(g.ids) == list(free.ids[:first])
    assert g.stats.done_reason == 0
    e.close()


@pytest.mark.parametrize("mode", [1, 0])
def test_embed_matches_oracle(tiny_gguf, mode):
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny_gguf)
    e = _engine(tiny_gguf, prefill_mode=mode)
    rng = np.random.Generator(np.random.PCG64(3000))
    seqs = [rng.integers(0, m.n_vocab - 3, size=n) for n in (5, 17, 1)]
    out, st = e.embed(seqs)
    orc = O.LlamaOracle(m, act="i16")
    for i, s in enumerate(seqs):
        ref = orc.embed(s)
        assert abs(np.linalg.norm(out[i]) - 1.0) < 1e-5
        assert np.abs(out[i] - ref).max() <= (2e-3 if mode == 1 else 5e-3)
    assert st.prompt_eval_count == 23
    e.close()


def test_context_overflow_is_an_error(tiny_gguf):
    from gridllm_b200 import native as N
    e = _engine(tiny_gguf, max_ctx=64)
    with pytest.raises(N.NativeError) as ei:
        e.generate(np.arange(60, dtype=np.int32), num_predict=16, ignore_eos=True)
    assert ei.value.code == -9
    e.close()


# ---- batched tensor-core prefill (fp16 inputs, fp32 accumulate) --------------------------------------
# Stated tolerance: logits after a batched prefill within 1e-2 * max|logit| of the exact-activation
# oracle and of the engine's own sequential (decode-kernel) prefill; token ids equal wherever the
# oracle's top-1/top-2 margin exceeds 5e-2.

@pytest.mark.parametrize("fixture", ["tiny_gguf", "tiny128_gguf", "tiny_q8_gguf", "tiny_f16_gguf"])
@pytest.mark.parametrize("n_tok", [8, 37, 128, 200])
def test_batched_prefill_matches_oracle(fixture, n_tok, request):
    from oracle import llama_oracle as O
    path = request.getfixturevalue(fixture)
    m = O.load_gguf(path)
    toks = np.random.Generator(np.random.PCG64(2000 + n_tok)).integers(0, m.n_vocab - 3, size=n_tok)
    eb = _engine(path, prefill_mode=0)
    lb = eb.prefill(toks)
    es = _engine(path, prefill_mode=1)
    ls = es.prefill(toks)
    orc = O.LlamaOracle(m, act="exact", kv_f16=True)
    for t in toks:
        ref = orc.step(int(t))
    scale = np.abs(ref).max()
    assert np.isfinite(lb).all()
    assert np.abs(lb - ref).max() <= 1e-2 * scale, (fixture, n_tok, np.abs(lb - ref).max(), scale)
    assert np.abs(lb - ls).max() <= 1e-2 * scale
    # the KV pages written by the prefill GEMMs must serve the decode kernels: continue stepwise
    nxt = int(np.argmax(ref))
    for _ in range(3):
        lg, am, _ = eb.decode_step(nxt)
        ref = orc.step(nxt)
        assert np.abs(lg - ref).max() <= 1e-2 * np.abs(ref).max()
        nxt = int(np.argmax(ref))
    assert eb.position() == n_tok + 3
    eb.close()
    es.close()


@pytest.fixture(scope="module")
def wide_ffn_gguf(tmp_models):
    """One layer, d = 256, n_ff = 6400: with a 300-token prompt the gate/up GEMM has 3 x 100 output tiles -- more than two
    per SM, so the persistent tcgen05 kernel reuses both halves of its double-buffered TMEM accumulator and its TMA ring
    runs on across tile boundaries."""
    from oracle import gguf_synth as S
    shape = S.LlamaShape("wide-ffn-synth", 1, 256, 4, 2, 6400, 512, 10000.0, 1e-5, 512)
    p = str(tmp_models / "wide_ffn.gguf")
    S.build_model(p, shape, "q4_k_m", seed=7)
    return p


@pytest.mark.parametrize("n_tok", [129, 300])
def test_batched_prefill_many_tiles_per_sm(wide_ffn_gguf, n_tok, monkeypatch):
    from oracle import llama_oracle as O
    m = O.load_gguf(wide_ffn_gguf)
    toks = np.random.Generator(np.random.PCG64(3000 + n_tok)).integers(0, m.n_vocab - 3, size=n_tok)
    orc = O.LlamaOracle(m, act="exact", kv_f16=True)
    for t in toks:
        ref = orc.step(int(t))
    scale = np.abs(ref).max()
    e = _engine(wide_ffn_gguf, prefill_mode=0)
    for _ in range(2):                                   # twice: barriers / TMEM are set up per launch
        e.kv_reset()
        lb = e.prefill(toks)
        assert np.isfinite(lb).all()
        assert np.abs(lb - ref).max() <= 1e-2 * scale, (n_tok, np.abs(lb - ref).max(), scale)
    e.close()
    # one tile per CTA (the non-persistent grid) must give the same numbers: same tiles, same K order
    monkeypatch.setenv("GL_PREFILL_TC5", "0")
    e2 = _engine(wide_ffn_gguf, prefill_mode=0)
    l2 = e2.prefill(toks)
    e2.close()
    assert np.abs(lb - l2).max() <= 1e-2 * scale


@pytest.mark.parametrize("fixture", ["tiny_gguf", "tiny128_gguf"])          # head dim 64 and 128
@pytest.mark.parametrize("n_tok", [37, 200, 300])
def test_fused_prompt_attention_equals_three_launch_path(fixture, n_tok, request, monkeypatch):
    """prefill_attn.cu (scores on the SM, online softmax) against the path it replaces (Q K^T GEMM -> causal softmax -> P V GEMM,
    GL_PREFILL_FLASH=0): same q / k / v bits in, P rounded to fp16 in both; stated tolerance 2e-3 * max|logit|."""
    from oracle import llama_oracle as O
    path = request.getfixturevalue(fixture)
    m = O.load_gguf(path)
    toks = np.random.Generator(np.random.PCG64(4000 + n_tok)).integers(0, m.n_vocab - 3, size=n_tok)
    monkeypatch.setenv("GL_PREFILL_FLASH", "1")
    e1 = _engine(path, prefill_mode=0)
    l1 = e1.prefill(toks)
    nxt = int(np.argmax(l1))
    d1, _, _ = e1.decode_step(nxt)                      # the KV rows the fused RoPE / split launch cached
    e1.close()
    monkeypatch.setenv("GL_PREFILL_FLASH", "0")
    e0 = _engine(path, prefill_mode=0)
    l0 = e0.prefill(toks)
    d0, _, _ = e0.decode_step(nxt)
    e0.close()
    scale = np.abs(l0).max()
    assert np.isfinite(l1).all()
    assert np.abs(l1 - l0).max() <= 2e-3 * scale, (fixture, n_tok, np.abs(l1 - l0).max(), scale)
    assert np.abs(d1 - d0).max() <= 2e-3 * np.abs(d0).max()


@pytest.mark.parametrize("n_tok", [37, 128, 300, 700])
def test_tcgen05_prompt_attention_equals_the_mma_sync_kernel(tiny128_gguf, n_tok, monkeypatch):
    """prefill_attn_tc5.cu (S, P and O in tensor memory, tcgen05.mma with the probabilities as the TMEM A operand,
    GL_PREFILL_ATTN_TC5=1) against prefill_attn.cu's mma.sync kernel on the same q / k / v bits: both round P to fp16 and sum the
    rounded values; KV tiles are 128 instead of 64 rows, so the online-softmax rescaling points differ -- 2e-3 * max|logit|.
    1 .. 6 query tiles: one KV tile (diagonal only), both TMEM score buffers in use, the third tile re-using the first buffer."""
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny128_gguf)
    toks = np.random.Generator(np.random.PCG64(8000 + n_tok)).integers(0, m.n_vocab - 3, size=n_tok)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GL_PREFILL_ATTN_TC5", mode)
        e = _engine(tiny128_gguf, prefill_mode=0, max_ctx=1024)
        for _ in range(2):                                   # twice: barriers / TMEM are set up per launch
            e.kv_reset()
            out[mode] = e.prefill(toks)
        e.close()
    scale = np.abs(out["0"]).max()
    assert np.isfinite(out["1"]).all()
    assert np.abs(out["1"] - out["0"]).max() <= 2e-3 * scale, (n_tok, np.abs(out["1"] - out["0"]).max(), scale)


@pytest.mark.parametrize("fixture", ["tiny_gguf", "tiny128_gguf"])
@pytest.mark.parametrize("n_tok", [37, 300])
def test_rope_split_in_the_qkv_epilogue_equals_the_kernel(fixture, n_tok, request, monkeypatch):
    """GEMM_EPI_ROPE_SPLIT (RoPE, Q / K / V^T split and the fp16 cache append out of the QKV projection's accumulator) against
    the stand-alone kernel reading the fp32 QKV matrix (GL_PREFILL_FUSE_ROPE=0): same arithmetic on the same accumulators --
    2e-3 * max|logit| covers a different fused-multiply-add contraction; the decode step that follows reads the cached K / V."""
    from oracle import llama_oracle as O
    path = request.getfixturevalue(fixture)
    m = O.load_gguf(path)
    toks = np.random.Generator(np.random.PCG64(7000 + n_tok)).integers(0, m.n_vocab - 3, size=n_tok)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GL_PREFILL_FUSE_ROPE", mode)
        e = _engine(path, prefill_mode=0)
        lg = e.prefill(toks)
        nxt = int(np.argmax(lg)) if mode == "1" else out["1"][2]
        dg, _, _ = e.decode_step(nxt)
        out[mode] = (lg, dg, nxt)
        e.close()
    scale = np.abs(out["0"][0]).max()
    assert np.isfinite(out["1"][0]).all() and np.isfinite(out["1"][1]).all()
    assert np.abs(out["1"][0] - out["0"][0]).max() <= 2e-3 * scale, (fixture, n_tok, np.abs(out["1"][0] - out["0"][0]).max(), scale)
    assert np.abs(out["1"][1] - out["0"][1]).max() <= 2e-3 * np.abs(out["0"][1]).max()


@pytest.mark.parametrize("n_tok", [129, 300, 520])
def test_cta_pair_gemm_equals_single_cta_gemm(wide_ffn_gguf, n_tok, monkeypatch):
    """tcgen05.mma.cta_group::2 (256 x 256 tiles on two SMs, GL_TC5_PAIR=2: wherever the shape allows) against the one-CTA kernel: same operands, same K
    order per output element.  Ragged M: the pair's second CTA owns rows past the prompt in the last tile."""
    from oracle import llama_oracle as O
    m = O.load_gguf(wide_ffn_gguf)
    toks = np.random.Generator(np.random.PCG64(6000 + n_tok)).integers(0, m.n_vocab - 3, size=n_tok)
    out = {}
    for mode in ("2", "0"):                               # 2 = pairs whenever the shape allows, 0 = never
        monkeypatch.setenv("GL_TC5_PAIR", mode)
        e = _engine(wide_ffn_gguf, prefill_mode=0, max_ctx=1024)
        for _ in range(2):
            e.kv_reset()
            out[mode] = e.prefill(toks)
        e.close()
    scale = np.abs(out["0"]).max()
    assert np.isfinite(out["2"]).all()
    assert np.abs(out["2"] - out["0"]).max() <= 1e-5 * scale, (n_tok, np.abs(out["2"] - out["0"]).max(), scale)


def test_generate_with_batched_prefill(tiny128_gguf):
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny128_gguf)
    e = _engine(tiny128_gguf)
    orc = O.LlamaOracle(m, act="exact", kv_f16=True)
    for seed in (1000, 1001):
        prompt = np.random.Generator(np.random.PCG64(seed)).integers(0, m.n_vocab - 3, size=150)
        ref = orc.generate(prompt, 8)
        g = e.generate(prompt, num_predict=8, ignore_eos=True, want_logits=True)
        assert g.stats.eval_count == 8 and g.stats.prompt_eval_count == 150
        for i in range(8):
            lg = e.last_logits(i)
            assert np.abs(lg - ref["logits"][i]).max() <= 1e-2 * np.abs(ref["logits"][i]).max(), (seed, i)
            assert abs(g.logprobs[i] - ref["logprobs"][i]) <= 2e-2
            if g.ids[i] != ref["ids"][i]:
                assert ref["margins"][i] <= 5e-2
                break
    e.close()


def test_long_context_decode(tiny128_gguf):
    """contexts beyond 512 tokens: more KV pages than attention splits, so a split owns a final page AND the page that
    holds the newest rows (two staging batches), splits own several pages, and the merge covers all 32 partials"""
    from oracle import llama_oracle as O
    m = O.load_gguf(tiny128_gguf)
    e = _engine(tiny128_gguf, max_ctx=1024)
    orc = O.LlamaOracle(m, act="exact", kv_f16=True)
    prompt = np.random.Generator(np.random.PCG64(77)).integers(0, m.n_vocab - 3, size=530)
    ref = orc.generate(prompt, 40)
    g = e.generate(prompt, num_predict=40, ignore_eos=True, want_logits=True)
    assert g.stats.eval_count == 40 and g.stats.prompt_eval_count == 530
    for i in range(40):
        lg = e.last_logits(i)
        assert np.isfinite(lg).all()
        assert np.abs(lg - ref["logits"][i]).max() <= 1e-2 * np.abs(ref["logits"][i]).max(), i
        if g.ids[i] != ref["ids"][i]:
            assert ref["margins"][i] <= 5e-2
            break
    e.close()


@pytest.mark.parametrize("fixture", ["tiny_gguf", "tiny128_gguf"])          # GQA group 2 / head dim 64 and group 4 / head dim 128
@pytest.mark.parametrize("splits", [8, 16])
def test_cluster_attention_matches_oracle(fixture, splits, request, monkeypatch):
    """GL_ATTN_CLUSTER=1: the splits of a KV head are one thread-block cluster (8 portable, 16 opt-in) and merge their partials
    through distributed shared memory instead of global partials + an atomic ticket.  Short contexts (idle splits still join the
    cluster barrier), a page boundary, and a context of more pages than one staging tile per split; same tolerances as the
    ticket path."""
    from oracle import llama_oracle as O
    path = request.getfixturevalue(fixture)
    m = O.load_gguf(path)
    monkeypatch.setenv("GL_ATTN_CLUSTER", "1")
    monkeypatch.setenv("GL_ATTN_SPLITS", str(splits))
    e = _engine(path)
    orc = O.LlamaOracle(m, act="i16", kv_f16=True)
    toks = np.random.Generator(np.random.PCG64(1000)).integers(0, m.n_vocab - 3, size=40)
    for i, t in enumerate(toks):
        logits, am, lp = e.decode_step(int(t))
        ref = orc.step(int(t))
        assert np.isfinite(logits).all()
        assert np.abs(logits - ref).max() <= 2e-3 * np.abs(ref).max(), (fixture, splits, i)
    e.close()
    # 700 cached positions = 44 pages: 3 - 6 pages per split, i.e. a second staging tile at 8 splits (the 512-context model: 450)
    n_long = 700 if fixture == "tiny128_gguf" else 450
    e = _engine(path, max_ctx=1024 if fixture == "tiny128_gguf" else 512)
    orc = O.LlamaOracle(m, act="exact", kv_f16=True)
    prompt = np.random.Generator(np.random.PCG64(78)).integers(0, m.n_vocab - 3, size=n_long)
    ref = orc.generate(prompt, 12)
    g = e.generate(prompt, num_predict=12, ignore_eos=True, want_logits=True)
    for i in range(12):
        lg = e.last_logits(i)
        assert np.isfinite(lg).all()
        assert np.abs(lg - ref["logits"][i]).max() <= 1e-2 * np.abs(ref["logits"][i]).max(), (fixture, splits, i)
        if g.ids[i] != ref["ids"][i]:
            assert ref["margins"][i] <= 5e-2
            break
    e.close()


@pytest.mark.parametrize("abits,warps,mega", [(16, 8, 1), (8, 8, 1), (16, 8, 0), (8, 8, 0), (16, 1, 0)])
def test_kernel_variants_match_oracle(tiny128_gguf, tiny_q8_gguf, abits, warps, mega, monkeypatch):
    if warps == 1:          # one CTA per SM with large stages instead of two with small ones
        monkeypatch.setenv("GL_CTAS_PER_SM", "1")
        warps = 8
    """every compiled (activation bits, consumer warps) variant of the GEMV core, inside the persistent kernel
    (mega=1) and as stand-alone per-op kernels under a CUDA graph with PDL (mega=0)"""
    from oracle import llama_oracle as O
    monkeypatch.setenv("GL_WARPS", str(warps))
    monkeypatch.setenv("GL_MEGA", str(mega))
    for path in (tiny128_gguf, tiny_q8_gguf):
        m = O.load_gguf(path)
        e = _engine(path, act_bits=abits, prefill_mode=1)
        orc = O.LlamaOracle(m, act="i16" if abits == 16 else "q8", kv_f16=True)
        toks = np.random.Generator(np.random.PCG64(77)).integers(0, m.n_vocab - 3, size=20)
        for t in toks:
            lg, am, lp = e.decode_step(int(t))
            ref = orc.step(int(t))
        # int8 activations are discontinuous (a rounding flip moves an activation by 1/127 of its block max): fp32-vs-fp64
        # differences upstream show at the 1e-2 level after 20 positions; the 15-bit default stays at 2e-3
        tol = 2e-3 if abits == 16 else 3e-2
        assert np.abs(lg - ref).max() <= tol * np.abs(ref).max(), (path, abits, warps, mega)
        g = e.generate(toks[:12], num_predict=6, ignore_eos=True)
        assert g.stats.eval_count == 6
        e.close()


# ---- cases shaped like BASELINE.json configs 4 and 5 at test size (both XPASSed on the round-1 driver box) ---------------------
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


# ---- the engine's own cancel path (job_cancellation, JobScheduler.ts:530-536 -> gl_token_cb returning non-zero) -----------------
def test_token_callback_cancels_the_real_engine(tiny_gguf):
    from gridllm_b200 import native as N
    e = _engine(tiny_gguf)
    prompt = np.random.Generator(np.random.PCG64(1000)).integers(0, e.info.n_vocab - 3, size=24)
    full = e.generate(prompt, num_predict=12, ignore_eos=True)
    seen = []

    def stop_at_three(tid, lp, piece):
        seen.append(tid)
        return len(seen) == 3                      # non-zero return at token 3

    g = e.generate(prompt, num_predict=12, ignore_eos=True, on_token=stop_at_three)
    assert g.stats.done_reason == 2 and g.stats.eval_count == 3 and len(seen) == 3
    assert list(g.ids) == list(full.ids[:3])
    # the raw status is GL_ERR_CANCELLED (the binding folds it into done_reason; check the C ABI itself once)
    import ctypes as C
    lib = N.load_library()
    so = N.SampleOpts()
    so.num_predict, so.ignore_eos, so.top_p = 12, 1, 1.0
    p = np.ascontiguousarray(prompt, dtype=np.int32)
    st = N.GenStats()
    n = [0]

    def raw_cb(_u, tid, lp, piece, plen):
        n[0] += 1
        return 1 if n[0] == 3 else 0
    rc = lib.gl_generate(e._h, N._i32p(p), len(p), C.byref(so), N.TOKEN_CB(raw_cb), None, None, None, C.byref(st))
    assert rc == N.GL_ERR_CANCELLED and st.done_reason == 2 and st.eval_count == 3
    # the engine is reusable afterwards and gives the same answer as before
    again = e.generate(prompt, num_predict=12, ignore_eos=True)
    assert list(again.ids) == list(full.ids)
    e.close()


# ---- RoPE variants that are also general.architecture == llama ----------------------------------------------
@pytest.mark.parametrize("variant", ["llama3_freqs", "linear"])
def test_rope_variants_match_oracle(tmp_models, variant):
    from oracle import gguf_synth as S, llama_oracle as O
    path = str(tmp_models / f"tiny128_rope_{variant}.gguf")
    kw = ({"rope_freqs": S.llama3_rope_factors(128, S.TINY128.rope_base, 8.0, 1.0, 4.0, 64)} if variant == "llama3_freqs"
          else {"rope_scaling": ("linear", 4.0)})
    S.build_model(path, S.TINY128, "q4_k_m", seed=4321, **kw)
    m = O.load_gguf(path)
    e = _engine(path)
    orc = O.LlamaOracle(m, act="i16", kv_f16=True)
    toks = np.random.Generator(np.random.PCG64(11)).integers(0, m.n_vocab - 3, size=40)
    ref = None
    for t in toks:
        ref = orc.step(int(t))
    lb = e.prefill(toks)                                   # batched prefill: rope_split kernel
    assert np.abs(lb - ref).max() <= 1e-2 * np.abs(ref).max()
    e.kv_reset()
    for t in toks:                                          # decode path: fused QKV epilogue
        lg, _, _ = e.decode_step(int(t))
    assert np.abs(lg - ref).max() <= 2e-3 * np.abs(ref).max()
    e.close()


def test_unsupported_rope_scaling_and_pretokenizer_fail_loudly(tmp_models):
    from gridllm_b200 import native as N
    from oracle import gguf_synth as S
    path = str(tmp_models / "tiny_yarn.gguf")
    S.build_model(path, S.TINY, "q4_k_m", seed=1, rope_scaling=("yarn", 4.0))
    with pytest.raises(N.NativeError) as ei:
        _engine(path)
    assert "rope.scaling.type" in str(ei.value)
    path = str(tmp_models / "tiny_tekken.gguf")
    S.build_model(path, S.TINY, "q4_k_m", seed=1, pre="tekken")
    e = _engine(path)
    assert e.info.has_tokenizer == 0                       # loads, but text requests are refused instead of mis-tokenised
    with pytest.raises(N.NativeError):
        e.tokenize("hello")
    e.close()


# ---- packed embeddings (/api/embed `input: string[]`, ollama.ts:574-643 -> OllamaService.ts:619-636) ----------
@pytest.mark.parametrize("fixture", ["tiny128_gguf", "tiny_q8_gguf"])
def test_packed_embeddings_ragged_batch_matches_oracle(fixture, request):
    """26 sequences of ragged lengths in ONE gl_embed call: they are packed 128-row-aligned into passes of <= 2048 rows (several
    packs here), share the linear layers and attend only inside themselves.  Every embedding equals the oracle's for that sequence
    alone (same tolerance as the single-sequence batched prefill), whatever its neighbours in the pack; order is preserved."""
    from oracle import llama_oracle as O
    path = request.getfixturevalue(fixture)
    m = O.load_gguf(path)
    rng = np.random.Generator(np.random.PCG64(5150))
    lens = [1, 9, 33, 2, 128, 129, 300, 7, 64, 255, 256, 257, 16, 3, 500, 31, 127, 200, 5, 90, 12, 384, 1, 77, 140, 20]
    seqs = [rng.integers(0, m.n_vocab - 3, size=n) for n in lens]
    e = _engine(path, max_ctx=512)
    out, st = e.embed(seqs)
    assert out.shape == (len(seqs), m.n_embd) and st.prompt_eval_count == sum(lens) and st.prompt_eval_duration_ns > 0
    orc = O.LlamaOracle(m, act="exact")
    worst = 0.0
    for i, s in enumerate(seqs):
        ref = orc.embed(s)
        assert abs(np.linalg.norm(out[i]) - 1.0) < 1e-5
        worst = max(worst, float(np.abs(out[i] - ref).max()))
        assert np.abs(out[i] - ref).max() <= 5e-3, (fixture, i, lens[i])
    # the same sequence alone, or first / last in another call: the same bits (its rows never see a neighbour)
    alone, _ = e.embed([seqs[6]])
    other, _ = e.embed([seqs[20], seqs[6], seqs[3]])
    assert np.array_equal(alone[0], out[6]) and np.array_equal(other[1], out[6])
    e.close()
