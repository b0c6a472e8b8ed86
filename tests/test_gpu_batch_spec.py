"""SPECIFICATION (not yet implemented): continuous batching inside one engine -- DESIGN.md section 10, SURVEY.md section 8f.1.
These tests state the contract of the planned entry points

    gl_seq_open(engine, prompt, n_prompt, opts, &slot)      prefill into the slot's own KV pages, first token pending
    gl_batch_step(engine, slots[], ids[], logprobs[], done[], cap, &n)   one token for every open slot
    gl_seq_close(engine, slot)                              pages back to the pool

through the Python binding methods Engine.seq_open / Engine.batch_step / Engine.seq_close.  They are skipped until the library
exports gl_seq_open, so they cost nothing today and become the parity bar of the feature the day it lands:
  * a sequence's tokens do not depend on who shares its batch (bit-identical ids to gl_generate on the same engine);
  * sequences join and leave between steps; closing frees pages; the pool's capacity is an error, not a crash;
  * per-sequence sampling options (greedy and seeded top-k) are honoured inside a batch."""
import numpy as np
import pytest


def _has_batching():
    try:
        from gridllm_b200 import native as N
        return hasattr(N.load_library(), "gl_seq_open")
    except Exception:
        return False


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _has_batching(), reason="gl_seq_open is not implemented yet (DESIGN.md section 10)")]


def _drain(e, want):
    """step until every slot in `want` (slot -> number of tokens) has produced its tokens; returns slot -> ids"""
    out = {s: [] for s in want}
    while any(len(out[s]) < want[s] for s in want):
        for slot, tok, _lp, done in e.batch_step():
            if slot in out and len(out[slot]) < want[slot]:
                out[slot].append(int(tok))
    return out


def test_batched_tokens_equal_single_sequence_generation(tiny128_gguf):
    from gridllm_b200 import native as N
    e = N.Engine(tiny128_gguf)
    prompts = [np.random.Generator(np.random.PCG64(100 + i)).integers(0, e.info.n_vocab - 3, size=n) for i, n in enumerate((5, 40, 17, 130))]
    ref = [list(e.generate(p, num_predict=12, ignore_eos=True).ids) for p in prompts]
    slots = [e.seq_open(p, num_predict=12, ignore_eos=True) for p in prompts]
    got = _drain(e, {s: 12 for s in slots})
    for s, r in zip(slots, ref):
        assert got[s] == [int(t) for t in r]            # who shares the batch does not change a sequence
        e.seq_close(s)
    e.close()


def test_sequences_join_and_leave_between_steps(tiny128_gguf):
    from gridllm_b200 import native as N
    e = N.Engine(tiny128_gguf)
    rng = np.random.Generator(np.random.PCG64(7))
    p = [rng.integers(0, e.info.n_vocab - 3, size=20) for _ in range(3)]
    ref = [list(e.generate(x, num_predict=10, ignore_eos=True).ids) for x in p]
    a = e.seq_open(p[0], num_predict=10, ignore_eos=True)
    out = {a: []}
    for _ in range(4):
        for slot, tok, _lp, _d in e.batch_step():
            out[slot].append(int(tok))
    b = e.seq_open(p[1], num_predict=10, ignore_eos=True)          # joins while a is mid-way
    out[b] = []
    for _ in range(3):
        for slot, tok, _lp, _d in e.batch_step():
            out[slot].append(int(tok))
    e.seq_close(a)                                                  # leaves early
    c = e.seq_open(p[2], num_predict=10, ignore_eos=True)          # may reuse a's pages
    out[c] = []
    while len(out[b]) < 10 or len(out[c]) < 10:
        for slot, tok, _lp, _d in e.batch_step():
            if len(out[slot]) < 10:
                out[slot].append(int(tok))
    assert out[a] == [int(t) for t in ref[0][:7]] and out[b] == [int(t) for t in ref[1]] and out[c] == [int(t) for t in ref[2]]
    e.close()


def test_per_sequence_sampling_and_capacity(tiny_gguf):
    from gridllm_b200 import native as N
    e = N.Engine(tiny_gguf)
    p = np.random.Generator(np.random.PCG64(9)).integers(0, e.info.n_vocab - 3, size=16)
    ref_greedy = list(e.generate(p, num_predict=8, ignore_eos=True).ids)
    ref_sampled = list(e.generate(p, num_predict=8, ignore_eos=True, temperature=0.8, top_k=40, top_p=0.9, seed=5).ids)
    g = e.seq_open(p, num_predict=8, ignore_eos=True)
    s = e.seq_open(p, num_predict=8, ignore_eos=True, temperature=0.8, top_k=40, top_p=0.9, seed=5)
    got = _drain(e, {g: 8, s: 8})
    assert got[g] == [int(t) for t in ref_greedy] and got[s] == [int(t) for t in ref_sampled]
    # the slot table / page pool is finite: running out is an error the caller can handle
    opened = [g, s]
    with pytest.raises(N.NativeError):
        for _ in range(10000):
            opened.append(e.seq_open(p, num_predict=8, ignore_eos=True))
    for slot in opened:
        e.seq_close(slot)
    again = e.seq_open(p, num_predict=8, ignore_eos=True)           # everything was returned to the pool
    assert _drain(e, {again: 8})[again] == [int(t) for t in ref_greedy]
    e.close()
