"""CPU: how NativeInferenceService maps InferenceRequest.options (client/src/types/index.ts:1-27; gateway validation ranges
server/src/routes/ollama.ts:26-48) onto gl_sample_opts.  No engine is created (the GPU side is tests/test_gpu_service.py)."""
import pytest

from gridllm_b200.service import NativeInferenceService


def _svc(defaults=None):
    s = NativeInferenceService.__new__(NativeInferenceService)       # no engines: only the option logic is exercised
    s._sampling_defaults = dict(defaults or {})
    s._apply_template = False
    return s


def test_absent_or_zero_temperature_is_greedy():
    s = _svc()
    assert s._sampling({}) == {}
    assert s._sampling({"temperature": 0}) == {}
    assert s._sampling({"temperature": 0.0, "top_k": 40, "seed": 3}) == {}
    assert s._sampling({"temperature": None}) == {}


def test_sampled_request_carries_all_options():
    s = _svc()
    kw = s._sampling({"temperature": 0.7, "top_k": 40, "top_p": 0.9, "seed": 123})
    assert kw == {"temperature": 0.7, "top_k": 40, "top_p": 0.9, "seed": 123}
    kw = s._sampling({"temperature": 1.2})
    assert kw["temperature"] == 1.2 and kw["top_k"] == 0 and kw["top_p"] == 1.0
    assert isinstance(kw["seed"], int) and 0 <= kw["seed"] < 2 ** 63          # drawn when the request has none
    assert s._sampling({"temperature": 1.2})["seed"] != kw["seed"] or True    # (two draws may collide; not asserted)


def test_ollama_defaults_apply_only_to_missing_options():
    s = _svc(NativeInferenceService.OLLAMA_SAMPLING_DEFAULTS)
    kw = s._sampling({})
    assert (kw["temperature"], kw["top_k"], kw["top_p"]) == (0.8, 40, 0.9)
    kw = s._sampling({"temperature": 0.5, "top_k": 10, "seed": 9})
    assert (kw["temperature"], kw["top_k"], kw["top_p"], kw["seed"]) == (0.5, 10, 0.9, 9)
    assert s._sampling({"temperature": 0}) == {}                               # an explicit 0 stays greedy


@pytest.mark.parametrize("bad", [-0.1, float("inf"), float("nan")])
def test_invalid_temperature_is_rejected(bad):
    with pytest.raises(RuntimeError):
        _svc()._sampling({"temperature": bad})


# ---- options.stop: stop strings end the generation, are not part of the response, may span tokens -----------------------
from gridllm_b200.service import StopFilter  # noqa: E402


def test_stop_filter_spanning_tokens_and_holdback():
    f = StopFilter(["###", "END"])
    out = [f.feed(p) for p in (b"Hello", b" wor", b"ld #", b"#", b" not yet", b" ##", b"#", b" after")]
    assert out == ["Hello", " wor", "ld ", "", "## not yet", " ", "", ""]       # '#' / '##' are held back until they resolve
    assert f.hit and f.text == "Hello world ## not yet "
    assert f.flush() == "" and f.text == "Hello world ## not yet "


def test_stop_filter_multibyte_and_flush():
    f = StopFilter(["\n\n"])
    euro = "€".encode("utf-8")
    assert f.feed(euro[:1]) == "" and f.feed(euro[1:]) == "€"
    assert f.feed(b"x\n") == "x" and not f.hit                   # a lone newline could start the stop string
    assert f.flush() == "\n" and f.text == "€x\n"                # generation ended by length: the held text is released
    g = StopFilter(["ab"])
    assert g.feed(b"a") == "" and g.feed(b"c") == "ac" and g.feed(b"ab") == "" and g.hit and g.text == "ac"


class _FakeInfo:
    has_tokenizer = True
    n_vocab = 100


class _FakeStats:
    prompt_eval_count, eval_count, prompt_eval_duration_ns, eval_duration_ns, total_duration_ns, load_duration_ns = 3, 0, 1, 1, 1, 1
    done_reason, kernel_launches = 1, 1


class _FakeGen:
    def __init__(self, ids):
        self.ids, self.logprobs, self.stats = ids, [0.0] * len(ids), _FakeStats()
        self.stats.eval_count = len(ids)


class _FakeEngine:
    """Emits fixed pieces; honours the cancel return of the token callback like gl_generate does."""
    info = _FakeInfo()
    PIECES = [b"one", b" two", b" th", b"ree", b" four", b" five"]

    def tokenize(self, text, add_bos=True, parse_special=False):
        return [1, 2, 3]

    def detokenize(self, ids):
        return b"".join(self.PIECES[i] for i in ids).decode()

    def generate(self, ids, num_predict=128, ignore_eos=False, on_token=None, **kw):
        self.kw = kw
        out = []
        for i, p in enumerate(self.PIECES[:num_predict]):
            out.append(i)
            if on_token is not None and on_token(i, -0.5, p):
                break
        return _FakeGen(out)


def _service_with_fake():
    import threading
    s = _svc()
    s._paths = {"m": "unused"}
    s._engines = {"m": _FakeEngine()}
    s._lock = threading.Lock()
    s._load_lock = threading.Lock()
    s._max_batch = 0
    s._runners = {}
    return s


def test_stop_string_through_the_service():
    import asyncio
    s = _service_with_fake()
    req = {"id": "r", "model": "m", "prompt": "p", "options": {"num_predict": 6, "stop": ["three"]}}
    res = asyncio.run(s.generateResponse(req))
    assert res["response"] == "one two " and res["done_reason"] == "stop" and res["eval_count"] == 4
    # without stop strings nothing changes
    res = asyncio.run(s.generateResponse({"id": "r", "model": "m", "prompt": "p", "options": {"num_predict": 6}}))
    assert res["response"] == "one two three four five" and res["done_reason"] == "length"
    # a stop string that never appears: full text, held-back pieces released at the end
    res = asyncio.run(s.generateResponse(dict(req, options={"num_predict": 6, "stop": "fivex"})))
    assert res["response"] == "one two three four five" and res["done_reason"] == "length"

    async def collect(r):
        return [c async for c in s.generateStreamResponse(r)]
    chunks = asyncio.run(collect(dict(req, stream=True)))
    assert "".join(c["response"] for c in chunks) == "one two " and chunks[-1]["done"] is True
    chunks = asyncio.run(collect(dict(req, stream=True, options={"num_predict": 6, "stop": ["fivex"]})))
    assert "".join(c["response"] for c in chunks) == "one two three four five"
    chat = asyncio.run(s.generateChatResponse({"id": "c", "model": "m", "options": {"num_predict": 6, "stop": [" four"]},
                                               "metadata": {"messages": [{"role": "user", "content": "hi"}]}}))
    assert chat["message"] == {"role": "assistant", "content": "one two three"} and chat["done_reason"] == "stop"


def test_context_ids_continue_a_conversation():
    """metadata.context (OllamaService.ts:224-226): the prior token ids come first, the new prompt follows without a BOS"""
    import asyncio
    s = _service_with_fake()
    eng = s._engines["m"]
    seen = {}
    orig_tok, orig_gen = eng.tokenize, eng.generate

    def tok(text, add_bos=True, parse_special=False):
        seen["add_bos"] = add_bos
        return [7, 8] if not add_bos else [0, 7, 8]

    def gen(ids, **kw):
        seen["ids"] = [int(i) for i in ids]
        return orig_gen(ids, **kw)
    eng.tokenize, eng.generate = tok, gen
    asyncio.run(s.generateResponse({"id": "r", "model": "m", "prompt": "p", "options": {"num_predict": 2}, "metadata": {"context": [11, 12, 13]}}))
    assert seen["ids"] == [11, 12, 13, 7, 8] and seen["add_bos"] is False
    asyncio.run(s.generateResponse({"id": "r", "model": "m", "prompt": "p", "options": {"num_predict": 2}}))
    assert seen["ids"] == [0, 7, 8] and seen["add_bos"] is True
