"""CPU: how NativeInferenceService maps InferenceRequest.options (client/src/types/index.ts:1-27; gateway validation ranges
server/src/routes/ollama.ts:26-48) onto gl_sample_opts.  No engine is created (the GPU side is tests/test_gpu_service.py)."""
import pytest

from gridllm_b200.service import NativeInferenceService


def _svc(defaults=None):
    s = NativeInferenceService.__new__(NativeInferenceService)       # no engines: only the option logic is exercised
    s._sampling_defaults = dict(defaults or {})
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
