"""CPU: the worker host speaks the reference's frozen pub/sub wire format.  The
engine is replaced by a test double here (no GPU); tests/test_gpu_service.py runs the real thing.
The checks mirror what the reference's only test does (tests/integration/integration.ts:6-35: same
key sets and typeof per key) but against golden shapes derived from the reference sources."""
import asyncio
import json
import os

import pytest

from conftest import ROOT

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "wire_format.json")))
PYT = {"str": str, "int": int, "dict": dict, "list": list, "bool": bool}


def _check(obj, spec, allow_extra=False):
    for k, t in spec.items():
        assert k in obj, f"missing key {k}"
        assert isinstance(obj[k], PYT[t]) and not (t == "int" and isinstance(obj[k], bool)), (k, type(obj[k]))
    if not allow_extra:
        assert set(obj) == set(spec), (sorted(obj), sorted(spec))


class FakeService:
    """Same 8-method surface as OllamaService / NativeInferenceService."""

    def __init__(self, fail=False):
        self.fail = fail
        self.calls = []

    async def checkHealth(self): return True
    async def getAvailableModels(self):
        return [{"name": "m:latest", "digest": "abc", "size": 10, "modified_at": "2026-01-01T00:00:00Z",
                 "details": {"format": "gguf", "family": "llama", "families": ["llama"], "parameter_size": "0.0B", "quantization_level": "Q4_K_M"}}]
    async def validateModel(self, name): return name == "m:latest"

    def _resp(self, r):
        return {"id": r["id"], "model": r["model"], "created_at": "2026-01-01T00:00:00.000Z", "response": "hi", "done": True, "done_reason": "length",
                "total_duration": 1, "load_duration": 1, "prompt_eval_count": 3, "prompt_eval_duration": 1, "eval_count": 2, "eval_duration": 1,
                "system_fingerprint": "fp"}

    async def generateResponse(self, r):
        self.calls.append("generateResponse")
        if self.fail:
            raise RuntimeError("Inference failed: boom")
        return self._resp(r)

    async def generateStreamResponse(self, r):
        self.calls.append("generateStreamResponse")
        for i, t in enumerate(["a", "b", ""]):
            yield {"id": r["id"], "response": t, "done": i == 2}

    async def generateChatResponse(self, r):
        self.calls.append("generateChatResponse")
        x = self._resp(r); x.pop("response"); x["message"] = {"role": "assistant", "content": "hi"}
        return x

    async def generateChatStreamResponse(self, r):
        self.calls.append("generateChatStreamResponse")
        for i, t in enumerate(["x", ""]):
            yield {"id": r["id"], "response": t, "done": i == 1}

    async def generateEmbedding(self, r):
        self.calls.append("generateEmbedding")
        return {"id": r["id"], "model": r["model"], "embeddings": [[0.0, 1.0]], "total_duration": 1, "load_duration": 1, "prompt_eval_count": 2}


def _assignment(jid, **kw):
    req = {"id": jid, "model": "m:latest", "prompt": "p", "stream": False, "options": {}, "priority": "medium", "timeout": 300000, "metadata": {}}
    req.update(kw)
    return json.dumps({"type": "job_assignment", "job": {"jobId": jid, "workerId": "w0", "request": req, "assignedAt": "2026-01-01T00:00:00.000Z", "timeout": 300000}})


def _run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


def test_registration_heartbeat_status_shapes():
    from gridllm_b200.worker import LocalBus, NativeWorker
    bus = LocalBus()
    w = NativeWorker("w0", FakeService(), bus)
    _run(w.start())
    _run(w.sendHeartbeat())
    _run(w.publishStatusUpdate())
    ch = dict((c, json.loads(m)) for c, m in bus.log)
    reg = ch["worker:registered"]
    _check(reg, GOLD["worker:registered"]["keys"])
    _check(reg["capabilities"], GOLD["worker:registered"]["capabilities_keys"])
    _check(reg["capabilities"]["availableModels"][0], GOLD["worker:registered"]["model_keys"])
    assert reg["status"] == "online" and reg["capabilities"]["supportedFormats"] == ["json", "text"]
    assert json.loads(bus.hashes["workers"]["w0"]) == reg                       # HSET workers <id> (WorkerClientService.ts:188-192)
    _check(ch["worker:heartbeat"], GOLD["worker:heartbeat"]["keys"])
    assert "heartbeat:w0" in bus.keys
    _check(ch["worker:status_update"], GOLD["worker:status_update"]["keys"])
    assert ch["worker:status_update"] == {"workerId": "w0", "status": "online", "currentJobs": 0}


@pytest.mark.parametrize("kind,expect_call", [
    ("generate", "generateResponse"), ("stream", "generateStreamResponse"), ("chat", "generateChatResponse"),
    ("chat_stream", "generateChatStreamResponse"), ("embedding", "generateEmbedding")])
def test_dispatch_and_result_messages(kind, expect_call):
    from gridllm_b200.worker import LocalBus, NativeWorker
    bus = LocalBus()
    svc = FakeService()
    w = NativeWorker("w0", svc, bus)
    _run(w.start())
    bus.log.clear()
    kw = {"generate": {}, "stream": {"stream": True}, "chat": {"metadata": {"requestType": "chat", "messages": [{"role": "user", "content": "q"}]}},
          "chat_stream": {"stream": True, "metadata": {"requestType": "chat", "messages": [{"role": "user", "content": "q"}]}},
          "embedding": {"input": ["doc"], "metadata": {"requestType": "embedding"}}}[kind]
    _run(bus.publish("worker:w0:job", _assignment("job-1", **kw)))
    assert svc.calls == [expect_call]                                         # type dispatch, WorkerClientService.ts:545-646
    chans = [c for c, _ in bus.log if c != "worker:w0:job"]
    assert chans[0] == "worker:status_update" and chans[-1] == "worker:status_update"
    assert json.loads(bus.log[1][1])["status"] == "busy" and json.loads(bus.log[-1][1]) == {"workerId": "w0", "status": "online", "currentJobs": 0}
    assert chans[-3:-1] == ["job:completed", "job:result:job-1"]
    done = json.loads([m for c, m in bus.log if c == "job:completed"][0])
    _check(done, GOLD["job:completed"]["keys"])
    assert done == json.loads([m for c, m in bus.log if c == "job:result:job-1"][0])
    streams = [json.loads(m) for c, m in bus.log if c == "job:stream:job-1"]
    if kind == "stream":
        assert len(streams) == 3
        for s in streams:
            _check(s, GOLD["job:stream"]["keys"]); _check(s["chunk"], GOLD["job:stream"]["chunk_keys"])
        assert done["result"] == {"id": "job-1", "response": "ab", "done": True}
    elif kind == "chat_stream":
        for s in streams:
            _check(s["chunk"], GOLD["job:stream(chat)"]["chunk_keys"])
        assert done["result"]["message"] == {"content": "x"}
    elif kind == "generate":
        _check(done["result"], GOLD["InferenceResponse(generate)"]["keys"], allow_extra=True)
    elif kind == "embedding":
        _check(done["result"], GOLD["InferenceResponse(embedding)"]["keys"])


def test_failure_and_unknown_model_publish_job_failed():
    from gridllm_b200.worker import LocalBus, NativeWorker
    for svc, kw, msg in ((FakeService(fail=True), {}, "Inference failed: boom"), (FakeService(), {"model": "nope"}, "Model nope is not available")):
        bus = LocalBus()
        w = NativeWorker("w0", svc, bus)
        _run(w.start())
        _run(bus.publish("worker:w0:job", _assignment("j2", **kw)))
        failed = json.loads([m for c, m in bus.log if c == "job:failed"][0])
        _check(failed, GOLD["job:failed"]["keys"])
        assert failed["error"] == msg
        assert json.loads([m for c, m in bus.log if c == "job:result:j2"][0]) == failed
        assert w.isProcessingJob is False and w.currentJobs == 0


def test_busy_worker_drops_second_assignment():
    """WorkerClientService.ts:500-505: an assignment received while busy is silently dropped."""
    from gridllm_b200.worker import LocalBus, NativeWorker
    bus = LocalBus()
    svc = FakeService()
    w = NativeWorker("w0", svc, bus)
    _run(w.start())
    w.isProcessingJob = True
    n = len(bus.log)
    _run(bus.publish("worker:w0:job", _assignment("j3")))
    assert svc.calls == [] and len(bus.log) == n + 1          # only the assignment itself was published
