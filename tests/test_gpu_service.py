"""GPU: the reference-facing host layer over the real engine -- NativeInferenceService (drop-in for
OllamaService) and NativeWorker (WorkerClientService job dispatch), driven by the scheduler stand-in."""
import asyncio
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


@pytest.fixture(scope="module")
def svc(tiny_gguf):
    from gridllm_b200.service import NativeInferenceService
    s = NativeInferenceService({"tiny:latest": tiny_gguf}, device=0)
    yield s
    s.close()


def test_health_models_validate(svc):
    assert _run(svc.checkHealth()) is True
    models = _run(svc.getAvailableModels())
    assert models[0]["name"] == "tiny:latest" and models[0]["details"]["format"] == "gguf" and models[0]["size"] > 0
    assert _run(svc.validateModel("tiny:latest")) and not _run(svc.validateModel("other"))


def test_generate_text_prompt_roundtrip(svc):
    from gridllm_b200 import native as N
    req = {"id": "r1", "model": "tiny:latest", "prompt": "hello world, the rain in spain", "options": {"num_predict": 6, "temperature": 0, "ignore_eos": True},
           "priority": "medium"}
    res = _run(svc.generateResponse(req))
    assert res["id"] == "r1" and res["done"] is True and res["done_reason"] == "length"
    assert res["eval_count"] == 6 and res["prompt_eval_count"] > 3 and res["eval_duration"] > 0 and res["prompt_eval_duration"] > 0
    eng = svc._engine("tiny:latest")
    ids = eng.tokenize(req["prompt"])
    assert eng.detokenize(ids[1:]) == req["prompt"]                 # byte-level BPE round trip (id 0 is BOS)
    g = eng.generate(ids, num_predict=6, ignore_eos=True)
    assert res["token_ids"] == [int(t) for t in g.ids]                # service == direct C-ABI call
    assert res["response"] == eng.detokenize(g.ids)
    # default generation length when num_predict is absent: OllamaService.ts:105 -> 128
    res2 = _run(svc.generateResponse({"id": "r2", "model": "tiny:latest", "prompt": "hi", "options": {"ignore_eos": True}, "priority": "low"}))
    assert res2["eval_count"] == 128


def test_stream_equals_non_stream(svc):
    req = {"id": "s1", "model": "tiny:latest", "prompt": "in the end", "stream": True, "options": {"num_predict": 9, "ignore_eos": True}, "priority": "medium"}

    async def collect():
        out = []
        async for c in svc.generateStreamResponse(req):
            out.append(c)
        return out
    chunks = _run(collect())
    assert [c["done"] for c in chunks] == [False] * 9 + [True]
    assert all(set(c) == {"id", "response", "done"} for c in chunks)
    full = _run(svc.generateResponse(dict(req, stream=False)))
    assert "".join(c["response"] for c in chunks) == full["response"]


def test_chat_and_embedding(svc):
    chat = {"id": "c1", "model": "tiny:latest", "options": {"num_predict": 4, "ignore_eos": True}, "priority": "high",
            "metadata": {"requestType": "chat", "messages": [{"role": "user", "content": "hello"}]}}
    res = _run(svc.generateChatResponse(chat))
    assert res["message"]["role"] == "assistant" and res["eval_count"] == 4 and "response" not in res
    emb = _run(svc.generateEmbedding({"id": "e1", "model": "tiny:latest", "input": ["the rain", "hello world"], "priority": "low",
                                      "metadata": {"requestType": "embedding"}}))
    e = np.asarray(emb["embeddings"])
    assert e.shape == (2, 256) and np.allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-5)
    with pytest.raises(RuntimeError, match="Embedding failed: Input is required"):
        _run(svc.generateEmbedding({"id": "e2", "model": "tiny:latest", "priority": "low"}))
    with pytest.raises(RuntimeError, match="Chat inference failed: Chat request must include messages"):
        _run(svc.generateChatResponse({"id": "c2", "model": "tiny:latest", "priority": "low"}))
    with pytest.raises(RuntimeError, match="Inference failed"):
        _run(svc.generateResponse({"id": "x", "model": "missing", "prompt": "a", "priority": "low"}))


def test_workers_shard_requests_like_the_scheduler(tiny_gguf):
    """Two in-process workers (two engines on this GPU; one per GPU on a multi-GPU box), sharded by the
    scheduler stand-in: least-loaded selection spreads jobs, priority is honoured, and every result equals
    the single-engine answer (replicas are independent: no collective on this path)."""
    from gridllm_b200.service import NativeInferenceService
    from gridllm_b200.worker import LocalBus, NativeWorker
    from sched_standin import SchedulerStandIn
    from gridllm_b200 import native as N
    ndev = N.device_count()
    bus = LocalBus()
    sched = SchedulerStandIn(bus)
    svcs = [NativeInferenceService({"tiny:latest": tiny_gguf}, device=i % ndev) for i in range(2)]
    workers = [NativeWorker(f"b200-{i}", s, bus) for i, s in enumerate(svcs)]

    async def go():
        await sched.start()
        for w in workers:
            await w.start()
        for i in range(6):
            ids = np.random.Generator(np.random.PCG64(1000 + i)).integers(0, 500, size=20).tolist()
            sched.add_job({"id": f"job-{i}", "model": "tiny:latest", "prompt": "", "stream": False, "priority": "high" if i == 5 else "medium",
                           "options": {"num_predict": 5, "ignore_eos": True}, "timeout": 300000, "metadata": {"prompt_token_ids": ids}})
        await sched.run_until_empty()
    _run(go())
    assert len(sched.results) == 6 and all("result" in r for r in sched.results.values())
    used = set(sched.assigned.values())
    assert used == {"b200-0", "b200-1"}
    first_two = [c for c, m in bus.log if c.startswith("worker:b200-") and c.endswith(":job")][:2]
    assigned_first = json.loads([m for c, m in bus.log if c == first_two[0]][0])["job"]["jobId"]
    assert assigned_first == "job-5"                                   # priority sort (JobScheduler.ts:145-151)
    ref = N.Engine(tiny_gguf)
    for i in range(6):
        ids = np.random.Generator(np.random.PCG64(1000 + i)).integers(0, 500, size=20)
        g = ref.generate(ids, num_predict=5, ignore_eos=True)
        assert sched.results[f"job-{i}"]["result"]["token_ids"] == [int(t) for t in g.ids]
    ref.close()
    for s in svcs:
        s.close()


def test_sampled_requests(svc):
    """options.temperature / top_k / top_p / seed reach the native sampler: same seed -> same text, streamed or not;
    another seed -> another draw; a sampled request without a seed still answers."""
    opts = {"num_predict": 16, "ignore_eos": True, "temperature": 0.9, "top_k": 40, "top_p": 0.95, "seed": 7}
    req = {"id": "t1", "model": "tiny:latest", "prompt": "once upon a time", "options": opts, "priority": "medium"}
    a = _run(svc.generateResponse(req))
    b = _run(svc.generateResponse(dict(req, id="t2")))
    c = _run(svc.generateResponse(dict(req, id="t3", options=dict(opts, seed=8))))
    assert a["context"] == b["context"] and a["response"] == b["response"] and a["eval_count"] == 16
    assert a["context"] != c["context"]
    greedy = _run(svc.generateResponse(dict(req, id="t4", options=dict(opts, temperature=0))))
    assert greedy["context"] != a["context"]

    async def collect():
        return [ch async for ch in svc.generateStreamResponse(dict(req, id="t5", stream=True))]
    chunks = _run(collect())
    assert "".join(ch["response"] for ch in chunks) == a["response"]
    d = _run(svc.generateResponse(dict(req, id="t6", options={k: v for k, v in opts.items() if k != "seed"})))
    assert d["eval_count"] == 16
    with pytest.raises(RuntimeError, match="Inference failed"):
        _run(svc.generateResponse(dict(req, id="t7", options=dict(opts, temperature=-1))))


# ---- the north_star's multi-GPU shape: N engines, one per GPU, in ONE process behind the scheduler rules ------
def test_in_process_workers_one_engine_per_gpu(tiny128_gguf):
    """One NativeWorker + NativeInferenceService + engine per visible GPU (up to 8), all in this process, each registered under its
    own worker id; the restated scheduler rules (tests/sched_standin.py: JobScheduler.ts:137-217, 317-360) shard 6 jobs per GPU.
    Every worker id is used, priority order is honoured, and every result equals the single-engine answer for that prompt.
    Needs >= 2 GPUs (`gpurun --gpus N`, the driver's multi-GPU tier); on one GPU the same path runs in
    test_workers_shard_requests_like_the_scheduler with both engines on device 0."""
    from gridllm_b200 import native as N
    from gridllm_b200.service import NativeInferenceService
    from gridllm_b200.worker import LocalBus, NativeWorker
    from sched_standin import SchedulerStandIn
    n_dev = min(8, N.device_count())
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    bus = LocalBus()
    sched = SchedulerStandIn(bus, max_jobs_per_worker=2)
    svcs = [NativeInferenceService({"tiny128:latest": tiny128_gguf}, device=d, max_batch=2) for d in range(n_dev)]
    workers = [NativeWorker(f"b200-{d}", s, bus, max_concurrent=2) for d, s in enumerate(svcs)]
    n_jobs = 6 * n_dev
    prompts = [np.random.Generator(np.random.PCG64(7000 + i)).integers(0, 1000, size=20 + i % 7).tolist() for i in range(n_jobs)]

    async def go():
        await sched.start()
        for w in workers:
            await w.start()
        for i, p in enumerate(prompts):
            sched.add_job({"id": f"job-{i}", "model": "tiny128:latest", "prompt": "", "stream": i % 3 == 0, "timeout": 300000,
                           "priority": "high" if i == n_jobs - 1 else "medium", "options": {"num_predict": 6, "ignore_eos": True},
                           "metadata": {"prompt_token_ids": p}})
            if i % 3 == 0:
                await sched.watch_stream(f"job-{i}")
        await sched.run_until_empty()
        for w in workers:
            await w.stop()
    asyncio.new_event_loop().run_until_complete(go())
    assert len(sched.results) == n_jobs and all("result" in r for r in sched.results.values())
    assert set(sched.assigned.values()) == {f"b200-{d}" for d in range(n_dev)}                  # every GPU's worker took jobs
    first_assignment = [json.loads(m) for c, m in bus.log if c.endswith(":job")][0]
    assert first_assignment["job"]["jobId"] == f"job-{n_jobs - 1}"                              # the high-priority job went out first
    ref_engine = N.Engine(tiny128_gguf, device=0, max_batch=2)                                  # the single-engine answers, same batched path
    for i, p in enumerate(prompts):
        slot = ref_engine.seq_open(p, num_predict=6, ignore_eos=True)
        ids = []
        while len(ids) < 6:
            ids += [t for s, t, _lp, _d in ref_engine.batch_step() if s == slot]
        ref_engine.seq_close(slot)
        r = sched.results[f"job-{i}"]["result"]
        if i % 3 == 0:
            assert sched.stream_chunks[f"job-{i}"] == 7 and r["done"] is True
        else:
            assert r["token_ids"] == ids, i
    ref_engine.close()
    for s in svcs:
        s.close()
