"""CPU: the reference-facing host layer end to end -- NativeInferenceService (drop-in for OllamaService) and NativeWorker
(WorkerClientService job dispatch) exactly as shipped, over an engine double backed by the numpy oracle and the product
tokenizer (tests/oracle_engine.py).  The same scenarios run against the real engine in tests/test_gpu_service.py."""
import asyncio
import json

import numpy as np
import pytest


def _run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


@pytest.fixture()
def svc(tiny_gguf, hostcheck_lib, monkeypatch):
    import oracle_engine
    from gridllm_b200 import service as SV
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 1)
    s = SV.NativeInferenceService({"tiny:latest": tiny_gguf}, device=0)
    yield s
    s.close()


def test_health_models_validate(svc):
    assert _run(svc.checkHealth()) is True
    models = _run(svc.getAvailableModels())
    assert models[0]["name"] == "tiny:latest" and models[0]["details"]["format"] == "gguf" and models[0]["size"] > 0
    assert _run(svc.validateModel("tiny:latest")) and not _run(svc.validateModel("other"))


def test_generate_matches_the_oracle_and_streams_the_same_text(svc):
    from oracle import llama_oracle as O
    req = {"id": "r1", "model": "tiny:latest", "prompt": "hello world, the rain in spain", "options": {"num_predict": 6, "temperature": 0, "ignore_eos": True},
           "priority": "medium"}
    res = _run(svc.generateResponse(req))
    assert res["id"] == "r1" and res["done"] is True and res["done_reason"] == "length" and res["eval_count"] == 6
    eng = svc._engine("tiny:latest")
    ids = eng.tokenize(req["prompt"])
    assert res["prompt_eval_count"] == len(ids) and eng.detokenize(ids[1:]) == req["prompt"]
    ref = O.LlamaOracle(eng.m, act="i16", kv_f16=True).generate(ids, 6)
    assert res["token_ids"] == [int(t) for t in ref["ids"]]                  # the service hands the oracle's greedy ids through
    assert res["response"] == eng.detokenize(ref["ids"])

    async def collect():
        return [c async for c in svc.generateStreamResponse(dict(req, id="s1", stream=True))]
    chunks = _run(collect())
    assert [c["done"] for c in chunks] == [False] * 6 + [True] and all(set(c) == {"id", "response", "done"} for c in chunks)
    assert "".join(c["response"] for c in chunks) == res["response"]
    # default generation length when num_predict is absent: OllamaService.ts:105 -> 128
    res2 = _run(svc.generateResponse({"id": "r2", "model": "tiny:latest", "prompt": "hi", "options": {"ignore_eos": True}, "priority": "low"}))
    assert res2["eval_count"] == 128


def test_sampled_and_stop_options_reach_the_engine(svc):
    opts = {"num_predict": 12, "ignore_eos": True, "temperature": 0.9, "top_k": 40, "top_p": 0.95, "seed": 7}
    req = {"id": "t1", "model": "tiny:latest", "prompt": "once upon a time", "options": opts, "priority": "medium"}
    a = _run(svc.generateResponse(req))
    b = _run(svc.generateResponse(dict(req, id="t2")))
    c = _run(svc.generateResponse(dict(req, id="t3", options=dict(opts, seed=8))))
    assert a["context"] == b["context"] and a["context"] != c["context"]
    call = svc._engine("tiny:latest").calls[-1]
    assert (call["temperature"], call["top_k"], call["top_p"], call["seed"]) == (0.9, 40, 0.95, 8)
    # a stop string taken from the middle of the sampled text ends the response there, streamed or not
    text = a["response"]
    assert len(text) >= 4
    stop = text[len(text) // 2: len(text) // 2 + 2]
    cut = text.index(stop)
    d = _run(svc.generateResponse(dict(req, id="t4", options=dict(opts, stop=[stop]))))
    assert d["response"] == text[:cut] and d["done_reason"] == "stop" and d["eval_count"] <= a["eval_count"]

    async def collect():
        return [ch async for ch in svc.generateStreamResponse(dict(req, id="t5", stream=True, options=dict(opts, stop=[stop])))]
    assert "".join(ch["response"] for ch in _run(collect())) == text[:cut]


def test_chat_embedding_and_errors(svc):
    chat = {"id": "c1", "model": "tiny:latest", "options": {"num_predict": 4, "ignore_eos": True}, "priority": "high",
            "metadata": {"requestType": "chat", "messages": [{"role": "user", "content": "hello"}]}}
    res = _run(svc.generateChatResponse(chat))
    assert res["message"]["role"] == "assistant" and res["eval_count"] == 4 and "response" not in res

    async def collect():
        return [c async for c in svc.generateChatStreamResponse(dict(chat, id="c3", stream=True))]
    assert "".join(c["response"] for c in _run(collect())) == res["message"]["content"]
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


def test_workers_shard_requests_like_the_scheduler(tiny_gguf, hostcheck_lib, monkeypatch):
    """two workers behind the scheduler stand-in: least-loaded selection, priority order, results equal to the oracle's"""
    import oracle_engine
    from gridllm_b200 import service as SV
    from gridllm_b200.worker import LocalBus, NativeWorker
    from oracle import llama_oracle as O
    from sched_standin import SchedulerStandIn
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 2)
    bus = LocalBus()
    sched = SchedulerStandIn(bus)
    svcs = [SV.NativeInferenceService({"tiny:latest": tiny_gguf}, device=i) for i in range(2)]
    workers = [NativeWorker(f"b200-{i}", s, bus) for i, s in enumerate(svcs)]

    async def go():
        await sched.start()
        for w in workers:
            await w.start()
        for i in range(6):
            ids = np.random.Generator(np.random.PCG64(1000 + i)).integers(0, 500, size=12).tolist()
            sched.add_job({"id": f"job-{i}", "model": "tiny:latest", "prompt": "", "stream": False, "priority": "high" if i == 5 else "medium",
                           "options": {"num_predict": 4, "ignore_eos": True}, "timeout": 300000, "metadata": {"prompt_token_ids": ids}})
        await sched.run_until_empty()
    _run(go())
    assert len(sched.results) == 6 and all("result" in r for r in sched.results.values())
    assert set(sched.assigned.values()) == {"b200-0", "b200-1"}
    first = [c for c, m in bus.log if c.startswith("worker:b200-") and c.endswith(":job")][0]
    assert json.loads([m for c, m in bus.log if c == first][0])["job"]["jobId"] == "job-5"          # priority sort (JobScheduler.ts:145-151)
    m = O.load_gguf(tiny_gguf)
    for i in range(6):
        ids = np.random.Generator(np.random.PCG64(1000 + i)).integers(0, 500, size=12)
        ref = O.LlamaOracle(m, act="i16", kv_f16=True).generate(ids, 4)
        assert sched.results[f"job-{i}"]["result"]["token_ids"] == [int(t) for t in ref["ids"]]


def test_chat_framing_follows_the_template_family(svc, monkeypatch):
    monkeypatch.setattr(svc, "_jinja", False)               # the fallback path: template FAMILY from its markers
    eng = svc._engine("tiny:latest")
    msgs = [{"role": "system", "content": "be brief"}, {"role": "user", "content": "hi"}, {"role": "assistant", "content": "hello"},
            {"role": "user", "content": "bye"}]
    eng.chat_template = ""                                   # no template in the file: Llama-3 framing
    p = svc._chat_prompt(eng, msgs)
    assert p.startswith("<|start_header_id|>system<|end_header_id|>\n\nbe brief<|eot_id|>") and p.endswith("<|start_header_id|>assistant<|end_header_id|>\n\n")
    ids = eng.tokenize(p, add_bos=True, parse_special=True)
    assert int(ids[0]) == eng.info.bos_id and (ids == eng.info.eot_id).sum() == 4      # control tokens parsed, not spelled out
    eng.chat_template = "{% for m in messages %}<|im_start|>{{ m.role }}\n{{ m.content }}<|im_end|>\n{% endfor %}"
    assert svc._chat_prompt(eng, msgs) == ("<|im_start|>system\nbe brief<|im_end|>\n<|im_start|>user\nhi<|im_end|>\n<|im_start|>assistant\nhello<|im_end|>\n"
                                           "<|im_start|>user\nbye<|im_end|>\n<|im_start|>assistant\n")
    eng.chat_template = "{{ bos_token }}{% for m in messages %}[INST] {{ m.content }} [/INST]{% endfor %}"
    assert svc._chat_prompt(eng, msgs) == "[INST] be brief\n\nhi [/INST] hello</s>[INST] bye [/INST]"
    eng.chat_template = "[INST] <<SYS>>\n{{ system }}\n<</SYS>>\n\n{{ user }} [/INST]"
    assert svc._chat_prompt(eng, msgs[:2]) == "[INST] <<SYS>>\nbe brief\n<</SYS>>\n\nhi [/INST]"

    class Broken:                                            # an engine whose accessor fails still gets the default framing
        @property
        def chat_template(self):
            raise RuntimeError("boom")
    assert svc._chat_prompt(Broken(), msgs[1:2]).startswith("<|start_header_id|>user<|end_header_id|>\n\nhi<|eot_id|>")


LLAMA3_TEMPLATE = ("{% set loop_messages = messages %}{% for message in loop_messages %}{% set content = '<|start_header_id|>' + message['role'] + "
                   "'<|end_header_id|>\n\n'+ message['content'] | trim + '<|eot_id|>' %}{% if loop.index0 == 0 %}{% set content = bos_token + content %}"
                   "{% endif %}{{ content }}{% endfor %}{% if add_generation_prompt %}{{ '<|start_header_id|>assistant<|end_header_id|>\n\n' }}{% endif %}")
CHATML_TEMPLATE = ("{% for message in messages %}{{'<|im_start|>' + message['role'] + '\n' + message['content'] + '<|im_end|>' + '\n'}}{% endfor %}"
                   "{% if add_generation_prompt %}{{ '<|im_start|>assistant\n' }}{% endif %}")
MISTRAL_TEMPLATE = ("{{ bos_token }}{% for message in messages %}{% if (message['role'] == 'user') != (loop.index0 % 2 == 0) %}"
                    "{{ raise_exception('Conversation roles must alternate user/assistant/user/assistant/...') }}{% endif %}"
                    "{% if message['role'] == 'user' %}{{ '[INST] ' + message['content'] + ' [/INST]' }}{% elif message['role'] == 'assistant' %}"
                    "{{ message['content'] + eos_token}}{% else %}{{ raise_exception('Only user and assistant roles are supported!') }}{% endif %}{% endfor %}")


def test_chat_template_is_rendered_as_jinja(svc):
    """tokenizer.chat_template is a Jinja program: the host renders it (messages, add_generation_prompt, bos_token / eos_token from
    the GGUF's own vocabulary, raise_exception) as Ollama's runner and HF apply_chat_template do [external]; the published Llama-3,
    ChatML and Mistral templates give their documented framings, and a template that refuses the conversation falls back to its family."""
    pytest.importorskip("jinja2")
    eng = svc._engine("tiny:latest")
    bos, eos = svc._control_text(eng, eng.info.bos_id), svc._control_text(eng, eng.info.eos_id)
    assert bos and eos                                       # control tokens have a spelling even though they stream as nothing
    msgs = [{"role": "system", "content": "be brief"}, {"role": "user", "content": "  hi  "}, {"role": "assistant", "content": "hello"},
            {"role": "user", "content": "bye"}]
    eng.chat_template = LLAMA3_TEMPLATE
    p = svc._chat_prompt(eng, msgs)
    # content is trimmed by the template itself ("| trim"); the leading BOS text is left to the tokenizer
    assert p == ("<|start_header_id|>system<|end_header_id|>\n\nbe brief<|eot_id|><|start_header_id|>user<|end_header_id|>\n\nhi<|eot_id|>"
                 "<|start_header_id|>assistant<|end_header_id|>\n\nhello<|eot_id|><|start_header_id|>user<|end_header_id|>\n\nbye<|eot_id|>"
                 "<|start_header_id|>assistant<|end_header_id|>\n\n")
    assert not p.startswith(bos)
    ids = eng.tokenize(p, add_bos=True, parse_special=True)
    assert int(ids[0]) == eng.info.bos_id and int(ids[1]) != eng.info.bos_id
    eng.chat_template = CHATML_TEMPLATE
    assert svc._chat_prompt(eng, msgs) == ("<|im_start|>system\nbe brief<|im_end|>\n<|im_start|>user\n  hi  <|im_end|>\n<|im_start|>assistant\nhello<|im_end|>\n"
                                           "<|im_start|>user\nbye<|im_end|>\n<|im_start|>assistant\n")
    eng.chat_template = MISTRAL_TEMPLATE
    assert svc._chat_prompt(eng, msgs[1:]) == "[INST]   hi   [/INST]hello" + eos + "[INST] bye [/INST]"
    # the Mistral template refuses a system role (raise_exception): the family framing folds it into the first user turn
    assert svc._chat_prompt(eng, msgs) == "[INST] be brief\n\n  hi   [/INST] hello</s>[INST] bye [/INST]"
    # family only, on request
    svc._jinja = False
    eng.chat_template = CHATML_TEMPLATE.replace("{% if add_generation_prompt %}", "{% if false %}")
    assert svc._chat_prompt(eng, msgs[:2]).endswith("<|im_start|>assistant\n")
    svc._jinja = True


def test_mistral_family_text_goes_through_sentencepiece_and_its_template(tmp_path, hostcheck_lib, monkeypatch):
    """BASELINE config 5's model family end to end on the host side: a GGUF whose tokenizer is SentencePiece BPE
    (tokenizer.ggml.model = "llama") -- document TEXT reaches the engine as the ids the `sentencepiece` library gives, the
    embeddings come back one per document, and a chat request is framed by the published Mistral template ([INST] ... [/INST],
    `eos_token` spelled as the vocabulary spells it) before it is tokenised."""
    spm = pytest.importorskip("sentencepiece")
    pytest.importorskip("jinja2")
    import io
    import random
    import oracle_engine
    from gridllm_b200 import service as SV
    from oracle import gguf_synth as S
    rnd = random.Random(3)
    words = ["the", "quick", "brown", "fox", "hello", "world", "embedding", "worker", "scheduler", "document", "query", "gpu", "naïve", "café"]
    lines = [" ".join(rnd.choice(words) for _ in range(rnd.randint(3, 10))) for _ in range(2000)]
    model = io.BytesIO()
    spm.SentencePieceTrainer.train(sentence_iterator=iter(lines), model_writer=model, vocab_size=400, model_type="bpe", byte_fallback=True,
                                   bos_id=1, eos_id=2, unk_id=0, pad_id=-1, normalization_rule_name="identity", add_dummy_prefix=True,
                                   remove_extra_whitespaces=False, minloglevel=2, user_defined_symbols=["[INST]", "[/INST]"])
    sp = spm.SentencePieceProcessor(model_proto=model.getvalue())
    n = sp.get_piece_size()
    toks = [sp.id_to_piece(i) for i in range(n)]
    types = [2 if sp.is_unknown(i) else 3 if sp.is_control(i) else 6 if sp.is_byte(i) else 4 if toks[i] in ("[INST]", "[/INST]") else 1 for i in range(n)]
    shape = S.LlamaShape("tiny-mistral-synth", 2, 256, 4, 2, 512, n, 10000.0, 1e-5, 512)
    path = str(tmp_path / "tiny_mistral.gguf")
    S.build_model(path, shape, "q8_0", seed=11, spm_vocab={"tokens": toks, "scores": [sp.get_score(i) for i in range(n)], "types": types,
                                                            "bos": 1, "eos": 2, "unk": 0, "chat_template": MISTRAL_TEMPLATE})
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 1)
    s = SV.NativeInferenceService({"mistral:tiny": path})
    eng = s._engine("mistral:tiny")
    assert eng.info.bos_id == 1 and eng.info.eos_id == 2
    docs = ["the quick brown fox", "hello world, naïve café", "an unseen wörd ✓"]
    for d in docs:
        assert list(eng.tokenize(d, add_bos=True)) == [1] + sp.encode(d)
    res = _run(s.generateEmbedding({"id": "e1", "model": "mistral:tiny", "input": docs, "metadata": {"requestType": "embedding"}}))
    emb = np.asarray(res["embeddings"])
    assert emb.shape == (3, 256) and np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-4) and res["prompt_eval_count"] == sum(1 + len(sp.encode(d)) for d in docs)
    # chat: the template is rendered (eos_token = "</s>" from the vocabulary), then tokenised with the markers as single pieces
    eng.chat_template = MISTRAL_TEMPLATE
    msgs = [{"role": "user", "content": "hello world"}, {"role": "assistant", "content": "the fox"}, {"role": "user", "content": "quick query"}]
    prompt = s._chat_prompt(eng, msgs)
    assert prompt == "[INST] hello world [/INST]the fox</s>[INST] quick query [/INST]"
    ids = list(eng.tokenize(prompt, add_bos=True, parse_special=True))
    inst, inst_end = sp.piece_to_id("[INST]"), sp.piece_to_id("[/INST]")
    assert ids[0] == 1 and ids.count(inst) == 2 and ids.count(inst_end) == 2 and ids.count(2) == 1
    s.close()


def test_generate_prompts_can_be_framed_like_ollama(tiny_gguf, hostcheck_lib, monkeypatch):
    """apply_template=True: /api/generate-style prompts become one user turn of the model's template (metadata.system as the
    system turn) unless metadata.raw; the default service tokenises the prompt text as it is"""
    import oracle_engine
    from gridllm_b200 import service as SV
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    s = SV.NativeInferenceService({"tiny:latest": tiny_gguf}, apply_template=True)
    eng = s._engine("tiny:latest")
    req = {"id": "g", "model": "tiny:latest", "prompt": "hi there", "metadata": {"system": "be brief"}}
    framed = eng.detokenize(s._prompt_ids(eng, req, req["prompt"])[1:])
    assert "be brief" in framed and "hi there" in framed and framed.index("be brief") < framed.index("hi there")
    ids = s._prompt_ids(eng, req, req["prompt"])
    assert (ids == eng.info.eot_id).sum() == 2
    raw = s._prompt_ids(eng, dict(req, metadata={"raw": True}), req["prompt"])
    assert list(raw) == list(eng.tokenize("hi there"))
    plain = SV.NativeInferenceService({"tiny:latest": tiny_gguf})
    assert list(plain._prompt_ids(plain._engine("tiny:latest"), req, req["prompt"])) == list(eng.tokenize("hi there"))


def test_embedding_inputs_are_truncated_to_the_context(tiny_gguf, hostcheck_lib, monkeypatch):
    import oracle_engine
    from gridllm_b200 import service as SV
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    s = SV.NativeInferenceService({"tiny:latest": tiny_gguf}, max_ctx=16)
    long_ids = list(range(1, 41))
    req = {"id": "e", "model": "tiny:latest", "metadata": {"requestType": "embedding", "input_token_ids": [long_ids, long_ids[:5]]}}
    res = _run(s.generateEmbedding(req))
    assert res["prompt_eval_count"] == 16 + 5 and len(res["embeddings"]) == 2
    ref = _run(s.generateEmbedding(dict(req, metadata=dict(req["metadata"], input_token_ids=[long_ids[:16], long_ids[:5]]))))
    assert np.allclose(res["embeddings"], ref["embeddings"])
    with pytest.raises(RuntimeError, match="Embedding failed"):
        _run(s.generateEmbedding(dict(req, metadata=dict(req["metadata"], truncate=False))))


def test_generation_ends_at_the_context(tiny_gguf, hostcheck_lib, monkeypatch):
    import oracle_engine
    from gridllm_b200 import service as SV
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    s = SV.NativeInferenceService({"tiny:latest": tiny_gguf}, max_ctx=16)
    req = {"id": "g", "model": "tiny:latest", "prompt": "", "options": {"num_predict": 50, "ignore_eos": True},
           "metadata": {"prompt_token_ids": list(range(1, 11))}}
    res = _run(s.generateResponse(req))
    assert res["prompt_eval_count"] == 10 and res["eval_count"] == 6 and res["done_reason"] == "length"
    assert s._engine("tiny:latest").calls[-1]["num_predict"] == 6


def test_heartbeats_keep_firing_while_a_job_runs(tiny_gguf, hostcheck_lib, monkeypatch):
    """the native call must not block the worker's event loop (heartbeats, WorkerClientService.ts:316-323)"""
    import oracle_engine
    from gridllm_b200 import service as SV
    from gridllm_b200.worker import LocalBus, NativeWorker
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 1)
    bus = LocalBus()
    w = NativeWorker("b200-0", SV.NativeInferenceService({"tiny:latest": tiny_gguf}), bus, heartbeat_interval_ms=10)
    job = {"type": "job_assignment", "job": {"jobId": "j1", "request": {
        "id": "j1", "model": "tiny:latest", "prompt": "", "stream": False, "priority": "medium", "options": {"num_predict": 60, "ignore_eos": True},
        "metadata": {"prompt_token_ids": list(range(1, 40))}}}}

    async def go():
        await w.start()
        w.start_heartbeats()
        await bus.publish("worker:b200-0:job", json.dumps(job))
        await asyncio.sleep(0.03)
        await w.stop()
    _run(go())
    beats = [json.loads(m) for c, m in bus.log if c == "worker:heartbeat"]
    busy = [b for b in beats if b["status"] == "busy"]
    assert len(busy) >= 3 and all(b["currentJobs"] == 1 for b in busy)      # fired DURING the job
    assert beats[-1]["status"] == "online"                                   # and after it
    assert any(c == "job:completed" for c, _ in bus.log)


def test_job_cancellation_stops_a_streaming_job(tiny_gguf, hostcheck_lib, monkeypatch):
    """job_cancellation (JobScheduler.ts:530-536; ignored by the reference worker, honoured here): the stream ends early, the
    engine call is cancelled through the token callback, the result carries what was generated so far"""
    import oracle_engine
    from gridllm_b200 import service as SV
    from gridllm_b200.worker import LocalBus, NativeWorker
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 1)
    bus = LocalBus()
    svc = SV.NativeInferenceService({"tiny:latest": tiny_gguf})
    w = NativeWorker("b200-0", svc, bus)
    job = {"type": "job_assignment", "job": {"jobId": "j1", "request": {
        "id": "j1", "model": "tiny:latest", "prompt": "", "stream": True, "priority": "medium", "options": {"num_predict": 120, "ignore_eos": True},
        "metadata": {"prompt_token_ids": list(range(1, 10))}}}}

    async def go():
        await w.start()
        task = asyncio.get_running_loop().create_task(bus.publish("worker:b200-0:job", json.dumps(job)))
        while not any(c == "job:stream:j1" for c, _ in bus.log):        # wait for the first streamed chunk
            await asyncio.sleep(0.005)
        await w.handleJobMessage(json.dumps({"type": "job_cancellation", "jobId": "j1"}))
        await task
    _run(go())
    chunks = [json.loads(m) for c, m in bus.log if c == "job:stream:j1"]
    assert 1 <= len(chunks) < 120
    done = [json.loads(m) for c, m in bus.log if c == "job:completed"]
    assert len(done) == 1 and done[0]["result"]["done"] is True
    # the engine stopped too: far fewer than the 120 requested steps were taken
    import time
    time.sleep(0.05)
    assert w.isProcessingJob is False and w.currentJobs == 0


def test_worker_can_hold_more_than_one_job(tiny_gguf, hostcheck_lib, monkeypatch):
    """max_concurrent = 2 (MAX_CONCURRENT_JOBS_PER_WORKER > 1 on the server side): a second assignment is accepted while the
    first runs and waits at the engine; a third is dropped like the reference drops a second; results are unaffected"""
    import oracle_engine
    from gridllm_b200 import service as SV
    from gridllm_b200.worker import LocalBus, NativeWorker
    from oracle import llama_oracle as O
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 1)
    bus = LocalBus()
    w = NativeWorker("b200-0", SV.NativeInferenceService({"tiny:latest": tiny_gguf}), bus, heartbeat_interval_ms=5, max_concurrent=2)
    prompts = {f"j{i}": np.random.Generator(np.random.PCG64(50 + i)).integers(0, 500, size=10).tolist() for i in range(3)}

    def job(jid):
        return json.dumps({"type": "job_assignment", "job": {"jobId": jid, "request": {
            "id": jid, "model": "tiny:latest", "prompt": "", "stream": False, "priority": "medium", "options": {"num_predict": 20, "ignore_eos": True},
            "metadata": {"prompt_token_ids": prompts[jid]}}}})

    async def go():
        await w.start()
        assert w.capabilities["maxConcurrentTasks"] == 2
        w.start_heartbeats()
        for jid in ("j0", "j1", "j2"):
            await bus.publish("worker:b200-0:job", job(jid))         # returns at once: the jobs run as tasks
        await asyncio.sleep(0.01)                                    # let the tasks start
        assert w.currentJobs == 2                                    # j2 was dropped
        await w.stop()
    _run(go())
    done = {json.loads(m)["jobId"]: json.loads(m) for c, m in bus.log if c == "job:completed"}
    assert set(done) == {"j0", "j1"} and not any(c == "job:failed" for c, _ in bus.log)
    m = O.load_gguf(tiny_gguf)
    for jid in ("j0", "j1"):
        ref = O.LlamaOracle(m, act="i16", kv_f16=True).generate(prompts[jid], 20)
        assert done[jid]["result"]["token_ids"] == [int(t) for t in ref["ids"]]
    assert max(json.loads(m)["currentJobs"] for c, m in bus.log if c == "worker:heartbeat") == 2
    assert w.currentJobs == 0 and w.isProcessingJob is False


def test_context_round_trip_continues_the_whole_conversation(svc):
    """Ollama's `context` is the conversation so far; the gateway returns it (ollama.ts:143) and forwards it back
    (ollama.ts:234 -> metadata.context): the second turn must run on prompt1 + reply1 + prompt2, BOS included once."""
    eng = svc._engine("tiny:latest")
    r1 = _run(svc.generateResponse({"id": "c1", "model": "tiny:latest", "prompt": "hello world", "options": {"num_predict": 4, "ignore_eos": True},
                                    "priority": "medium"}))
    p1 = [int(t) for t in eng.tokenize("hello world")]
    assert r1["context"] == p1 + r1["token_ids"] and len(r1["token_ids"]) == 4
    n_calls = len(eng.calls)
    r2 = _run(svc.generateResponse({"id": "c2", "model": "tiny:latest", "prompt": " the rain", "options": {"num_predict": 3, "ignore_eos": True},
                                    "priority": "medium", "metadata": {"context": r1["context"]}}))
    p2 = [int(t) for t in eng.tokenize(" the rain", add_bos=False, parse_special=False)]
    assert eng.calls[n_calls]["n_prompt"] == len(p1) + 4 + len(p2)
    assert r2["context"] == r1["context"] + p2 + r2["token_ids"] and r2["prompt_eval_count"] == len(r1["context"]) + len(p2)


def test_num_predict_minus_one_runs_until_the_context_is_full(svc):
    """the gateway lets num_predict -1 through (ollama.ts:47, Ollama's 'until EOS'): it must not fail, and ends at the context"""
    eng = svc._engine("tiny:latest")
    ids = eng.tokenize("hi")
    res = _run(svc.generateResponse({"id": "n1", "model": "tiny:latest", "prompt": "hi", "options": {"num_predict": -1, "ignore_eos": True},
                                     "priority": "low"}))
    assert res["eval_count"] == eng.info.n_ctx - len(ids) and res["done_reason"] == "length"


@pytest.mark.timeout(30)
def test_drain_does_not_spin_on_tasks_that_finished_a_moment_ago(svc):
    """NativeWorker.drain() entered while a job task is done but its done-callback (which removes it from the worker's task set) has
    not run yet: gather() of finished tasks completes without yielding, so the loop must prune and yield itself -- it used to spin
    for ever (bench.py --workload config3 --gpus 2, 1 s dispatch tick)."""
    from gridllm_b200.worker import LocalBus, NativeWorker
    w = NativeWorker("b200-0", svc, LocalBus(), max_concurrent=4)

    async def go():
        async def job():
            return None
        t = asyncio.get_running_loop().create_task(job())
        w._tasks.add(t)
        t.add_done_callback(w._tasks.discard)
        await asyncio.sleep(0)          # the task's only step and this coroutine's resumption run in the same loop iteration
        assert t.done() and t in w._tasks
        await w.drain()
        assert not w._tasks
    _run(go())


# ---- continuous batching on the host side (gridllm_b200/batching.py): concurrent requests share the engine's batched step ----
@pytest.fixture()
def bsvc(tiny_gguf, hostcheck_lib, monkeypatch):
    import oracle_engine
    from gridllm_b200 import service as SV
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 1)
    s = SV.NativeInferenceService({"tiny:latest": tiny_gguf}, device=0, max_batch=4)
    yield s
    s.close()


def test_concurrent_requests_are_decoded_together_and_match_the_oracle(bsvc):
    from oracle import llama_oracle as O
    eng = bsvc._engine("tiny:latest")
    prompts = [np.random.Generator(np.random.PCG64(500 + i)).integers(0, 500, size=6 + i).tolist() for i in range(6)]
    reqs = [{"id": f"b{i}", "model": "tiny:latest", "prompt": "", "options": {"num_predict": 3 + i % 3, "ignore_eos": True}, "priority": "medium",
             "metadata": {"prompt_token_ids": p}} for i, p in enumerate(prompts)]

    async def go():
        return await asyncio.gather(*[bsvc.generateResponse(r) for r in reqs])
    res = _run(go())
    for r, p, q in zip(res, prompts, reqs):
        ref = O.LlamaOracle(eng.m, act="i16", kv_f16=True).generate(np.asarray(p), q["options"]["num_predict"])
        assert r["token_ids"] == [int(t) for t in ref["ids"]] and r["eval_count"] == q["options"]["num_predict"]
        assert r["context"] == p + r["token_ids"] and r["done_reason"] == "length" and r["prompt_eval_count"] == len(p)
    assert max(eng.batch_sizes) == 4                     # six requests, four slots: the engine stepped four sequences at once,
    assert all(c.get("batched") for c in eng.calls)      # ... every request went through gl_seq_open, none through gl_generate
    runner = bsvc._runners["tiny:latest"]
    assert runner.max_rows == 4 and runner.steps >= 3
    assert max(eng.open_many_calls) >= 2                 # requests that were waiting together were admitted in one engine call


def test_batched_streams_stop_strings_and_cancellation(bsvc):
    eng = bsvc._engine("tiny:latest")
    base = {"model": "tiny:latest", "prompt": "hello world, the rain in spain", "options": {"num_predict": 8, "ignore_eos": True}, "priority": "medium"}
    plain = _run(bsvc.generateResponse(dict(base, id="p0")))
    text = plain["response"]
    assert plain["eval_count"] == 8 and len(text) >= 4
    stop = text[len(text) // 2: len(text) // 2 + 2]

    async def collect(req, cancel_after=None):
        out = []
        agen = bsvc.generateStreamResponse(req)
        async for c in agen:
            out.append(c)
            if cancel_after is not None and len(out) >= cancel_after:
                await agen.aclose()
                break
        return out

    async def go():
        return await asyncio.gather(
            collect(dict(base, id="s1", stream=True)),
            collect(dict(base, id="s2", stream=True, options=dict(base["options"], stop=[stop]))),
            collect(dict(base, id="s3", stream=True), cancel_after=2),
            bsvc.generateResponse(dict(base, id="s4")))
    full, stopped, cancelled, again = _run(go())
    assert "".join(c["response"] for c in full) == text and [c["done"] for c in full] == [False] * 8 + [True]
    assert "".join(c["response"] for c in stopped) == text[:text.index(stop)] and stopped[-1]["done"] is True
    assert len(cancelled) == 2
    assert again["response"] == text                     # a stop string or a cancellation in one sequence leaves the others alone
    # every sequence was returned to the engine: nothing is left open
    import time
    for _ in range(200):
        if not getattr(eng, "_seqs", {}):
            break
        time.sleep(0.01)
    assert not eng._seqs


def test_workers_with_concurrent_jobs_behind_the_scheduler(tiny_gguf, hostcheck_lib, monkeypatch):
    """MAX_CONCURRENT_JOBS_PER_WORKER = 3 on the server side, max_concurrent = 3 / max_batch = 3 on the worker side: the
    scheduler's least-loaded rule spreads nine jobs over two workers, each worker decodes the jobs it holds together"""
    import oracle_engine
    from gridllm_b200 import service as SV
    from gridllm_b200.worker import LocalBus, NativeWorker
    from oracle import llama_oracle as O
    from sched_standin import SchedulerStandIn
    oracle_engine.use_hostcheck(hostcheck_lib)
    monkeypatch.setattr(SV.N, "Engine", oracle_engine.OracleEngine)
    monkeypatch.setattr(SV.N, "device_count", lambda: 2)
    bus = LocalBus()
    sched = SchedulerStandIn(bus, max_jobs_per_worker=3)
    svcs = [SV.NativeInferenceService({"tiny:latest": tiny_gguf}, device=i, max_batch=3) for i in range(2)]
    workers = [NativeWorker(f"b200-{i}", s, bus, max_concurrent=3) for i, s in enumerate(svcs)]

    async def go():
        await sched.start()
        for w in workers:
            await w.start()
        for i in range(9):
            ids = np.random.Generator(np.random.PCG64(2000 + i)).integers(0, 500, size=10).tolist()
            sched.add_job({"id": f"job-{i}", "model": "tiny:latest", "prompt": "", "stream": i % 2 == 0, "priority": "medium",
                           "options": {"num_predict": 4, "ignore_eos": True}, "timeout": 300000, "metadata": {"prompt_token_ids": ids}})
            if i % 2 == 0:
                await sched.watch_stream(f"job-{i}")
        await sched.run_until_empty()
        for w in workers:
            await w.stop()
    _run(go())
    assert len(sched.results) == 9 and all("result" in r for r in sched.results.values())
    per_worker = {w: sum(1 for v in sched.assigned.values() if v == w) for w in ("b200-0", "b200-1")}
    assert min(per_worker.values()) >= 3                 # both workers were used, neither ever held more than three jobs
    assert all(max(s._engine("tiny:latest").batch_sizes) <= 3 for s in svcs)
    assert max(max(s._engine("tiny:latest").batch_sizes) for s in svcs) >= 2         # jobs really shared steps
    m = O.load_gguf(tiny_gguf)
    for i in range(9):
        ids = np.random.Generator(np.random.PCG64(2000 + i)).integers(0, 500, size=10)
        ref = O.LlamaOracle(m, act="i16", kv_f16=True).generate(ids, 4)
        r = sched.results[f"job-{i}"]["result"]
        if i % 2 == 0:
            assert sched.stream_chunks[f"job-{i}"] == 5 and r["done"] is True       # 4 tokens + the done chunk on job:stream:<id>
        else:
            assert r["token_ids"] == [int(t) for t in ref["ids"]]
    for s in svcs:
        s.close()
