"""TEST INFRASTRUCTURE: an engine double with the surface of gridllm_b200.native.Engine, backed by the numpy oracle and the
product tokenizer compiled for the host (tests/hostcheck).  It lets the CPU suite drive the UNMODIFIED host layer
(gridllm_b200/service.py, worker.py) end to end; the real engine takes its place in tests/test_gpu_service.py."""
import ctypes
import time
from types import SimpleNamespace

import numpy as np

from oracle import llama_oracle as O, sampler as SM

_LIB = None


def use_hostcheck(lib):
    global _LIB
    _LIB = lib


class OracleEngine:
    def __init__(self, gguf_path, device=0, max_ctx=0, **kw):
        assert _LIB is not None, "call oracle_engine.use_hostcheck(hostcheck_lib) first"
        self.path = gguf_path
        self.m = O.load_gguf(gguf_path)
        self.info = SimpleNamespace(has_tokenizer=1, n_vocab=self.m.n_vocab, n_embd=self.m.n_embd, n_params=int(1e6), quantization=b"Q4_K_M", n_ctx=int(max_ctx) if max_ctx else 512,
                                    eos_id=self.m.n_vocab - 2, eot_id=self.m.n_vocab - 1, bos_id=self.m.n_vocab - 3)
        try:                                     # bos / eos / eot ids as the file states them (a SentencePiece vocabulary keeps them at 1 / 2)
            import gguf
            fields = gguf.GGUFReader(gguf_path).fields

            def _u(key, dflt):
                f = fields.get(key)
                return int(f.parts[-1][0]) if f is not None else dflt
            self.info.bos_id = _u("tokenizer.ggml.bos_token_id", self.info.bos_id)
            self.info.eos_id = _u("tokenizer.ggml.eos_token_id", self.info.eos_id)
            self.info.eot_id = _u("tokenizer.ggml.eot_token_id", -1 if fields.get("tokenizer.ggml.scores") is not None else self.info.eot_id)
        except Exception:
            pass
        self.calls = []
        self.chat_template = ""
        self.max_batch = int(kw.get("max_batch", 0) or 0)

    # ---- tokenizer: the product's C++ code --------------------------------------------------------------------------
    def tokenize(self, text, add_bos=True, parse_special=False):
        ids = np.zeros(8192, np.int32)
        n = _LIB.hc_tokenize(self.path.encode(), text.encode("utf-8"), int(add_bos), int(parse_special), ids.ctypes.data_as(ctypes.c_void_p), 8192)
        assert n >= 0
        return ids[:n].copy()

    def _bytes(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        buf = ctypes.create_string_buffer(16 * max(1, len(a)) + 16)
        n = _LIB.hc_detokenize(self.path.encode(), a.ctypes.data_as(ctypes.c_void_p), len(a), buf, len(buf))
        assert n >= 0
        return buf.raw[:n]

    def detokenize(self, ids):
        return self._bytes(ids).decode("utf-8", "replace")

    # ---- generation: the oracle's forward, greedy or the oracle's seeded draw -----------------------------------------
    def generate(self, prompt, num_predict=128, ignore_eos=False, on_token=None, want_logits=False, stop_ids=(), temperature=0.0, top_k=0,
                 top_p=1.0, seed=0):
        self.calls.append(dict(n_prompt=len(prompt), num_predict=num_predict, temperature=temperature, top_k=top_k, top_p=top_p, seed=seed))
        t0 = time.perf_counter_ns()
        orc = O.LlamaOracle(self.m, act="i16", kv_f16=True)
        logits = None
        for t in prompt:
            logits = orc.step(int(t))
        t1 = time.perf_counter_ns()
        stops = set(int(s) for s in stop_ids) | ({self.info.eos_id, self.info.eot_id} if not ignore_eos else set())
        ids, lps, reason = [], [], 1
        for i in range(num_predict):
            tok, lp, _ = SM.sample(logits, temperature, top_k, top_p, seed, i)
            if tok in stops:
                reason = 0
                break
            ids.append(tok)
            lps.append(lp)
            if on_token is not None and on_token(tok, lp, self._bytes([tok])):
                reason = 2
                break
            logits = orc.step(tok)
        t2 = time.perf_counter_ns()
        stats = SimpleNamespace(prompt_eval_count=len(prompt), eval_count=len(ids), prompt_eval_duration_ns=t1 - t0, eval_duration_ns=max(1, t2 - t1),
                                total_duration_ns=t2 - t0, load_duration_ns=1, done_reason=reason, kernel_launches=0)
        return SimpleNamespace(ids=np.array(ids, dtype=np.int32), logprobs=np.array(lps, dtype=np.float32), stats=stats)

    # ---- continuous batching: the same oracle, one instance per open sequence, stepped together -------------------------
    def seq_open(self, prompt, num_predict=128, ignore_eos=False, temperature=0.0, top_k=0, top_p=1.0, seed=0, stop_ids=()):
        from gridllm_b200.native import NativeError
        if not getattr(self, "max_batch", 0):
            raise NativeError(-4, "continuous batching is off")
        if not hasattr(self, "_seqs"):
            self._seqs = {}
        if len(prompt) + num_predict > self.info.n_ctx:
            raise NativeError(-9, "prompt + num_predict exceeds the engine context")
        free = [i for i in range(self.max_batch) if i not in self._seqs]
        if not free:
            raise NativeError(-6, "no free sequence slot")
        self.calls.append(dict(n_prompt=len(prompt), num_predict=num_predict, temperature=temperature, top_k=top_k, top_p=top_p, seed=seed, batched=True))
        orc = O.LlamaOracle(self.m, act="i16", kv_f16=True)
        logits = None
        for t in prompt:
            logits = orc.step(int(t))
        stops = set(int(s) for s in stop_ids) | ({self.info.eos_id, self.info.eot_id} if not ignore_eos else set())
        self._seqs[free[0]] = SimpleNamespace(orc=orc, logits=logits, n=0, n_pred=num_predict, stops=stops, opts=(temperature, top_k, top_p, seed),
                                              n_prompt=len(prompt), done=False, stopped=False, t0=time.perf_counter_ns())
        return free[0]

    def seq_open_many(self, prompts, options):
        from gridllm_b200.native import NativeError
        self.open_many_calls = getattr(self, "open_many_calls", []) + [len(prompts)]
        for p, o in zip(prompts, options):          # the engine validates everything before it opens anything
            if len(p) + o.get("num_predict", 128) > self.info.n_ctx:
                raise NativeError(-9, "prompt + num_predict exceeds the engine context")
        slots = []
        for p, o in zip(prompts, options):
            try:
                slots.append(self.seq_open(p, **o))
            except NativeError as ex:
                if ex.code != -6:
                    raise
                slots.append(-1)
        if all(s < 0 for s in slots):
            raise NativeError(-6, "no free sequence slot")
        return slots

    def batch_step(self, cap=128):
        out = []
        self.batch_sizes = getattr(self, "batch_sizes", [])
        live = [(slot, q) for slot, q in sorted(getattr(self, "_seqs", {}).items()) if not q.done]
        self.batch_sizes.append(len(live))
        for slot, q in live:
            tok, lp, _ = SM.sample(q.logits, *q.opts, q.n)
            if tok in q.stops:
                q.done = q.stopped = True
                out.append((slot, -1, 0.0, True))
                continue
            q.n += 1
            q.done = q.n >= q.n_pred
            out.append((slot, int(tok), float(lp), q.done))
            if not q.done:
                q.logits = q.orc.step(int(tok))
        return out

    def seq_stats(self, slot):
        q = self._seqs[slot]
        dt = time.perf_counter_ns() - q.t0
        return SimpleNamespace(prompt_eval_count=q.n_prompt, eval_count=q.n, prompt_eval_duration_ns=1, eval_duration_ns=max(1, dt), total_duration_ns=dt,
                               load_duration_ns=1, done_reason=0 if q.stopped else 1, kernel_launches=0)

    def seq_close(self, slot):
        from gridllm_b200.native import NativeError
        if slot not in getattr(self, "_seqs", {}):
            raise NativeError(-1, "seq_close: no such open sequence")
        del self._seqs[slot]

    def token_piece(self, tid):
        return self._bytes([tid])

    def token_text(self, tid):
        """gl_token_text: the vocabulary's spelling, control tokens included (the double reads it from the GGUF metadata)"""
        toks = getattr(self, "_vocab_texts", None)
        if toks is None:
            import gguf                                   # gguf-py, the ggml project's own reader
            f = gguf.GGUFReader(self.path).fields.get("tokenizer.ggml.tokens")
            toks = self._vocab_texts = [bytes(f.parts[i]).decode("utf-8", "replace") for i in f.data] if f is not None else []
        return toks[tid] if 0 <= tid < len(toks) else ""

    def embed(self, seqs):
        if any(len(s) > self.info.n_ctx for s in seqs):
            raise RuntimeError("GL_ERR_CONTEXT: sequence exceeds the engine context")
        orc = O.LlamaOracle(self.m, act="i16")
        out = np.stack([orc.embed(s) for s in seqs]).astype(np.float32)
        return out, SimpleNamespace(prompt_eval_count=int(sum(len(s) for s in seqs)), load_duration_ns=1)

    def close(self):
        pass
