"""TEST INFRASTRUCTURE -- numpy restatement of the inference path the native worker replaces.

What the reference does on this path: ``OllamaService.generateResponse`` /
``generateStreamResponse`` / ``generateEmbedding``
(/root/reference/client/src/services/OllamaService.ts:97-184, 186-284, 601-665) POST the
prompt to an Ollama daemon and map the answer back.  The arithmetic itself -- GGUF load,
tokenise, prefill, decode loop, greedy sampling, embedding pooling -- is in Ollama's bundled
llama.cpp/ggml, which is NOT in /root/reference and is not version-pinned
(docs/deployment/docker-compose.dependencies.yml:14).  This file restates that published
algorithm (Llama architecture over GGUF tensors) so the CUDA path has something to be
checked against.  PARITY UNPINNED w.r.t. the reference (it holds no golden vectors); pinned
instead against gguf-py dequantisers and transformers' LlamaForCausalLM
(tests/test_oracle_pin.py).

Numerics modes
  act="exact" : y = W_deq @ x with exact dequantised weights, float64 accumulate.  The
                specification ("mode A").
  act="i16"   : x is first snapped to the engine's per-32-block 15-bit fixed point
                (x ~ sx*(128*hi+lo), hi/lo int8) -- what the CUDA GEMV consumes.  Lets the
                tests separate "kernel arithmetic wrong" from "activation rounding".
  act="q8"    : x snapped to int8 per 32-block (the engine's act_bits=8 option).
  act="ggml"  : x snapped the way ggml's CPU backend pairs activations with each weight type
                (K-quants: Q8_K, one scale per 256 columns; Q8_0: one per 32) -- "mode B", used
                to pin the C baseline and to QUANTIFY the divergence a real Ollama would show.
  kv_f16=True : K/V rounded to fp16 when written to the cache, as the engine stores them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import gguf_synth as S


# --------------------------------------------------------------------------------------
# GGUF loading through the INDEPENDENT gguf-py reader + dequantisers
# --------------------------------------------------------------------------------------

@dataclass
class OracleModel:
    arch: str
    n_layer: int
    n_embd: int
    n_head: int
    n_head_kv: int
    n_ff: int
    n_vocab: int
    head_dim: int
    rope_base: float
    rms_eps: float
    n_ctx: int
    raw: Dict[str, Tuple[int, Tuple[int, ...], np.ndarray]]   # name -> (type, shape, bytes)
    meta: Dict[str, object]
    _cache: Dict[str, np.ndarray]

    def w(self, name: str) -> np.ndarray:
        """Dequantised fp32 tensor [rows, cols] (cached)."""
        if name not in self._cache:
            t, shape, data = self.raw[name]
            self._cache[name] = dequantize(data, t, shape)
        return self._cache[name]

    def has(self, name: str) -> bool:
        return name in self.raw

    def drop_cache(self):
        self._cache.clear()


def dequantize(data: np.ndarray, ggml_type: int, shape: Sequence[int]) -> np.ndarray:
    from gguf import quants, GGMLQuantizationType as T
    rows = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
    cols = shape[-1]
    b = np.ascontiguousarray(data).view(np.uint8).reshape(rows, -1)
    if ggml_type == S.F32:
        return b.view(np.float32).reshape(shape).copy()
    if ggml_type == S.F16:
        return b.view(np.float16).astype(np.float32).reshape(shape)
    if ggml_type == S.BF16:
        return (b.view(np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(shape)
    return quants.dequantize(b, T(ggml_type)).reshape(shape).astype(np.float32)


def load_gguf(path: str) -> OracleModel:
    from gguf import GGUFReader
    r = GGUFReader(path)

    def field(key, default=None):
        f = r.get_field(key)
        if f is None:
            return default
        return f.contents()

    arch = field("general.architecture")
    raw = {}
    for t in r.tensors:
        shape = tuple(int(d) for d in reversed(t.shape.tolist()))
        raw[t.name] = (int(t.tensor_type), shape, np.asarray(t.data).view(np.uint8).reshape(-1))
    n_embd = int(field(f"{arch}.embedding_length"))
    n_head = int(field(f"{arch}.attention.head_count"))
    meta = {k: None for k in r.fields}
    m = OracleModel(
        arch=arch,
        n_layer=int(field(f"{arch}.block_count")),
        n_embd=n_embd,
        n_head=n_head,
        n_head_kv=int(field(f"{arch}.attention.head_count_kv", n_head)),
        n_ff=int(field(f"{arch}.feed_forward_length")),
        n_vocab=raw["token_embd.weight"][1][0],
        head_dim=int(field(f"{arch}.rope.dimension_count", n_embd // n_head)),
        rope_base=float(field(f"{arch}.rope.freq_base", 10000.0)),
        rms_eps=float(field(f"{arch}.attention.layer_norm_rms_epsilon", 1e-5)),
        n_ctx=int(field(f"{arch}.context_length", 2048)),
        raw=raw, meta=meta, _cache={},
    )
    st = field(f"{arch}.rope.scaling.type", "") or ""
    if st not in ("", "none", "linear"):
        raise ValueError(f"rope.scaling.type {st!r} is not restated (none / linear only)")
    m.rope_linear = float(field(f"{arch}.rope.scaling.factor", 1.0)) if st == "linear" else 1.0
    m.meta["reader"] = r
    return m


# --------------------------------------------------------------------------------------
# activation fixed point (the engine's GEMV input format) and ggml-style int8
# --------------------------------------------------------------------------------------

def snap_i16(x: np.ndarray) -> np.ndarray:
    """Per-32-block 15-bit fixed point: sx = amax/16256 (fp32), v = rint(x/sx) -> sx*v."""
    x32 = np.asarray(x, dtype=np.float32).reshape(-1, 32)
    amax = np.abs(x32).max(axis=1, keepdims=True)
    sx = (amax / np.float32(16256.0)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(amax > 0, np.float32(16256.0) / amax, np.float32(0)).astype(np.float32)
    v = np.rint(x32 * inv)
    return (v.astype(np.float64) * sx.astype(np.float64)).reshape(np.shape(x))


def snap_q8(x: np.ndarray) -> np.ndarray:
    x32 = np.asarray(x, dtype=np.float32).reshape(-1, 32)
    amax = np.abs(x32).max(axis=1, keepdims=True)
    sx = (amax / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(amax > 0, np.float32(127.0) / amax, np.float32(0)).astype(np.float32)
    v = np.rint(x32 * inv)
    return (v.astype(np.float64) * sx.astype(np.float64)).reshape(np.shape(x))


def snap_q8k(x: np.ndarray) -> np.ndarray:
    """ggml Q8_K-style: one int8 scale per 256 columns (d = amax/127)."""
    x32 = np.asarray(x, dtype=np.float32).reshape(-1, 256)
    amax = np.abs(x32).max(axis=1, keepdims=True)
    sx = (amax / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(amax > 0, np.float32(127.0) / amax, np.float32(0)).astype(np.float32)
    v = np.rint(x32 * inv)
    return (v.astype(np.float64) * sx.astype(np.float64)).reshape(np.shape(x))


def _snap(x: np.ndarray, act: str, wtype: Optional[int] = None) -> np.ndarray:
    if act == "ggml":      # the activation format ggml's CPU backend pairs with each weight type
        if wtype in (S.Q4_K, S.Q6_K):
            return snap_q8k(x)
        if wtype == S.Q8_0:
            return snap_q8(x)
        return np.asarray(x, dtype=np.float64)
    if act == "exact":
        return np.asarray(x, dtype=np.float64)
    if act == "i16":
        return snap_i16(x)
    if act == "q8":
        return snap_q8(x)
    raise ValueError(act)


def gemv(w: np.ndarray, x: np.ndarray, act: str = "exact", wtype: Optional[int] = None) -> np.ndarray:
    """y = W @ snap(x); W fp32 exact dequantised values, float64 accumulate."""
    return w.astype(np.float64) @ _snap(x, act, wtype)


# --------------------------------------------------------------------------------------
# Llama building blocks
# --------------------------------------------------------------------------------------

def rmsnorm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return x / np.sqrt(np.mean(x * x, axis=-1, keepdims=True) + eps) * w.astype(np.float64)


def rope_table(n_pos: int, n_rot: int, base: float, freq_factors: Optional[np.ndarray] = None,
               linear_factor: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """cos/sin [n_pos, n_rot/2] as fp32.  inv_freq_i = base^(-2i/n_rot) rounded to fp32, the
    angle pos*inv_freq_i is formed in fp32, cos/sin evaluated in float64 and rounded to fp32.
    (ggml 'NORM' rope, adjacent pairs (x[2i], x[2i+1]); [external] llama.cpp ggml_rope.)"""
    i = np.arange(n_rot // 2, dtype=np.float64)
    inv = (base ** (-2.0 * i / n_rot)).astype(np.float32)
    if freq_factors is not None:          # Llama-3.1+ rope_freqs.weight: theta_i / factor_i ([external] llama.cpp rope with freq_factors)
        inv = (inv / np.asarray(freq_factors, dtype=np.float32)).astype(np.float32)
    p = (np.arange(n_pos, dtype=np.float32) / np.float32(linear_factor)).astype(np.float32)      # rope.scaling.type == linear
    ang = (p[:, None] * inv[None, :]).astype(np.float32)
    return np.cos(ang.astype(np.float64)).astype(np.float32), np.sin(ang.astype(np.float64)).astype(np.float32)


def apply_rope(v: np.ndarray, pos: int, n_heads: int, head_dim: int, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64).reshape(n_heads, head_dim // 2, 2)
    c = cos[pos].astype(np.float64)[None, :]
    s = sin[pos].astype(np.float64)[None, :]
    out = np.empty_like(v)
    out[:, :, 0] = v[:, :, 0] * c - v[:, :, 1] * s
    out[:, :, 1] = v[:, :, 0] * s + v[:, :, 1] * c
    return out.reshape(-1)


def silu(x: np.ndarray) -> np.ndarray:
    return x / (1.0 + np.exp(-x))


class LlamaOracle:
    """Token-at-a-time forward with a KV cache (prefill == repeated decode steps; the result
    does not depend on batching in exact arithmetic)."""

    def __init__(self, model: OracleModel, act: str = "exact", kv_f16: bool = True):
        self.m = model
        self.act = act
        self.kv_f16 = kv_f16
        ff = model.w("rope_freqs.weight").reshape(-1) if model.has("rope_freqs.weight") else None
        self.cos, self.sin = rope_table(model.n_ctx, model.head_dim, model.rope_base, ff, getattr(model, "rope_linear", 1.0))
        self.reset()

    def reset(self):
        self.k: List[List[np.ndarray]] = [[] for _ in range(self.m.n_layer)]
        self.v: List[List[np.ndarray]] = [[] for _ in range(self.m.n_layer)]
        self.pos = 0

    def _kv_round(self, a: np.ndarray) -> np.ndarray:
        if self.kv_f16:
            return a.astype(np.float32).astype(np.float16).astype(np.float64)
        return a

    def hidden_step(self, token: int) -> np.ndarray:
        """Run one token through all layers; returns the final residual stream (pre output_norm)."""
        m, act = self.m, self.act
        H, KV, hd = m.n_head, m.n_head_kv, m.head_dim
        x = m.w("token_embd.weight")[token].astype(np.float64)
        pos = self.pos

        def mv(name, vec):
            return gemv(m.w(name), vec, act, m.raw[name][0])
        for il in range(m.n_layer):
            p = f"blk.{il}."
            h = rmsnorm(x, m.w(p + "attn_norm.weight"), m.rms_eps)
            q = mv(p + "attn_q.weight", h)
            k = mv(p + "attn_k.weight", h)
            v = mv(p + "attn_v.weight", h)
            q = apply_rope(q, pos, H, hd, self.cos, self.sin)
            k = apply_rope(k, pos, KV, hd, self.cos, self.sin)
            self.k[il].append(self._kv_round(k))
            self.v[il].append(self._kv_round(v))
            Kc = np.stack(self.k[il]).reshape(pos + 1, KV, hd)
            Vc = np.stack(self.v[il]).reshape(pos + 1, KV, hd)
            qh = q.reshape(H, hd)
            grp = H // KV
            att = np.empty((H, hd), dtype=np.float64)
            for hh in range(H):
                kvh = hh // grp
                s = Kc[:, kvh, :] @ qh[hh] / np.sqrt(hd)
                s = s - s.max()
                pw = np.exp(s)
                pw /= pw.sum()
                att[hh] = pw @ Vc[:, kvh, :]
            x = x + mv(p + "attn_output.weight", att.reshape(-1))
            h2 = rmsnorm(x, m.w(p + "ffn_norm.weight"), m.rms_eps)
            g = mv(p + "ffn_gate.weight", h2)
            u = mv(p + "ffn_up.weight", h2)
            x = x + mv(p + "ffn_down.weight", silu(g) * u)
        self.pos += 1
        return x

    def logits_from_hidden(self, x: np.ndarray) -> np.ndarray:
        m = self.m
        h = rmsnorm(x, m.w("output_norm.weight"), m.rms_eps)
        wname = "output.weight" if m.has("output.weight") else "token_embd.weight"
        return gemv(m.w(wname), h, self.act, m.raw[wname][0])

    def step(self, token: int) -> np.ndarray:
        return self.logits_from_hidden(self.hidden_step(token))

    def generate(self, prompt: Sequence[int], n_predict: int) -> Dict[str, np.ndarray]:
        """Greedy decode.  Returns ids, logprobs of the chosen ids, top-1/top-2 logit margins,
        and the full logits of every generation step."""
        self.reset()
        logits = None
        for t in prompt:
            logits = self.step(int(t))
        ids, lps, margins, all_logits = [], [], [], []
        for _ in range(n_predict):
            top = int(np.argmax(logits))
            lse = logits.max() + np.log(np.exp(logits - logits.max()).sum())
            srt = np.partition(logits, -2)[-2:]
            ids.append(top)
            lps.append(float(logits[top] - lse))
            margins.append(float(srt[1] - srt[0]))
            all_logits.append(logits.astype(np.float32))
            logits = self.step(top)
        return {"ids": np.array(ids, dtype=np.int32), "logprobs": np.array(lps, dtype=np.float32),
                "margins": np.array(margins, dtype=np.float32), "logits": np.stack(all_logits)}

    def embed(self, tokens: Sequence[int], pooling: str = "mean", normalize: bool = True) -> np.ndarray:
        """Embedding path (OllamaService.generateEmbedding, OllamaService.ts:601-665): prefill only,
        output_norm applied, pool over positions, L2-normalise ([external] Ollama /api/embed
        normalises; llama-family GGUFs without pooling metadata mean-pool -- pinned here)."""
        self.reset()
        hs = []
        for t in tokens:
            x = self.hidden_step(int(t))
            hs.append(rmsnorm(x, self.m.w("output_norm.weight"), self.m.rms_eps))
        hs = np.stack(hs)
        e = hs.mean(axis=0) if pooling == "mean" else hs[-1]
        if normalize:
            e = e / max(np.linalg.norm(e), 1e-12)
        return e.astype(np.float32)
