"""TEST INFRASTRUCTURE -- synthetic GGUF v3 writer, block quantisers and model builders.

There is no network, no real checkpoint and no tokenizer file in this image
(SURVEY.md section 0.4), so every model the tests, the smoke and the bench use is a
synthetic GGUF produced here.  The writer is a from-scratch restatement of the public
GGUF v3 container layout; ``tests/test_gguf_synth.py`` pins it by reading its files back
with the independent ``gguf.GGUFReader`` and by dequantising its blocks with
``gguf.quants.dequantize``.

Block layouts restated (ggml public format; cross-checked against gguf-py
``gguf/quants.py`` Q8_0 l.378-401, Q4_K l.475-522, Q6_K l.552-572):
  Q8_0  34 B / 32 w : f16 d | int8 q[32]                      w = d*q
  Q4_K 144 B /256 w : f16 d | f16 dmin | u8 scales[12] | u8 qs[128]
                      w = d*sc[j]*q - dmin*m[j],  j = sub-block of 32
  Q6_K 210 B /256 w : u8 ql[128] | u8 qh[64] | int8 scales[16] | f16 d
                      w = d*scales[j]*(q-32),     j = sub-block of 16
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# ggml type ids (public enum)
F32, F16, Q8_0, Q4_K, Q6_K, BF16 = 0, 1, 8, 12, 14, 30
TYPE_NAMES = {F32: "F32", F16: "F16", Q8_0: "Q8_0", Q4_K: "Q4_K", Q6_K: "Q6_K", BF16: "BF16"}
# (weights per block, bytes per block)
BLOCK = {F32: (1, 4), F16: (1, 2), BF16: (1, 2), Q8_0: (32, 34), Q4_K: (256, 144), Q6_K: (256, 210)}

GGUF_MAGIC = 0x46554747
ALIGN = 32
# GGUF metadata value types
_U8, _I8, _U16, _I16, _U32, _I32, _F32, _BOOL, _STR, _ARR, _U64, _I64, _F64 = range(13)


def row_bytes(ggml_type: int, cols: int) -> int:
    w, b = BLOCK[ggml_type]
    assert cols % w == 0, (ggml_type, cols)
    return cols // w * b


# --------------------------------------------------------------------------------------
# quantisers: fp32 -> *valid* blocks.  Not bit-identical to ggml's search-based
# quantisers (not on the hot path) -- only the block FORMAT matters.
# --------------------------------------------------------------------------------------

def quantize_q8_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 32)
    d = np.abs(x).max(axis=1, keepdims=True) / 127.0
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(d == 0, 0.0, 1.0 / d)
    q = np.rint(x * inv).astype(np.int8)
    out = np.empty((x.shape[0], 34), dtype=np.uint8)
    out[:, :2] = d.astype(np.float16).view(np.uint8)
    out[:, 2:] = q.view(np.uint8)
    return out


def pack_q4k_scales(sc: np.ndarray, mn: np.ndarray) -> np.ndarray:
    """sc, mn: (n, 8) uint8 in [0, 63] -> (n, 12) packed 6-bit fields."""
    sc = sc.astype(np.uint8)
    mn = mn.astype(np.uint8)
    out = np.empty((sc.shape[0], 12), dtype=np.uint8)
    out[:, 0:4] = (sc[:, 0:4] & 63) | ((sc[:, 4:8] >> 4) << 6)
    out[:, 4:8] = (mn[:, 0:4] & 63) | ((mn[:, 4:8] >> 4) << 6)
    out[:, 8:12] = (sc[:, 4:8] & 0xF) | ((mn[:, 4:8] & 0xF) << 4)
    return out


def quantize_q4_k(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 8, 32)
    n = x.shape[0]
    lo = np.minimum(x.min(axis=2), 0.0)           # (n, 8)  <= 0
    hi = np.maximum(x.max(axis=2), lo)
    sub_scale = (hi - lo) / 15.0
    sub_min = -lo
    d = (sub_scale.max(axis=1, keepdims=True) / 63.0).astype(np.float16).astype(np.float32)
    dmin = (sub_min.max(axis=1, keepdims=True) / 63.0).astype(np.float16).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        sc = np.where(d > 0, np.rint(sub_scale / d), 0).clip(0, 63)
        mn = np.where(dmin > 0, np.rint(sub_min / dmin), 0).clip(0, 63)
        eff = (d * sc)[:, :, None]
        q = np.where(eff > 0, np.rint((x + (dmin * mn)[:, :, None]) / eff), 0).clip(0, 15)
    q = q.astype(np.uint8)                         # (n, 8, 32)
    qs = (q[:, 0::2, :] | (q[:, 1::2, :] << 4)).reshape(n, 128)
    out = np.empty((n, 144), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8)
    out[:, 2:4] = dmin.astype(np.float16).view(np.uint8)
    out[:, 4:16] = pack_q4k_scales(sc, mn)
    out[:, 16:] = qs
    return out


def quantize_q6_k(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 16, 16)
    n = x.shape[0]
    amax = np.abs(x).max(axis=2)                  # (n, 16)
    sub_scale = amax / 31.0
    d = (sub_scale.max(axis=1, keepdims=True) / 127.0).astype(np.float16).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        sc = np.where(d > 0, np.rint(sub_scale / d), 0).clip(1, 127)
        eff = (d * sc)[:, :, None]
        q = np.where(eff > 0, np.rint(x / eff), 0).clip(-32, 31)
    q = (q.astype(np.int16) + 32).astype(np.uint8).reshape(n, 256)   # 0..63
    # element e = h*128 + s*64 + i (ql nibble s of byte h*64+i); e = h*128 + t*32 + j (qh bits 2t of byte h*32+j)
    qe = q.reshape(n, 2, 2, 64)
    ql = ((qe[:, :, 0, :] & 0xF) | ((qe[:, :, 1, :] & 0xF) << 4)).reshape(n, 128)
    qt = (q.reshape(n, 2, 4, 32) >> 4) & 3
    qh = (qt[:, :, 0, :] | (qt[:, :, 1, :] << 2) | (qt[:, :, 2, :] << 4) | (qt[:, :, 3, :] << 6)).reshape(n, 64)
    out = np.empty((n, 210), dtype=np.uint8)
    out[:, 0:128] = ql
    out[:, 128:192] = qh
    out[:, 192:208] = sc.astype(np.int8).view(np.uint8)
    out[:, 208:210] = d.astype(np.float16).view(np.uint8)
    return out


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    u = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))     # round to nearest even
    return (u >> np.uint32(16)).astype(np.uint16)


def quantize(x: np.ndarray, ggml_type: int) -> np.ndarray:
    """fp32 [rows, cols] -> uint8 [rows, row_bytes] in GGUF layout."""
    rows, cols = x.shape
    if ggml_type == F32:
        return np.ascontiguousarray(x, dtype=np.float32).view(np.uint8).reshape(rows, -1)
    if ggml_type == F16:
        return np.ascontiguousarray(x.astype(np.float16)).view(np.uint8).reshape(rows, -1)
    if ggml_type == BF16:
        return to_bf16_bits(x).view(np.uint8).reshape(rows, -1)
    fn = {Q8_0: quantize_q8_0, Q4_K: quantize_q4_k, Q6_K: quantize_q6_k}[ggml_type]
    return fn(x).reshape(rows, row_bytes(ggml_type, cols))


def random_blocks(rng: np.random.Generator, ggml_type: int, rows: int, cols: int,
                  gain: float = 1.0) -> np.ndarray:
    """Raw random but well-formed blocks ('kernel-only tests may use raw
    random blocks').  Scale fields are chosen so the dequantised weights are zero-mean with
    standard deviation ~ gain/sqrt(cols); cheap enough to build Llama-3-8B shapes in seconds."""
    nb = rows * cols // BLOCK[ggml_type][0]
    bb = BLOCK[ggml_type][1]
    target = gain / np.sqrt(cols)
    if ggml_type in (F16, BF16) and rows * cols >= (1 << 24):
        # big 16-bit matrices (the bf16 Llama-3-8B of BASELINE config 4 is 16 GB): random BIT PATTERNS instead of 8 G normal
        # draws -- random sign and mantissa, exponent one of two adjacent values around `target`, so |w| is spread over
        # [target/2, 2*target) with zero mean: finite, well-scaled weights at a few GB/s of generation
        n = rows * cols
        man_bits = 7 if ggml_type == BF16 else 10
        bias = 127 if ggml_type == BF16 else 15
        e0 = int(np.floor(np.log2(target))) + bias
        raw = rng.bit_generator.random_raw((n + 3) // 4).view(np.uint16)[:n]           # 16 random bits per weight, ~GB/s
        ebit = raw & np.uint16(1 << man_bits)                                            # one random bit picks the exponent
        raw &= np.uint16(0x8000 | ((1 << man_bits) - 1))                                 # keep sign and mantissa
        raw += np.uint16(e0 << man_bits)
        raw -= ebit
        out16 = raw
        return out16.view(np.uint8).reshape(rows, cols * 2)
    if ggml_type in (F32, F16, BF16):
        w = rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(target)
        return quantize(w, ggml_type)
    out = rng.integers(0, 256, size=(nb, bb), dtype=np.uint8)
    u = rng.random(nb, dtype=np.float32) + 1.0                       # U[1,2)
    if ggml_type == Q4_K:
        # w = d*sc*q - dmin*m ; sc,m ~ U{0..63}, q ~ U{0..15}; dmin = 7.5 d makes E[w] = 0.
        # std(w)/d ~ 258 (analytic and measured); d chosen accordingly.
        d = (u * np.float32(target / (1.5 * 258.0))).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
        out[:, 2:4] = (d.astype(np.float32) * np.float32(7.5)).astype(np.float16).view(np.uint8).reshape(nb, 2)
    elif ggml_type == Q6_K:
        # w = d*s*(q-32); s ~ U{-128..127}, q-32 ~ U{-32..31}: std(w)/d ~ 74*18.5 ~ 1365
        d = (u * np.float32(target / (1.5 * 1365.0))).astype(np.float16)
        out[:, 208:210] = d.view(np.uint8).reshape(nb, 2)
    elif ggml_type == Q8_0:
        # w = d*q, q ~ U{-128..127}: std/d ~ 74
        d = (u * np.float32(target / (1.5 * 74.0))).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
    else:
        raise ValueError(ggml_type)
    return out.reshape(rows, row_bytes(ggml_type, cols))


# --------------------------------------------------------------------------------------
# GGUF v3 container writer
# --------------------------------------------------------------------------------------

def _s(b: str) -> bytes:
    e = b.encode("utf-8")
    return struct.pack("<Q", len(e)) + e


def _kv(key: str, vtype: int, val) -> bytes:
    out = _s(key) + struct.pack("<I", vtype)
    return out + _val(vtype, val)


def _val(vtype: int, val) -> bytes:
    fmt = {_U8: "<B", _I8: "<b", _U16: "<H", _I16: "<h", _U32: "<I", _I32: "<i", _F32: "<f",
           _BOOL: "<?", _U64: "<Q", _I64: "<q", _F64: "<d"}
    if vtype in fmt:
        return struct.pack(fmt[vtype], val)
    if vtype == _STR:
        return _s(val)
    if vtype == _ARR:
        etype, items = val
        if etype == _STR:
            body = b"".join(_s(i) for i in items)
        else:
            np_t = {_U8: "<u1", _I8: "<i1", _U16: "<u2", _I16: "<i2", _U32: "<u4", _I32: "<i4",
                    _F32: "<f4", _U64: "<u8", _I64: "<i8", _F64: "<f8"}[etype]
            body = np.asarray(items, dtype=np_t).tobytes()
        return struct.pack("<IQ", etype, len(items)) + body
    raise ValueError(vtype)


@dataclass
class TensorSpec:
    name: str
    ggml_type: int
    shape: Tuple[int, ...]          # numpy order: (rows, cols) -> GGUF ne = (cols, rows)
    data: np.ndarray                # uint8 bytes in GGUF layout


@dataclass
class GGUFFile:
    kv: List[bytes] = field(default_factory=list)
    tensors: List[TensorSpec] = field(default_factory=list)

    def add_u32(self, k, v): self.kv.append(_kv(k, _U32, int(v)))
    def add_f32(self, k, v): self.kv.append(_kv(k, _F32, float(v)))
    def add_str(self, k, v): self.kv.append(_kv(k, _STR, v))
    def add_bool(self, k, v): self.kv.append(_kv(k, _BOOL, bool(v)))
    def add_arr(self, k, etype, items): self.kv.append(_kv(k, _ARR, (etype, items)))

    def add_tensor(self, name: str, ggml_type: int, shape: Sequence[int], data: np.ndarray):
        data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        rows = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
        assert data.nbytes == rows * row_bytes(ggml_type, shape[-1]), (name, data.nbytes)
        self.tensors.append(TensorSpec(name, ggml_type, tuple(int(s) for s in shape), data))

    def write(self, path: str) -> int:
        head = struct.pack("<IIQQ", GGUF_MAGIC, 3, len(self.tensors), len(self.kv))
        body = b"".join(self.kv)
        infos, off = [], 0
        for t in self.tensors:
            ne = tuple(reversed(t.shape))
            infos.append(_s(t.name) + struct.pack("<I", len(ne)) + struct.pack(f"<{len(ne)}Q", *ne)
                         + struct.pack("<IQ", t.ggml_type, off))
            off += (t.data.nbytes + ALIGN - 1) // ALIGN * ALIGN
        pre = head + body + b"".join(infos)
        pad = (-len(pre)) % ALIGN
        with open(path, "wb") as f:
            f.write(pre + b"\0" * pad)
            for t in self.tensors:
                t.data.tofile(f)
                p = (-t.data.nbytes) % ALIGN
                if p:
                    f.write(b"\0" * p)
            return f.tell()


# --------------------------------------------------------------------------------------
# model builders
# --------------------------------------------------------------------------------------

@dataclass
class LlamaShape:
    name: str
    n_layer: int
    n_embd: int
    n_head: int
    n_head_kv: int
    n_ff: int
    n_vocab: int
    rope_base: float = 10000.0
    rms_eps: float = 1e-5
    n_ctx: int = 2048

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head


TINY = LlamaShape("tiny-llama-synth", 2, 256, 4, 2, 512, 512, 10000.0, 1e-5, 512)
# second tiny shape with n_ff that is not a power of two and GQA 4:1, head_dim 128 like Llama-3
TINY128 = LlamaShape("tiny128-llama-synth", 2, 512, 4, 1, 768, 1024, 500000.0, 1e-5, 1024)
# four K-blocks per row and 8-12 output tiles per GEMM: the smallest shape on which the batched step's quantised GEMMs take their
# cluster path (tile-aligned split-K, folded RMSNorm) -- the 256 / 512-wide models above run the stream-K path only
MID = LlamaShape("mid-llama-synth", 2, 1024, 8, 2, 2048, 1024, 500000.0, 1e-5, 1024)
TINYLLAMA_1B = LlamaShape("tinyllama-1.1b-synth", 22, 2048, 32, 4, 5632, 32000, 10000.0, 1e-5, 2048)
LLAMA3_8B = LlamaShape("llama3-8b-synth", 32, 4096, 32, 8, 14336, 128256, 500000.0, 1e-5, 8192)
MISTRAL_7B = LlamaShape("mistral-7b-synth", 32, 4096, 32, 8, 14336, 32000, 10000.0, 1e-5, 8192)


def llama3_rope_factors(head_dim: int, base: float, factor: float = 8.0, low_freq_factor: float = 1.0,
                         high_freq_factor: float = 4.0, original_ctx: int = 8192) -> np.ndarray:
    """rope_freqs.weight as the Llama-3.1 GGUF converter computes it from the checkpoint's rope_scaling block
    ([external] convert_hf_to_gguf.py, rope_type "llama3"): 1 for short wavelengths, `factor` for long ones, a smooth
    ramp between.  inv_freq_i / factor_i is what the model then rotates by."""
    freqs = 1.0 / (base ** (np.arange(0, head_dim, 2, dtype=np.float64) / head_dim))
    low_wl, high_wl = original_ctx / low_freq_factor, original_ctx / high_freq_factor
    out = []
    for f in freqs:
        wl = 2 * np.pi / f
        if wl < high_wl:
            out.append(1.0)
        elif wl > low_wl:
            out.append(factor)
        else:
            smooth = (original_ctx / wl - low_freq_factor) / (high_freq_factor - low_freq_factor)
            out.append(1.0 / ((1 - smooth) / factor + smooth))
    return np.asarray(out, dtype=np.float32)


def q4_k_m_uses_q6(i_layer: int, n_layer: int) -> bool:
    """llama.cpp q4_K_M 'use_more_bits' rule for attn_v / ffn_down ([external])."""
    return (i_layer < n_layer // 8 or i_layer >= 7 * n_layer // 8
            or (i_layer - n_layer // 8) % 3 == 2)


def tensor_plan(shape: LlamaShape, recipe: str) -> List[Tuple[str, int, Tuple[int, int]]]:
    """(name, ggml_type, (rows, cols)) for every matrix; recipe in {q4_k_m, q8_0, f16, bf16, f32}."""
    E, H, KV, FF, V, hd = shape.n_embd, shape.n_head, shape.n_head_kv, shape.n_ff, shape.n_vocab, shape.head_dim

    def ty(kind: str, il: int) -> int:
        if recipe == "q4_k_m":
            if kind == "output":
                return Q6_K
            if kind in ("attn_v", "ffn_down") and q4_k_m_uses_q6(il, shape.n_layer):
                return Q6_K
            return Q4_K
        return {"q8_0": Q8_0, "f16": F16, "bf16": BF16, "f32": F32, "q4_k": Q4_K, "q6_k": Q6_K}[recipe]

    plan = [("token_embd.weight", ty("token_embd", 0), (V, E))]
    for il in range(shape.n_layer):
        p = f"blk.{il}."
        plan += [
            (p + "attn_norm.weight", F32, (1, E)),
            (p + "attn_q.weight", ty("attn_q", il), (H * hd, E)),
            (p + "attn_k.weight", ty("attn_k", il), (KV * hd, E)),
            (p + "attn_v.weight", ty("attn_v", il), (KV * hd, E)),
            (p + "attn_output.weight", ty("attn_output", il), (E, H * hd)),
            (p + "ffn_norm.weight", F32, (1, E)),
            (p + "ffn_gate.weight", ty("ffn_gate", il), (FF, E)),
            (p + "ffn_up.weight", ty("ffn_up", il), (FF, E)),
            (p + "ffn_down.weight", ty("ffn_down", il), (E, FF)),
        ]
    plan += [("output_norm.weight", F32, (1, E)), ("output.weight", ty("output", 0), (V, E))]
    return plan


def synth_vocab(n_vocab: int) -> Tuple[List[str], List[str], List[int]]:
    """A tiny byte-level BPE vocabulary in GPT-2 style: 256 byte tokens, then merges of
    adjacent byte tokens, then filler; last three ids are specials.  Token types: 1 normal,
    3 control."""
    b2u = gpt2_byte_to_unicode()
    toks = [b2u[b] for b in range(256)]
    merges: List[str] = []
    common = [" t", "he", " a", "in", "re", "on", " the", "er", " s", "at", "en", " w", "or",
              "nd", " c", "it", "es", "is", " b", "an", " p", "ou", "ing", " f", "al", "ar",
              " m", "ll", " o", " d", "ed", " in", "lo", "el", "hel", "hello", " wor", "ld", " world"]
    have = set(toks)
    for m in common:
        u = "".join(b2u[c] for c in m.encode())
        if u in have or len(toks) >= n_vocab - 3:
            continue
        # find a split into two existing tokens
        for k in range(1, len(u)):
            if u[:k] in have and u[k:] in have:
                merges.append(f"{u[:k]} {u[k:]}")
                toks.append(u)
                have.add(u)
                break
    i = 0
    while len(toks) < n_vocab - 3:
        toks.append(f"<filler_{i}>")
        i += 1
    toks = toks[: n_vocab - 3] + ["<|begin_of_text|>", "<|end_of_text|>", "<|eot_id|>"]
    types = [1] * (n_vocab - 3) + [3, 3, 3]
    for j in range(256 + len(merges), n_vocab - 3):
        types[j] = 5          # unused filler
    return toks, merges, types


def gpt2_byte_to_unicode() -> Dict[int, str]:
    bs = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


def build_model(path: str, shape: LlamaShape, recipe: str = "q4_k_m", seed: int = 1234,
                mode: str = "quantize", gain: float = 1.0, with_vocab: bool = True,
                arch: str = "llama", rope_freqs: Optional[np.ndarray] = None,
                rope_scaling: Optional[Tuple[str, float]] = None, pre: str = "llama-bpe",
                spm_vocab: Optional[Dict[str, object]] = None) -> Dict[str, object]:
    """Write a synthetic Llama-architecture GGUF.

    mode="quantize": fp32 master weights N(0, gain^2/fan_in) (embedding N(0,1), norm weights
                     1 + 0.1 N(0,1)) quantised with the valid-block quantisers above.  Use for
                     small models whose logits must be non-degenerate.
    mode="random":   raw random well-formed blocks (fast; for Llama-3-8B shaped benches).
    rope_freqs:   per-pair frequency factors written as rope_freqs.weight F32[head_dim/2] (what the Llama-3.1 converter emits).
    rope_scaling: (type, factor) -> {arch}.rope.scaling.type / .factor.   pre: tokenizer.ggml.pre.
    spm_vocab:    a SentencePiece vocabulary instead of the synthetic byte-level BPE one (tokenizer.ggml.model = "llama", the
                  Llama-2 / Mistral family): {"tokens": [...], "scores": [...], "types": [...], "bos": id, "eos": id, "unk": id,
                  "chat_template": str or None}; shape.n_vocab must equal len(tokens).
    Returns {"bytes": file size, "n_params": ..., "weights_bytes": matrix payload}.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    g = GGUFFile()
    g.add_str("general.architecture", arch)
    g.add_str("general.name", shape.name)
    g.add_u32("general.alignment", ALIGN)
    ftype = {"q4_k_m": 15, "q8_0": 7, "f16": 1, "bf16": 32, "f32": 0, "q4_k": 14, "q6_k": 18}[recipe]
    g.add_u32("general.file_type", ftype)
    g.add_u32("general.quantization_version", 2)
    g.add_u32(f"{arch}.block_count", shape.n_layer)
    g.add_u32(f"{arch}.context_length", shape.n_ctx)
    g.add_u32(f"{arch}.embedding_length", shape.n_embd)
    g.add_u32(f"{arch}.feed_forward_length", shape.n_ff)
    g.add_u32(f"{arch}.attention.head_count", shape.n_head)
    g.add_u32(f"{arch}.attention.head_count_kv", shape.n_head_kv)
    g.add_f32(f"{arch}.attention.layer_norm_rms_epsilon", shape.rms_eps)
    g.add_f32(f"{arch}.rope.freq_base", shape.rope_base)
    g.add_u32(f"{arch}.rope.dimension_count", shape.head_dim)
    g.add_u32(f"{arch}.vocab_size", shape.n_vocab)
    if rope_scaling is not None:
        g.add_str(f"{arch}.rope.scaling.type", rope_scaling[0])
        g.add_f32(f"{arch}.rope.scaling.factor", rope_scaling[1])
    if spm_vocab is not None:
        assert len(spm_vocab["tokens"]) == shape.n_vocab, (len(spm_vocab["tokens"]), shape.n_vocab)
        g.add_str("tokenizer.ggml.model", "llama")
        g.add_arr("tokenizer.ggml.tokens", _STR, list(spm_vocab["tokens"]))
        g.add_arr("tokenizer.ggml.scores", _F32, list(spm_vocab["scores"]))
        g.add_arr("tokenizer.ggml.token_type", _I32, list(spm_vocab["types"]))
        g.add_u32("tokenizer.ggml.bos_token_id", int(spm_vocab["bos"]))
        g.add_u32("tokenizer.ggml.eos_token_id", int(spm_vocab["eos"]))
        g.add_u32("tokenizer.ggml.unknown_token_id", int(spm_vocab["unk"]))
        g.add_bool("tokenizer.ggml.add_bos_token", True)
        g.add_bool("tokenizer.ggml.add_space_prefix", True)
        if spm_vocab.get("chat_template"):
            g.add_str("tokenizer.chat_template", str(spm_vocab["chat_template"]))
    elif with_vocab:
        toks, merges, types = synth_vocab(shape.n_vocab)
        g.add_str("tokenizer.ggml.model", "gpt2")
        g.add_str("tokenizer.ggml.pre", pre)
        g.add_arr("tokenizer.ggml.tokens", _STR, toks)
        g.add_arr("tokenizer.ggml.token_type", _I32, types)
        g.add_arr("tokenizer.ggml.merges", _STR, merges)
        g.add_u32("tokenizer.ggml.bos_token_id", shape.n_vocab - 3)
        g.add_u32("tokenizer.ggml.eos_token_id", shape.n_vocab - 2)
        g.add_u32("tokenizer.ggml.eot_token_id", shape.n_vocab - 1)
        g.add_bool("tokenizer.ggml.add_bos_token", True)
        g.add_str("tokenizer.chat_template",
                  "{% for m in messages %}<|{{ m.role }}|>\n{{ m.content }}<|eot_id|>{% endfor %}<|assistant|>\n")
    n_params = 0
    wbytes = 0
    if rope_freqs is not None:
        ff = np.ascontiguousarray(rope_freqs, dtype=np.float32).reshape(-1)
        assert ff.size == shape.head_dim // 2
        g.add_tensor("rope_freqs.weight", F32, (ff.size,), ff)
    for name, t, (rows, cols) in tensor_plan(shape, recipe):
        n_params += rows * cols
        if name.endswith("_norm.weight"):
            w = (1.0 + 0.1 * rng.standard_normal(cols)).astype(np.float32)
            g.add_tensor(name, F32, (cols,), w)
            continue
        if mode == "random":
            data = random_blocks(rng, t, rows, cols, gain=(np.sqrt(cols) if name == "token_embd.weight" else gain))
        else:
            std = 1.0 if name == "token_embd.weight" else gain / np.sqrt(cols)
            w = rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(std)
            data = quantize(w, t)
        wbytes += data.nbytes
        g.add_tensor(name, t, (rows, cols), data)
    size = g.write(path)
    return {"bytes": size, "n_params": n_params, "weights_bytes": wbytes}
