"""CPU restatement of the seeded temperature / top-k / top-p draw (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
CPU legs may import this; the product path is gridllm_b200/csrc/sampler.cu).

Follows the sampling options the reference forwards to Ollama (InferenceRequest.options.temperature / top_k / top_p / seed,
/root/reference/client/src/types/index.ts:1-27; OllamaService.generateResponse, /root/reference/client/src/services/
OllamaService.ts:101-134).  The arithmetic itself lives in Ollama [external, not under /root/reference, unpinned image]; the
published order of its sampler is restated here: top-k -> temperature -> softmax -> top-p -> draw by inverse CDF.  PARITY
UNPINNED against a real Ollama: its random generator is not reproducible from outside, so parity is defined on the
distribution (the kept candidates and their cumulative masses) and on this repo's own counter-based generator.

Definitions shared with the kernel:
  * candidates: the k best logits ordered by (logit descending, index ascending); k = top_k, or MAX_K when top_k is "off"
    (<= 0) or larger than MAX_K;
  * weights w_j = exp((l_j - l_0) / T); running sum c_j in candidate order;
  * top-p keeps the shortest prefix with c_j >= top_p * c_last (off for top_p <= 0 or >= 1);
  * u = (splitmix64(seed + GOLDEN * (out_index + 1)) >> 40) / 2**24, r = u * c_keep; the draw is the first j with c_j > r;
  * reported logprob = log-softmax of the drawn logit over the WHOLE vocabulary at T = 1 (what the greedy sampler reports).
"""
from __future__ import annotations

import numpy as np

MAX_K = 1024
_M64 = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15


def uniform24(seed: int, out_index: int) -> float:
    z = (seed + _GOLDEN * (out_index + 1)) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return float(z >> 40) / 16777216.0


def candidates(logits: np.ndarray, top_k: int) -> np.ndarray:
    """Indices of the candidates in draw order."""
    n = len(logits)
    k = top_k if 0 < top_k <= MAX_K else MAX_K
    k = min(k, n)
    order = np.lexsort((np.arange(n), -logits.astype(np.float64)))      # logit descending, index ascending
    return order[:k]


def distribution(logits: np.ndarray, temperature: float, top_k: int, top_p: float):
    """(candidate ids kept, cumulative masses of the kept candidates) in float64."""
    l32 = np.asarray(logits, dtype=np.float32)
    ids = candidates(l32, top_k)
    inv_t = np.float32(1.0) / np.float32(temperature)                  # the kernel's fp32 reciprocal
    w = np.exp((l32[ids].astype(np.float64) - float(l32[ids[0]])) * float(inv_t))
    c = np.cumsum(w)
    keep = len(ids)
    if 0.0 < top_p < 1.0:
        keep = int(np.argmax(c >= np.float32(top_p).astype(np.float64) * c[-1])) + 1
    return ids[:keep], c[:keep]


def sample(logits: np.ndarray, temperature: float, top_k: int = 0, top_p: float = 1.0, seed: int = 0, out_index: int = 0):
    """-> (token id, logprob, margin): margin = distance of the draw from the nearest CDF boundary, relative to the kept mass
    (a kernel working in fp32 may legitimately land on the neighbour when margin is ~1e-6)."""
    l32 = np.asarray(logits, dtype=np.float32)
    if temperature <= 0:
        i = int(np.argmax(l32))       # first maximum = lowest index
        m = float(l32[i])
        return i, -float(np.log(np.sum(np.exp(l32.astype(np.float64) - m)))), 1.0
    ids, c = distribution(l32, temperature, top_k, top_p)
    u = uniform24(seed, out_index)
    r = u * c[-1]
    j = int(np.searchsorted(c, r, side="right"))
    j = min(j, len(ids) - 1)
    lo = c[j - 1] if j > 0 else 0.0
    margin = min(r - lo, c[j] - r) / c[-1]
    m = float(l32.max())
    lse = m + float(np.log(np.sum(np.exp(l32.astype(np.float64) - m))))
    return int(ids[j]), float(l32[ids[j]]) - lse, float(margin)


def interval_error(logits: np.ndarray, token: int, temperature: float, top_k: int = 0, top_p: float = 1.0, seed: int = 0,
                   out_index: int = 0) -> float:
    """How far the draw u * mass lies OUTSIDE the cumulative interval of `token`, relative to the kept mass (0 = inside,
    inf = token is not among the kept candidates).  The statement a fp32 implementation is held to."""
    ids, c = distribution(np.asarray(logits, dtype=np.float32), temperature, top_k, top_p)
    where = np.nonzero(ids == token)[0]
    if len(where) == 0:
        return float("inf")
    j = int(where[0])
    r = uniform24(seed, out_index) * c[-1]
    lo = c[j - 1] if j > 0 else 0.0
    return float(max(lo - r, r - c[j], 0.0) / c[-1])
