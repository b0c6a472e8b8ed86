"""Generates the committed golden vectors.  The reference holds none for this path, so
these are produced by the pinned oracle + the independent gguf-py dequantisers:

  dequant_<type>.npz   : seeded random blocks -> gguf.quants.dequantize output            (block-format KAT)
  gemv_<type>.npz      : blocks + x -> W_deq @ x in float64 (exact) and with x snapped to the engine's fixed point
  tiny_model_logits.npz: the 2-layer d=256 synthetic q4_K_M model (oracle/gguf_synth.py TINY, seed 1234):
                         logits of 16 prompt positions + 8 greedy steps, ids, logprobs, margins (exact mode, fp16 KV)
  sampler_draws.npz    : two 512-entry logit vectors (smooth / heavily tied) x 24 settings of (temperature, top_k, top_p,
                         seed, output index) -> token id, logprob, distance of the draw from the nearest CDF boundary
                         (oracle/sampler.py; the uniform generator is pinned to splitmix64's published outputs)

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gguf_synth as S, llama_oracle as O, sampler as SM  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    from gguf import quants, GGMLQuantizationType as T
    rng = np.random.Generator(np.random.PCG64(1234))
    for name, t, gt in (("q4_k", S.Q4_K, T.Q4_K), ("q6_k", S.Q6_K, T.Q6_K), ("q8_0", S.Q8_0, T.Q8_0)):
        blocks = S.random_blocks(rng, t, 4, 512)
        np.savez_compressed(os.path.join(OUT, f"dequant_{name}.npz"), blocks=blocks, dequant=quants.dequantize(blocks, gt).astype(np.float32))
        rows, cols = 24, 1024
        b2 = S.random_blocks(rng, t, rows, cols)
        x = np.random.Generator(np.random.PCG64(42)).standard_normal(cols).astype(np.float32)
        wd = O.dequantize(b2, t, (rows, cols))
        np.savez_compressed(os.path.join(OUT, f"gemv_{name}.npz"), blocks=b2, x=x, rows=rows, cols=cols,
                            y_exact=O.gemv(wd, x, "exact"), y_i16=O.gemv(wd, x, "i16"), y_q8=O.gemv(wd, x, "q8"))
    path = os.path.join(OUT, "_tiny_tmp.gguf")
    S.build_model(path, S.TINY, "q4_k_m", seed=1234)
    m = O.load_gguf(path)
    orc = O.LlamaOracle(m, act="exact", kv_f16=True)
    prompt = np.random.Generator(np.random.PCG64(1000)).integers(0, m.n_vocab - 3, size=16)
    prompt_logits = []
    orc.reset()
    for t in prompt:
        prompt_logits.append(orc.step(int(t)).astype(np.float32))
    g = orc.generate(prompt, 8)
    np.savez_compressed(os.path.join(OUT, "tiny_model_logits.npz"), prompt=prompt.astype(np.int32),
                        prompt_logits=np.stack(prompt_logits), gen_ids=g["ids"], gen_logprobs=g["logprobs"],
                        gen_margins=g["margins"], gen_logits=g["logits"], file_sha_hint=np.array([os.path.getsize(path)]))
    os.remove(path)
    # sampler draws
    srng = np.random.Generator(np.random.PCG64(77))
    logits = np.stack([(srng.standard_normal(512) * 3.0).astype(np.float32), srng.integers(-3, 4, size=512).astype(np.float32)])
    settings, results = [], []
    for v in range(2):
        for _ in range(24):
            t = float(srng.choice([0.2, 0.8, 1.0, 1.5]))
            k = int(srng.choice([1, 5, 40, 64, 100, 0]))
            p = float(srng.choice([1.0, 0.9, 0.5]))
            seed, idx = int(srng.integers(0, 2 ** 62)), int(srng.integers(0, 32))
            tok, lp, margin = SM.sample(logits[v], t, k, p, seed, idx)
            settings.append((v, t, k, p, seed, idx))
            results.append((tok, lp, margin))
    np.savez_compressed(os.path.join(OUT, "sampler_draws.npz"), logits=logits,
                        vector=np.array([s[0] for s in settings], np.int32), temperature=np.array([s[1] for s in settings], np.float32),
                        top_k=np.array([s[2] for s in settings], np.int32), top_p=np.array([s[3] for s in settings], np.float32),
                        seed=np.array([s[4] for s in settings], np.int64), out_index=np.array([s[5] for s in settings], np.int32),
                        token=np.array([r[0] for r in results], np.int32), logprob=np.array([r[1] for r in results], np.float64),
                        margin=np.array([r[2] for r in results], np.float64))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
