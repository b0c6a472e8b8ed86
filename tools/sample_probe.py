"""Cost of the seeded top-k / top-p sampler inside a request: 512-in / 128-out on the Llama-3-8B-shaped synthetic model,
greedy against temperature 0.8 / top_k 40 / top_p 0.9 (and top_k off = 1024 candidates).  Device time of the decode part."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from gridllm_b200 import native as N
    path = "/dev/shm/prof_llama3_8b.gguf"
    if not os.path.exists(path):
        from oracle import gguf_synth as S
        S.build_model(path, S.LLAMA3_8B, "q4_k_m", seed=1234, mode="random", with_vocab=False)
    pdl = os.environ.get("GL_PDL", "1")
    e = N.Engine(path, max_ctx=2048)
    prompt = np.random.Generator(np.random.PCG64(1000)).integers(0, 128000, size=512)
    for name, kw in (("greedy", {}), ("t0.8_k40_p0.9", dict(temperature=0.8, top_k=40, top_p=0.9, seed=1)),
                     ("t0.8_k1024", dict(temperature=0.8, top_k=0, top_p=1.0, seed=1))):
        best = None
        for _ in range(3):
            g = e.generate(prompt, num_predict=128, ignore_eos=True, **kw)
            ms = g.stats.eval_duration_ns / 1e6 / g.stats.eval_count
            best = ms if best is None else min(best, ms)
        print(json.dumps({"pdl": pdl, "sampler": name, "decode_ms_per_token": round(best, 4), "distinct_tokens": int(len(set(g.ids.tolist())))}), flush=True)
    e.close()


if __name__ == "__main__":
    main()
