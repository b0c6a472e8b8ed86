"""Quick end-to-end probe on the GPU box: build the Llama-3-8B-shaped synthetic q4_K_M GGUF, load it,
time decode steps and one 512-in/128-out request."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gguf_synth as S  # noqa: E402


def main():
    from gridllm_b200 import native as N
    path = "/dev/shm/llama3_8b_q4km_synth.gguf" if os.path.isdir("/dev/shm") else "/tmp/llama3_8b_q4km_synth.gguf"
    t0 = time.time()
    info = S.build_model(path, S.LLAMA3_8B, "q4_k_m", seed=1234, mode="random", with_vocab=False)
    print("build_s", round(time.time() - t0, 1), info, flush=True)
    cfgs = ({}, {"GL_ACT_BITS": "8"})
    for cfg in cfgs:
        for k in ("GL_ACT_BITS", "GL_PDL", "GL_GRAPH", "GL_MEGA", "GL_MEGA_SLOT_BYTES", "GL_MEGA_SLOTS", "GL_MEGA_INFLIGHT", "GL_ATTN_SPLITS", "GL_WARPS", "GL_RING_DEPTH", "GL_MEGA_TRACKS", "GL_LEAN_RINGS", "GL_POLITE_TRACKS", "GL_GREEDY_PDL", "GL_XRAW", "GL_HB256", "GL_XRAW_WIDE", "GL_SMEM_KB"):
            os.environ.pop(k, None)
        os.environ.update(cfg)
        t0 = time.time()
        os.environ["GL_PREFILL"] = "1"
        gen = False
        e = N.Engine(path, max_ctx=2048)
        load_s = time.time() - t0
        bpt = e.info.decode_bytes_per_token
        for ctx in (1, 576):
            ms, nl = e.time_decode(ctx, 32)
            print(json.dumps({"cfg": cfg, "ctx": ctx, "ms_per_token": round(ms, 4), "tok_s": round(1000 / ms, 1), "launches": nl,
                              "weights_GBps": round(bpt / ms / 1e6, 1), "load_s": round(load_s, 1)}), flush=True)
        if not gen:
            e.close()
            continue
        prompt = np.random.Generator(np.random.PCG64(1000)).integers(0, 128000, size=512)
        g = e.generate(prompt, num_predict=128, ignore_eos=True)
        st = g.stats
        print(json.dumps({"cfg": cfg, "gen": int(st.eval_count), "prefill_ms": st.prompt_eval_duration_ns / 1e6,
                          "decode_ms": st.eval_duration_ns / 1e6, "decode_tok_s": round(st.eval_count / (st.eval_duration_ns / 1e9), 1),
                          "ids_head": g.ids[:8].tolist(), "finite_lp": bool(np.isfinite(g.logprobs).all())}), flush=True)
        e.close()
    os.remove(path)


if __name__ == "__main__":
    main()
