"""Profiling driver (run under ncu on the GPU box): a few batched decode steps of the Llama-3-8B q4_K_M bench model.
usage: batch_probe.py <batch> <ctx> <iters> <batch_weights 1|2>   -- prints ms per batched step (ignore the number under ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from gridllm_b200 import native as N
    path = bench.build_model_once("llama3_8b_q4km", 0, lambda: None)
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 576
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    e = N.Engine(path, max_ctx=1024, max_batch=max(2, batch), batch_weights=mode)
    ms, nl, wb = e.time_batch_step(batch, ctx, iters)
    print({"batch": batch, "ctx": ctx, "mode": mode, "ms_per_step": ms, "launches": nl, "weight_bytes": wb,
           "tok_s": batch / ms * 1e3}, flush=True)
    e.close()


if __name__ == "__main__":
    main()
