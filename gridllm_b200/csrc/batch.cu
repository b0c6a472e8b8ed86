// Per-sequence kernels of a batched decode step (batch.h): everything of the step that is NOT a weight GEMM.
// Stands in for what Ollama's runner does when it decodes several requests of one model together [external]; on the
// reference side only MAX_CONCURRENT_JOBS_PER_WORKER (server/src/config/index.ts:31) and the worker's busy-drop
// (client/src/services/WorkerClientService.ts:500-505) decide whether a worker ever sees more than one request.
#include "batch.h"

#include "attn_core.cuh"
#include "common.cuh"

namespace gl {

namespace {

__global__ void __launch_bounds__(MAX_BATCH) batch_gather_tokens_kernel(const BatchCtl* __restrict__ ctl, const StepState* __restrict__ st,
                                                                        int* __restrict__ ids, int bucket) {
    const int r = threadIdx.x;
    if (r >= bucket) return;
    ids[r] = r < ctl->n_rows ? st[ctl->row_slot[r]].token : 0;
}

// Same arithmetic as the EPI_QKV epilogue of the decode GEMV (fp32 rotate of adjacent pairs, fp16 cache), one CTA per row.
__global__ void __launch_bounds__(256) batch_rope_kv_kernel(const float* __restrict__ qkv, const BatchCtl* __restrict__ ctl,
                                                            const StepState* __restrict__ st, const int* __restrict__ tables, int table_stride,
                                                            int n_head, int n_kv, int hd, const float* __restrict__ cos_t,
                                                            const float* __restrict__ sin_t, float* __restrict__ q_out,
                                                            __half* __restrict__ k_cache, __half* __restrict__ v_cache) {
    const int r = blockIdx.x;
    if (r >= ctl->n_rows) return;
    const int slot = ctl->row_slot[r];
    const int pos = st[slot].pos;
    const int page = tables[(size_t)slot * table_stride + pos / KV_PAGE_TOKENS], tok = pos % KV_PAGE_TOKENS;
    const int qd = n_head * hd, kvd = n_kv * hd, ld = qd + 2 * kvd;
    const float* row = qkv + (size_t)r * ld;
    for (int i = threadIdx.x; i < (qd + kvd) / 2; i += 256) {
        const int e = 2 * i, d = e % hd;
        const float c = cos_t[(size_t)pos * (hd / 2) + d / 2], s = sin_t[(size_t)pos * (hd / 2) + d / 2];
        const float a = row[e], b = row[e + 1];
        const float o0 = a * c - b * s, o1 = a * s + b * c;
        if (e < qd) {
            *reinterpret_cast<float2*>(q_out + (size_t)r * qd + e) = make_float2(o0, o1);
        } else {
            const int ek = e - qd, kvh = ek / hd;
            const size_t off = (((size_t)page * n_kv + kvh) * KV_PAGE_TOKENS + tok) * hd + d;
            *reinterpret_cast<__half2*>(k_cache + off) = __halves2half2(__float2half_rn(o0), __float2half_rn(o1));
        }
    }
    for (int i = threadIdx.x; i < kvd / 2; i += 256) {
        const int e = 2 * i, kvh = e / hd, d = e % hd;
        const size_t off = (((size_t)page * n_kv + kvh) * KV_PAGE_TOKENS + tok) * hd + d;
        *reinterpret_cast<__half2*>(v_cache + off) =
            __halves2half2(__float2half_rn(row[qd + kvd + e]), __float2half_rn(row[qd + kvd + e + 1]));
    }
}

// Paged decode attention for B rows at once: grid (KV head, split, row); a CTA = the GQA group of one KV head of one row
// (one warp per query head, lanes own head dims), split s owns that row's pages s, s + S, ...; pages travel as 1-D TMA bulk
// copies (a page of one KV head is one contiguous 4 KB block), four pages per round trip; partials are merged by the last
// CTA of each (row, KV head) -- atomic ticket -- and the merged row is written as fp16, the operand of the attn_output GEMM.
// The page arithmetic is the single-sequence kernel's (attn_core.cuh): same scores, same online softmax.
constexpr int B_TILE_PAGES = 4;
constexpr int B_MAX_GRP = 8;

template <int DPL>
__global__ void __launch_bounds__(32 * B_MAX_GRP) batch_attn_kernel(const __grid_constant__ BatchAttnParams p) {
    constexpr int HD = DPL * 32;
    constexpr int PAGE_ELEMS = KV_PAGE_TOKENS * HD;
    constexpr uint32_t PAGE_BYTES = PAGE_ELEMS * sizeof(__half);
    __shared__ __align__(128) __half ks[B_TILE_PAGES * PAGE_ELEMS];
    __shared__ __align__(128) __half vs[B_TILE_PAGES * PAGE_ELEMS];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int is_last;

    const int row = blockIdx.z;
    if (row >= p.ctl->n_rows) return;
    const int slot = p.ctl->row_slot[row];
    const int* table = p.tables + (size_t)slot * p.table_stride;
    const int kvh = blockIdx.x, split = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = p.n_head / p.n_kv_heads;
    const int head = kvh * grp + warp;
    const int S = p.n_splits;
    const int L = p.st[slot].pos + 1;                 // this step's K / V row was appended by the launch before this one
    const int n_pages = (L + KV_PAGE_TOKENS - 1) / KV_PAGE_TOKENS;
    const int active = min(n_pages, S);
    if (split >= active) return;
    const int my_pages = (n_pages - split + S - 1) / S;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    float q[DPL], o[DPL];
    {
        const float* qp = p.q + ((size_t)row * p.n_head + head) * HD + lane * DPL;
        if (DPL == 4) {
            const float4 t = *reinterpret_cast<const float4*>(qp);
            q[0] = t.x; q[1] = t.y; q[DPL - 2] = t.z; q[DPL - 1] = t.w;
        } else {
            const float2 t = *reinterpret_cast<const float2*>(qp);
            q[0] = t.x; q[1] = t.y;
        }
#pragma unroll
        for (int d = 0; d < DPL; ++d) { q[d] *= p.scale; o[d] = 0.f; }
    }
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t ph = 0;
    for (int t0 = 0; t0 < my_pages; t0 += B_TILE_PAGES) {
        const int np = min(B_TILE_PAGES, my_pages - t0);
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(&bar, 2u * np * PAGE_BYTES);
            __syncwarp();
            if (lane < np) {
                const int page = table[split + (t0 + lane) * S];
                const size_t off = ((size_t)page * p.n_kv_heads + kvh) * PAGE_ELEMS;
                tma_load_1d(ks + lane * PAGE_ELEMS, p.k_cache + off, PAGE_BYTES, &bar);
                tma_load_1d(vs + lane * PAGE_ELEMS, p.v_cache + off, PAGE_BYTES, &bar);
            }
        }
        mbar_wait(&bar, ph);
        ph ^= 1;
        for (int i = 0; i < np; ++i) {
            const int pg = split + (t0 + i) * S;
            attn_page_math_smem<DPL>(ks + i * PAGE_ELEMS + lane * DPL, vs + i * PAGE_ELEMS + lane * DPL,
                                     min(KV_PAGE_TOKENS, L - pg * KV_PAGE_TOKENS), q, o, m_run, l_run);
        }
        // every thread has seen this phase complete (and is done with the tile) before the barrier is armed again
        if (t0 + B_TILE_PAGES < my_pages) __syncthreads();
    }

    __half* out = p.out16 + ((size_t)row * p.n_head + head) * HD + lane * DPL;
    if (active == 1) {
        const float inv = 1.0f / l_run;
#pragma unroll
        for (int d = 0; d < DPL; d += 2)
            *reinterpret_cast<__half2*>(out + d) = __halves2half2(__float2half_rn(o[d] * inv), __float2half_rn(o[d + 1] * inv));
        return;
    }
    const size_t pbase = ((size_t)row * p.n_head + head) * S;
    {
        float* po = p.part_o + (pbase + split) * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) po[d] = o[d];
        if (lane == 0) {
            p.part_ml[(pbase + split) * 2] = m_run;
            p.part_ml[(pbase + split) * 2 + 1] = l_run;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* ctr = p.counters + (size_t)row * p.n_kv_heads + kvh;
        unsigned ticket;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(ctr) : "memory");
        is_last = (ticket == (unsigned)active - 1);
        if (is_last) *ctr = 0;                         // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    // merge in split order (fixed order: the result does not depend on which CTA came last)
    float M = -INFINITY;
    for (int s2 = 0; s2 < active; ++s2) M = fmaxf(M, __ldcg(p.part_ml + (pbase + s2) * 2));
    float den = 0.f, acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[d] = 0.f;
    for (int s2 = 0; s2 < active; ++s2) {
        const float w = expf(__ldcg(p.part_ml + (pbase + s2) * 2) - M);
        den += w * __ldcg(p.part_ml + (pbase + s2) * 2 + 1);
        const float* po = p.part_o + (pbase + s2) * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) acc[d] += w * __ldcg(po + d);
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int d = 0; d < DPL; d += 2)
        *reinterpret_cast<__half2*>(out + d) = __halves2half2(__float2half_rn(acc[d] * inv), __float2half_rn(acc[d + 1] * inv));
}

// Greedy rows: one CTA per row scans the row's logits once (max / lowest argmax / sum of exponentials, the single-sequence
// sampler's arithmetic: ties go to the lowest index; logprob = -log(sum exp(l - max))) and advances the row's StepState.
constexpr int BS_THREADS = 1024;
struct BCand { float m; int i; float s; };
__device__ __forceinline__ BCand bcand_merge(BCand a, BCand b) {
    if (b.m > a.m || (b.m == a.m && b.i < a.i)) { const BCand t = a; a = b; b = t; }
    // a is the winner; fold b's mass in
    if (b.m != -INFINITY) a.s += b.s * expf(b.m - a.m);
    return a;
}
__global__ void __launch_bounds__(BS_THREADS) batch_sample_greedy_kernel(const float* __restrict__ logits, int n_vocab,
                                                                         const BatchCtl* __restrict__ ctl, StepState* __restrict__ stv,
                                                                         int* __restrict__ out_ids, float* __restrict__ out_lps, int max_out) {
    const int row = blockIdx.x;
    if (row >= ctl->n_rows) return;
    const int slot = ctl->row_slot[row];
    StepState* st = stv + slot;
    if (st->temperature > 0.f) return;               // sampled rows: the seeded top-k sampler runs on them after this kernel
    __shared__ BCand sc[BS_THREADS / 32];
    const float* lg = logits + (size_t)row * n_vocab;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    BCand c{-INFINITY, 0x7fffffff, 0.f};
    for (int i = tid; i < n_vocab; i += BS_THREADS) {
        const float v = lg[i];
        if (v > c.m) {                               // ascending indices per thread: strict > keeps the lowest on ties
            c.s = (c.m == -INFINITY ? 0.f : c.s * expf(c.m - v)) + 1.f;
            c.m = v;
            c.i = i;
        } else if (v != -INFINITY) {
            c.s += expf(v - c.m);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        BCand t{__shfl_xor_sync(0xffffffffu, c.m, o), __shfl_xor_sync(0xffffffffu, c.i, o), __shfl_xor_sync(0xffffffffu, c.s, o)};
        c = bcand_merge(c, t);
    }
    if (lane == 0) sc[warp] = c;
    __syncthreads();
    if (tid != 0) return;
    BCand t = sc[0];
    for (int w = 1; w < BS_THREADS / 32; ++w) t = bcand_merge(t, sc[w]);
    if (st->done) return;
    const int out_idx = st->out_idx;
    if (out_idx < max_out) {
        out_ids[(size_t)slot * max_out + out_idx] = t.i;
        out_lps[(size_t)slot * max_out + out_idx] = -logf(t.s);
    }
    st->token = t.i;
    st->pos = st->pos + 1;
    st->out_idx = out_idx + 1;
    if (!st->ignore_eos) {
        for (int k = 0; k < st->n_stop; ++k)
            if (st->stop_ids[k] == t.i) st->done = 1;
    }
}

__global__ void __launch_bounds__(MAX_BATCH) batch_collect_kernel(const BatchCtl* __restrict__ ctl, const StepState* __restrict__ st,
                                                                  const float* __restrict__ out_lps, int max_out, BatchOut* __restrict__ out,
                                                                  int bucket) {
    const int r = threadIdx.x;
    if (r >= bucket) return;
    BatchOut o{0, 0.f, 0, 0};
    if (r < ctl->n_rows) {
        const int slot = ctl->row_slot[r];
        const StepState& s = st[slot];
        o.token = s.token;
        o.done = s.done;
        o.pos = s.pos;
        const int k = s.out_idx - 1;
        o.logprob = (k >= 0 && k < max_out) ? out_lps[(size_t)slot * max_out + k] : 0.f;
    }
    out[r] = o;
}

__global__ void __launch_bounds__(256) batch_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, int n, float eps,
                                                            __half* __restrict__ y) {
    const int r = blockIdx.x;
    __shared__ float red[8];
    const float* xr = x + (size_t)r * n;
    __half* yr = y + (size_t)r * n;
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) { const float v = xr[i]; ss += v * v; }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int k = 0; k < 8; ++k) tot += red[k];
    const float rstd = 1.0f / sqrtf(tot / (float)n + eps);
    for (int i = threadIdx.x; i < n; i += 256) yr[i] = __float2half_rn((xr[i] * rstd) * w[i]);
}

}  // namespace

cudaError_t batch_gather_tokens_launch(const BatchCtl* ctl, const StepState* st, int* ids, int bucket, cudaStream_t s) {
    batch_gather_tokens_kernel<<<1, MAX_BATCH, 0, s>>>(ctl, st, ids, bucket);
    return cudaGetLastError();
}

cudaError_t batch_rope_kv_launch(const float* qkv, int bucket, const BatchCtl* ctl, const StepState* st, const int* tables, int table_stride,
                                 int n_head, int n_kv, int hd, const float* cos_t, const float* sin_t, float* q_out, __half* k_cache,
                                 __half* v_cache, cudaStream_t s) {
    batch_rope_kv_kernel<<<bucket, 256, 0, s>>>(qkv, ctl, st, tables, table_stride, n_head, n_kv, hd, cos_t, sin_t, q_out, k_cache, v_cache);
    return cudaGetLastError();
}

cudaError_t batch_attn_launch(const BatchAttnParams& p, int bucket, cudaStream_t s) {
    const int grp = p.n_head / p.n_kv_heads;
    if (grp < 1 || grp > B_MAX_GRP || p.n_head % p.n_kv_heads || p.n_splits < 1 || p.n_splits > 32) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)p.n_kv_heads, (unsigned)p.n_splits, (unsigned)bucket);
    if (p.head_dim == 128) batch_attn_kernel<4><<<grid, 32 * grp, 0, s>>>(p);
    else if (p.head_dim == 64) batch_attn_kernel<2><<<grid, 32 * grp, 0, s>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t batch_sample_greedy_launch(const float* logits, int n_vocab, int bucket, const BatchCtl* ctl, StepState* st, int* out_ids,
                                       float* out_lps, int max_out, cudaStream_t s) {
    batch_sample_greedy_kernel<<<bucket, BS_THREADS, 0, s>>>(logits, n_vocab, ctl, st, out_ids, out_lps, max_out);
    return cudaGetLastError();
}

cudaError_t batch_collect_launch(const BatchCtl* ctl, const StepState* st, const float* out_lps, int max_out, BatchOut* out, int bucket,
                                 cudaStream_t s) {
    batch_collect_kernel<<<1, MAX_BATCH, 0, s>>>(ctl, st, out_lps, max_out, out, bucket);
    return cudaGetLastError();
}

cudaError_t batch_rmsnorm_launch(const float* x, const float* w, int rows, int n, float eps, __half* y, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    batch_rmsnorm_kernel<<<rows, 256, 0, s>>>(x, w, n, eps, y);
    return cudaGetLastError();
}

}  // namespace gl
