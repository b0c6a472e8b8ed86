// Paged-KV decode attention (one query token, GQA), split over KV pages.
//
// Stands in for the attention inside Ollama's decode step, reached in the reference only through
// OllamaService.generate*Response (/root/reference/client/src/services/OllamaService.ts:142-145,
// 235-237).  Bound: HBM (KV pages) in principle, launch/latency in practice at the 512+128-token
// workloads of BASELINE.json -- so the kernel is a single launch: split-KV partials are merged by
// the last CTA of each KV head (atomic ticket), no second kernel.
//
// KV pages ([page][kv_head][16 tokens][head_dim] fp16) are staged into shared memory with 1-D TMA
// bulk copies (one 16-token page of one KV head is one contiguous 16*head_dim*2-byte block).
#include "common.cuh"
#include "kernels.h"

namespace gl {

namespace {

constexpr int TILE_PAGES = 4;
constexpr int TILE_POS = TILE_PAGES * KV_PAGE_TOKENS;   // 64
constexpr int MAX_GRP = 8;

template <int DPL>   // dims per lane = head_dim / 32
__global__ void __launch_bounds__(32 * MAX_GRP) attn_decode_kernel(const __grid_constant__ AttnParams p) {
    constexpr int HD = DPL * 32;
    __shared__ __align__(128) __half ks[TILE_POS * HD];
    __shared__ __align__(128) __half vs[TILE_POS * HD];
    __shared__ float probs[MAX_GRP][TILE_POS];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int is_last;

    const int kvh = blockIdx.x, split = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = p.n_head / p.n_kv_heads;
    const int head = kvh * grp + warp;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();

    const int L = __ldcg(&p.st->pos) + 1;
    const int n_pages = (L + KV_PAGE_TOKENS - 1) / KV_PAGE_TOKENS;
    const int pps = (n_pages + p.n_splits - 1) / p.n_splits;
    const int pg0 = split * pps, pg1 = min(n_pages, pg0 + pps);

    float q[DPL];
    {
        const float* qp = p.q + (size_t)head * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) q[d] = __ldcg(qp + d) * p.scale;
    }
    float m_run = -INFINITY, l_run = 0.f;
    float o[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[d] = 0.f;

    uint32_t ph = 0;
    for (int t0 = pg0; t0 < pg1; t0 += TILE_PAGES) {
        const int np = min(TILE_PAGES, pg1 - t0);
        if (threadIdx.x == 0) {
            const uint32_t page_bytes = KV_PAGE_TOKENS * HD * sizeof(__half);
            mbar_expect_tx(&bar, 2u * np * page_bytes);
            for (int i = 0; i < np; ++i) {
                const int page = __ldcg(p.page_table + t0 + i);
                const size_t off = ((size_t)page * p.n_kv_heads + kvh) * KV_PAGE_TOKENS * HD;
                tma_load_1d(ks + i * KV_PAGE_TOKENS * HD, p.k_cache + off, page_bytes, &bar);
                tma_load_1d(vs + i * KV_PAGE_TOKENS * HD, p.v_cache + off, page_bytes, &bar);
            }
        }
        mbar_wait(&bar, ph);
        ph ^= 1;
        const int pos0 = t0 * KV_PAGE_TOKENS;
        const int npos = min(np * KV_PAGE_TOKENS, L - pos0);

        // scores of this head against the tile
        float m_t = -INFINITY;
        for (int j = 0; j < npos; ++j) {
            const __half* kr = ks + j * HD + lane * DPL;
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < DPL; d += 2) {
                const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(kr + d));
                a += q[d] * kk.x + q[d + 1] * kk.y;
            }
            a = warp_sum(a);
            if (lane == 0) probs[warp][j] = a;
            m_t = fmaxf(m_t, a);
        }
        __syncwarp();
        const float m_new = fmaxf(m_run, m_t);
        const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
        float lsum = 0.f;
        for (int j = lane; j < npos; j += 32) {
            const float w = expf(probs[warp][j] - m_new);
            probs[warp][j] = w;
            lsum += w;
        }
        lsum = warp_sum(lsum);
        __syncwarp();
        l_run = l_run * corr + lsum;
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] *= corr;
        for (int j = 0; j < npos; ++j) {
            const float w = probs[warp][j];
            const __half* vr = vs + j * HD + lane * DPL;
#pragma unroll
            for (int d = 0; d < DPL; d += 2) {
                const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(vr + d));
                o[d] += w * vv.x;
                o[d + 1] += w * vv.y;
            }
        }
        m_run = m_new;
        __syncthreads();   // tile buffers are re-filled by the next TMA
    }

    // partial result of (head, split)
    {
        float* po = p.part_o + ((size_t)head * p.n_splits + split) * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) po[d] = o[d];
        if (lane == 0) {
            p.part_ml[((size_t)head * p.n_splits + split) * 2] = m_run;
            p.part_ml[((size_t)head * p.n_splits + split) * 2 + 1] = l_run;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(p.counters + kvh, 1u);
        is_last = (ticket == (unsigned)p.n_splits - 1);
        if (is_last) p.counters[kvh] = 0;      // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // merge all splits of this head
    float M = -INFINITY;
    for (int s = 0; s < p.n_splits; ++s) M = fmaxf(M, __ldcg(p.part_ml + ((size_t)head * p.n_splits + s) * 2));
    float den = 0.f;
    float acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[d] = 0.f;
    for (int s = 0; s < p.n_splits; ++s) {
        const float ms = __ldcg(p.part_ml + ((size_t)head * p.n_splits + s) * 2);
        const float ls = __ldcg(p.part_ml + ((size_t)head * p.n_splits + s) * 2 + 1);
        if (ms == -INFINITY) continue;
        const float w = expf(ms - M);
        den += w * ls;
        const float* po = p.part_o + ((size_t)head * p.n_splits + s) * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) acc[d] += w * __ldcg(po + d);
    }
    const float inv = 1.0f / den;
    float* out = p.out + (size_t)head * HD + lane * DPL;
#pragma unroll
    for (int d = 0; d < DPL; ++d) out[d] = acc[d] * inv;
}

}  // namespace

cudaError_t attn_decode_launch(const AttnParams& p, bool pdl, cudaStream_t s) {
    const int grp = p.n_head / p.n_kv_heads;
    if (grp < 1 || grp > MAX_GRP || p.n_head % p.n_kv_heads) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)p.n_kv_heads, (unsigned)p.n_splits);
    cfg.blockDim = dim3(32u * grp);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    if (p.head_dim == 128) return cudaLaunchKernelEx(&cfg, attn_decode_kernel<4>, p);
    if (p.head_dim == 64) return cudaLaunchKernelEx(&cfg, attn_decode_kernel<2>, p);
    return cudaErrorInvalidValue;
}

}  // namespace gl
