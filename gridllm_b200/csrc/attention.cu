// Paged-KV decode attention (one query token, GQA), split over KV pages -- stand-alone kernel of the per-op path.
//
// Stands in for the attention inside Ollama's decode step, reached in the reference only through
// OllamaService.generate*Response (/root/reference/client/src/services/OllamaService.ts:142-145, 235-237).
// Bound: HBM (KV pages) in principle, latency in practice at the 512+128-token workloads of BASELINE.json (2.6 MB
// of KV per layer), so everything is arranged to shorten the dependent chain: one launch; all pages of a split
// (<= 4) staged at once with 1-D TMA bulk copies (a 16-token page of one KV head is one contiguous block); 16
// independent dot/shuffle chains per page; partials merged by the last CTA of each KV head (atomic ticket) with
// batched loads -- no second kernel.  The grid is fixed (it lives in a CUDA graph) but the number of splits that take
// part is chosen from the context length at run time -- one page per split at least: surplus CTAs leave at once,
// and a context of one page is written straight to the output without partials, ticket or merge.
#include "attn_core.cuh"

namespace gl {

namespace {

constexpr int TILE_PAGES = 4;
constexpr int MAX_GRP = 8;

template <int DPL>   // dims per lane = head_dim / 32
__global__ void __launch_bounds__(32 * MAX_GRP) attn_decode_kernel(const __grid_constant__ AttnParams p) {
    constexpr int HD = DPL * 32;
    constexpr int PAGE_ELEMS = KV_PAGE_TOKENS * HD;
    __shared__ __align__(128) __half ks[TILE_PAGES * PAGE_ELEMS];
    __shared__ __align__(128) __half vs[TILE_PAGES * PAGE_ELEMS];
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
    const int active = min(n_pages, p.n_splits);
    if (split >= active) return;
    const int pg0 = (split * n_pages) / active, pg1 = ((split + 1) * n_pages) / active;

    float q[DPL], o[DPL];
    {
        const float* qp = p.q + (size_t)head * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) { q[d] = __ldcg(qp + d) * p.scale; o[d] = 0.f; }
    }
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t ph = 0;
    for (int t0 = pg0; t0 < pg1; t0 += TILE_PAGES) {
        const int np = min(TILE_PAGES, pg1 - t0);
        if (threadIdx.x == 0) {
            const uint32_t page_bytes = PAGE_ELEMS * sizeof(__half);
            mbar_expect_tx(&bar, 2u * np * page_bytes);
            for (int i = 0; i < np; ++i) {
                const int page = __ldcg(p.page_table + t0 + i);
                const size_t off = ((size_t)page * p.n_kv_heads + kvh) * PAGE_ELEMS;
                tma_load_1d(ks + i * PAGE_ELEMS, p.k_cache + off, page_bytes, &bar);
                tma_load_1d(vs + i * PAGE_ELEMS, p.v_cache + off, page_bytes, &bar);
            }
        }
        mbar_wait(&bar, ph);
        ph ^= 1;
        for (int i = 0; i < np; ++i) {
            uint2 kk[KV_PAGE_TOKENS], vv[KV_PAGE_TOKENS];
            const __half* kb = ks + i * PAGE_ELEMS + lane * DPL;
            const __half* vb = vs + i * PAGE_ELEMS + lane * DPL;
#pragma unroll
            for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
                if (DPL == 4) {
                    kk[j] = *reinterpret_cast<const uint2*>(kb + j * HD);
                    vv[j] = *reinterpret_cast<const uint2*>(vb + j * HD);
                } else {
                    kk[j] = make_uint2(*reinterpret_cast<const unsigned*>(kb + j * HD), 0u);
                    vv[j] = make_uint2(*reinterpret_cast<const unsigned*>(vb + j * HD), 0u);
                }
            }
            attn_page_math<DPL>(kk, vv, min(KV_PAGE_TOKENS, L - (t0 + i) * KV_PAGE_TOKENS), q, o, m_run, l_run);
        }
        if (t0 + TILE_PAGES < pg1) __syncthreads();   // tile buffers are re-filled by the next TMA
    }

    if (active == 1) {
        const float inv = 1.0f / l_run;
        float* out = p.out + (size_t)head * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) out[d] = o[d] * inv;
        return;
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
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ticket;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(p.counters + kvh) : "memory");
        is_last = (ticket == (unsigned)active - 1);
        if (is_last) p.counters[kvh] = 0;      // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    attn_merge_head<DPL>(p.part_o, p.part_ml, p.out, head, p.n_splits, active, lane);
}

}  // namespace

cudaError_t attn_decode_launch(const AttnParams& p, bool pdl, cudaStream_t s) {
    const int grp = p.n_head / p.n_kv_heads;
    if (grp < 1 || grp > MAX_GRP || p.n_head % p.n_kv_heads || p.n_splits < 1 || p.n_splits > 32) return cudaErrorInvalidValue;
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
