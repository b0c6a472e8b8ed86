// Persistent decode kernel: the WHOLE decode step (and several steps back to back) in one cooperative
// launch, one CTA per SM.  Replaces the 163 dependent launches of the per-op path: at ~0.7 ms of HBM
// traffic per token, launch + ramp + drain of every small kernel was costing more than the traffic.
//
//   producer warp : walks the phase table of the token (QKV, O, gate/up, down of every layer, lm_head) and
//                   streams this CTA's weight rows into the shared-memory ring with 1-D TMA bulk copies.  It
//                   never waits for a grid barrier -- weights do not depend on activations -- so while the
//                   consumers synchronise, the ring fills with the NEXT phase's rows and HBM stays busy.
//   consumer warps: phase by phase: fused prologue (RMSNorm + activation snap), dp4a block decode from the
//                   ring, fused epilogue; split-KV paged attention; greedy sampling; a grid-wide barrier
//                   (one atomic + acquire spin per CTA) between phases.
//
// Reference call site replaced: the decode loop inside Ollama behind OllamaService.generate*Response
// (/root/reference/client/src/services/OllamaService.ts:142-145, 235-237).
#include "decode_mega.h"
#include "attn_core.cuh"
#include "gemv_core.cuh"

namespace gl {

namespace {

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all consumer threads of every CTA call this the same number of times
template <int NT>
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target, int tid) {
    // bar.sync orders every consumer thread's writes before thread 0's release; the acquire load + bar.sync hand the
    // other CTAs' writes to every thread of this one.  No full fences.
    named_bar_sync(1, NT);
    if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned spins = 0;
        while ((int)(ld_acquire_u32(counter) - target) < 0) {
            if (++spins > (1u << 28)) __trap();
        }
    }
    named_bar_sync(1, NT);
}

__device__ __forceinline__ float dequant_native_elem(const uint8_t* row, int type, int c) {
    switch (type) {
        case T_F32: return reinterpret_cast<const float*>(row)[c];
        case T_F16: return __half2float(reinterpret_cast<const __half*>(row)[c]);
        case T_BF16: return __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(row)[c] << 16);
        case T_Q8_0: {
            const uint8_t* b = row + (size_t)(c >> 5) * 34;
            return half_bits_to_float(*reinterpret_cast<const uint16_t*>(b)) * (float)(int8_t)b[2 + (c & 31)];
        }
        case T_Q4_K: return dequant_native_q4k(row, c);
        case T_Q6_K: {
            const uint8_t* b = row + (size_t)(c >> 8) * 210;
            const int e = c & 255, h = e >> 7, r = e & 127;
            const int qlv = (b[h * 64 + (r & 63)] >> (4 * (r >> 6))) & 0xF;
            const int qhv = (b[128 + h * 32 + (r & 31)] >> (2 * (r >> 5))) & 3;
            const float d = half_bits_to_float(*reinterpret_cast<const uint16_t*>(b + 208));
            return d * (float)(int8_t)b[192 + (e >> 4)] * (float)((qlv | (qhv << 4)) - 32);
        }
        default: return 0.f;
    }
}

// ---- split-KV paged attention for one (kv head, split) item; K/V rows read straight from HBM/L2 ----------
// One warp per query head of the GQA group; a lane owns DPL dims.
// The phase is a chain of dependent L2 round trips, so it is cut in two around the grid barrier that follows the
// QKV phase: K/V rows of OLD positions are final before that barrier, and attn_prefetch() requests up to two
// 16-token pages of them (64 independent 8-B loads per lane) before the CTA waits; after the barrier only q and
// the single row of the newest position remain to be fetched.
constexpr int ATTN_PRE_PAGES = 2;
constexpr int ATTN_SMEM_BYTES = 2 * ATTN_PRE_PAGES * KV_PAGE_TOKENS * 128 * 2 + 64;   // K + V tiles (head_dim <= 128) + mbarrier

// 16 rows of one page of K (or V) for this lane's dims, straight from global memory
template <int DPL>
__device__ __forceinline__ void attn_load_rows(const MegaParams& mp, const __half* cache, int kvh, int pg, int L, int lane, uint2* rows) {
    constexpr int HD = DPL * 32;
    const int page = __ldcg(mp.page_table + pg);
    const size_t base = ((size_t)page * mp.n_kv + kvh) * KV_PAGE_TOKENS * HD + lane * DPL;
    const int npos = min(KV_PAGE_TOKENS, L - pg * KV_PAGE_TOKENS);
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        rows[j] = make_uint2(0u, 0u);
        if (j < npos) {
            if (DPL == 4) rows[j] = __ldcg(reinterpret_cast<const uint2*>(cache + base + (size_t)j * HD));
            else rows[j].x = __ldcg(reinterpret_cast<const unsigned*>(cache + base + (size_t)j * HD));
        }
    }
}

// one thread: TMA-stage up to two KV pages of this CTA's attention item into shared memory (old positions are final)
__device__ __forceinline__ void attn_prefetch(uint8_t* abuf, const MegaParams& mp, const __half* kc, const __half* vc, int item, int pos) {
    const int HD = mp.head_dim;
    const int n_splits = mp.attn_splits;
    const int kvh = item / n_splits, split = item % n_splits;
    const int L = pos + 1;
    const int n_pages = (L + KV_PAGE_TOKENS - 1) / KV_PAGE_TOKENS;
    const int pps = (n_pages + n_splits - 1) / n_splits;
    const int pg0 = split * pps, pg1 = min(n_pages, pg0 + pps);
    const int np = min(ATTN_PRE_PAGES, pg1 - pg0);
    if (np <= 0) return;
    uint64_t* bar = reinterpret_cast<uint64_t*>(abuf + ATTN_SMEM_BYTES - 64);
    const uint32_t page_bytes = (uint32_t)(KV_PAGE_TOKENS * HD * sizeof(__half));
    mbar_expect_tx(bar, 2u * np * page_bytes);
    for (int i = 0; i < np; ++i) {
        const int page = __ldcg(mp.page_table + pg0 + i);
        const size_t off = ((size_t)page * mp.n_kv + kvh) * KV_PAGE_TOKENS * HD;
        tma_load_1d(abuf + (size_t)i * page_bytes, kc + off, page_bytes, bar);
        tma_load_1d(abuf + (size_t)(ATTN_PRE_PAGES + i) * page_bytes, vc + off, page_bytes, bar);
    }
}

template <int DPL, int NT>
__device__ __forceinline__ void attn_item(uint8_t* abuf, bool prefetched, uint32_t aparity, const MegaParams& mp, const __half* kc,
                                          const __half* vc, int item, int warp, int lane, int tid, int pos, int* smem_flag) {
    constexpr int HD = DPL * 32;
    const int n_splits = mp.attn_splits;
    const int kvh = item / n_splits, split = item % n_splits;
    const int grp = mp.n_head / mp.n_kv;
    const int head = kvh * grp + warp;
    const bool active = warp < grp;
    const int L = pos + 1;
    const int n_pages = (L + KV_PAGE_TOKENS - 1) / KV_PAGE_TOKENS;
    const int pps = (n_pages + n_splits - 1) / n_splits;
    const int pg0 = split * pps, pg1 = min(n_pages, pg0 + pps);
    const int npre = prefetched ? max(0, min(ATTN_PRE_PAGES, pg1 - pg0)) : 0;
    if (active) {
        float q[DPL], o[DPL];
        const float* qp = mp.q + (size_t)head * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) { q[d] = __ldcg(qp + d) * mp.attn_scale; o[d] = 0.f; }
        // the newest position's K/V row was written by the QKV phase that just ended (the staged copy may predate it):
        // fetch it in the same round trip as q
        const int pg_new = pos / KV_PAGE_TOKENS, j_new = pos % KV_PAGE_TOKENS;
        uint2 kn = make_uint2(0u, 0u), vn = make_uint2(0u, 0u);
        const bool mine_new = pg_new >= pg0 && pg_new < pg0 + npre;
        if (mine_new) {
            const int page = __ldcg(mp.page_table + pg_new);
            const size_t off = ((size_t)page * mp.n_kv + kvh) * KV_PAGE_TOKENS * HD + (size_t)j_new * HD + lane * DPL;
            if (DPL == 4) { kn = __ldcg(reinterpret_cast<const uint2*>(kc + off)); vn = __ldcg(reinterpret_cast<const uint2*>(vc + off)); }
            else { kn.x = __ldcg(reinterpret_cast<const unsigned*>(kc + off)); vn.x = __ldcg(reinterpret_cast<const unsigned*>(vc + off)); }
        }
        float m_run = -INFINITY, l_run = 0.f;
        if (npre > 0) mbar_wait(reinterpret_cast<uint64_t*>(abuf + ATTN_SMEM_BYTES - 64), aparity);
        const uint32_t page_bytes = (uint32_t)(KV_PAGE_TOKENS * HD * sizeof(__half));
        for (int i = 0; i < npre; ++i) {
            const int pg = pg0 + i;
            uint2 kk[KV_PAGE_TOKENS], vv[KV_PAGE_TOKENS];
            const uint8_t* kb = abuf + (size_t)i * page_bytes + lane * DPL * 2;
            const uint8_t* vb = abuf + (size_t)(ATTN_PRE_PAGES + i) * page_bytes + lane * DPL * 2;
#pragma unroll
            for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
                if (DPL == 4) {
                    kk[j] = *reinterpret_cast<const uint2*>(kb + (size_t)j * HD * 2);
                    vv[j] = *reinterpret_cast<const uint2*>(vb + (size_t)j * HD * 2);
                } else {
                    kk[j] = make_uint2(*reinterpret_cast<const unsigned*>(kb + (size_t)j * HD * 2), 0u);
                    vv[j] = make_uint2(*reinterpret_cast<const unsigned*>(vb + (size_t)j * HD * 2), 0u);
                }
                if (mine_new && pg == pg_new && j == j_new) { kk[j] = kn; vv[j] = vn; }
            }
            attn_page_math<DPL>(kk, vv, min(KV_PAGE_TOKENS, L - pg * KV_PAGE_TOKENS), q, o, m_run, l_run);
        }
        for (int pg = pg0 + npre; pg < pg1; ++pg) {       // not staged (long contexts / no prefetch): straight from global memory
            uint2 kk[KV_PAGE_TOKENS], vv[KV_PAGE_TOKENS];
            attn_load_rows<DPL>(mp, kc, kvh, pg, L, lane, kk);
            attn_load_rows<DPL>(mp, vc, kvh, pg, L, lane, vv);
            attn_page_math<DPL>(kk, vv, min(KV_PAGE_TOKENS, L - pg * KV_PAGE_TOKENS), q, o, m_run, l_run);
        }
        float* po = mp.part_o + ((size_t)head * n_splits + split) * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) po[d] = o[d];
        if (lane == 0) {
            mp.part_ml[((size_t)head * n_splits + split) * 2] = m_run;
            mp.part_ml[((size_t)head * n_splits + split) * 2 + 1] = l_run;
        }
    }
    named_bar_sync(1, NT);
    if (tid == 0) {
        unsigned ticket;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(mp.attn_counters + kvh) : "memory");
        const int last = (ticket == (unsigned)n_splits - 1);
        if (last) mp.attn_counters[kvh] = 0;
        *smem_flag = last;
    }
    named_bar_sync(1, NT);
    if (*smem_flag && active) {
        attn_merge_head<DPL>(mp.part_o, mp.part_ml, mp.attn_out, head, n_splits, n_splits, lane);
    }
}

template <int ABITS, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1) decode_mega_kernel(const __grid_constant__ MegaParams mp) {
    constexpr int NT = NW * 32;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, G = gridDim.x;
    const int fixed = gemv_fixed_smem(mp.max_cols);
    constexpr int DESC_SLOT = 384;                                                // bytes per descriptor buffer
    MegaPhase* sdesc = reinterpret_cast<MegaPhase*>(smem + fixed);              // 2 x 384 B: phase descriptors, double-buffered
    int* sflag = reinterpret_cast<int*>(smem + fixed + 2 * DESC_SLOT);
    float* sstat = reinterpret_cast<float*>(smem + fixed + 2 * DESC_SLOT + 16);  // 3 x NW floats (< 1024 in all)
    uint8_t* abuf = smem + fixed + 1024;                                          // attention staging: K/V tiles + mbarrier
    uint64_t* abar = reinterpret_cast<uint64_t*>(abuf + ATTN_SMEM_BYTES - 64);
    uint32_t aparity = 0;
    Ring ring;
    ring.init(smem, smem + fixed + 1024 + ATTN_SMEM_BYTES, mp.n_tracks, mp.depth, mp.slot_bytes);
    ring.init_barriers(tid);
    if (tid == 0) mbar_init(abar, 1);
    if (tid < RING_MAX_SLOTS) fence_mbar_init();
    __syncthreads();

    if (warp == NW) {
        // ============ producer: every weight phase of every step, never blocked by grid barriers ============
        Track tr{0u, 0u};
        for (int step = 0; step < mp.n_steps; ++step)
            for (int i = 0; i < mp.n_prod; ++i) gemv_produce(mp.prod[i], ring, tr, lane, cta, G);
        return;
    }

    // ============ consumers ==================================================================================
    static_assert(sizeof(MegaPhase) <= DESC_SLOT && sizeof(MegaPhase) % 4 == 0, "descriptor slot");
    constexpr int DESC_WORDS = sizeof(MegaPhase) / 4;
    auto prefetch_desc = [&](int ph, int slot) {      // static data: plain loads; visible after the next barrier
        if (tid < DESC_WORDS)
            reinterpret_cast<uint32_t*>(sdesc)[slot * (DESC_SLOT / 4) + tid] = reinterpret_cast<const uint32_t*>(mp.phases + ph)[tid];
    };
    StepState* st = mp.st;
    const unsigned bar_base = __ldcg(&st->bar_base);
    unsigned nbar = 0;
    Track trk{0u, 0u};            // this warp's position in its ring track (mirrors its producer lane's)
    int dslot = 0;
    for (int step = 0; step < mp.n_steps; ++step) {
        // ---- token embedding (CTA 0) ----------------------------------------------------------------------
        if (cta == 0) {
            int tok;
            {
                const int pos = __ldcg(&st->pos), np = __ldcg(&st->n_prompt);
                tok = __ldcg(&st->token);
                if (pos < np) tok = __ldcg(mp.prompt_ids + pos);
            }
            const uint8_t* row = mp.embd_w + (size_t)tok * mp.embd_row_bytes;
            for (int c = tid; c < mp.n_embd; c += NT) mp.x[c] = dequant_native_elem(row, mp.embd_type, c);
        }
        prefetch_desc(0, dslot);
        ++nbar;
        grid_barrier<NT>(mp.bar_counter, bar_base + nbar * G, tid);
        const int step_pos = __ldcg(&st->pos);        // fixed for the whole step
        const EpiCtx step_ec{step_pos, __ldcg(mp.page_table + step_pos / KV_PAGE_TOKENS)};

        for (int ph = 0; ph < mp.n_phases; ++ph) {
            const MegaPhase& P = *reinterpret_cast<const MegaPhase*>(reinterpret_cast<const uint8_t*>(sdesc) + dslot * DESC_SLOT);
            const int kind = P.kind;
            unsigned long long* tr = (mp.trace != nullptr && tid == 0) ? mp.trace + ((size_t)cta * (mp.n_phases + 1) + ph) * 4 : nullptr;
            if (tr) tr[0] = gtime();
            if (kind == PH_GEMV) {
                PrologueStatic ps;
                gemv_prologue_static<NW>(P.g, tid, ps);
                const float scale = gemv_prologue<ABITS, NW>(P.g, smem, tid, ps);
                if (tr) tr[1] = gtime();
                gemv_consume<ABITS>(P.g, ring, trk, smem, tid, scale, step_ec, cta, G);
                if (P.flags & PHF_HEAD) {
                    // per-CTA softmax statistics over the logits rows this CTA produced
                    named_bar_sync(1, NT);
                    // the logits rows this CTA produced: its item range of the (single) lm_head matrix
                    WorkRange wr = cta_range(P.g.pd.seg[0].n_items, cta, G);
                    wr.a = min(wr.a * (int)P.g.pd.rpi[0], P.g.pd.seg[0].rows);
                    wr.b = min(wr.b * (int)P.g.pd.rpi[0], P.g.pd.seg[0].rows);
                    float best = -INFINITY;
                    int bi = 0x7fffffff;
                    for (int i = wr.a + tid; i < wr.b; i += NT) {
                        const float v = __ldcg(mp.logits + i);
                        if (v > best) { best = v; bi = i; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                    }
                    if (lane == 0) { sstat[warp] = best; reinterpret_cast<int*>(sstat)[NW + warp] = bi; }
                    named_bar_sync(1, NT);
                    best = sstat[0]; bi = reinterpret_cast<int*>(sstat)[NW];
                    for (int w = 1; w < NW; ++w) {
                        const float ov = sstat[w];
                        const int oi = reinterpret_cast<int*>(sstat)[NW + w];
                        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                    }
                    float s = 0.f;
                    for (int i = wr.a + tid; i < wr.b; i += NT) s += expf(__ldcg(mp.logits + i) - best);
                    s = warp_sum(s);
                    if (lane == 0) sstat[2 * NW + warp] = s;
                    named_bar_sync(1, NT);
                    if (tid == 0) {
                        float tot = 0.f;
                        for (int w = 0; w < NW; ++w) tot += sstat[2 * NW + w];
                        mp.head_part[cta * 4 + 0] = best;
                        reinterpret_cast<int*>(mp.head_part)[cta * 4 + 1] = bi;
                        mp.head_part[cta * 4 + 2] = tot;
                    }
                    if (mp.logits_keep != nullptr) {
                        const int oi = __ldcg(&st->out_idx);
                        if (!__ldcg(&st->done) && oi < mp.max_out) {
                            float* dst = mp.logits_keep + (size_t)oi * P.g.pd.seg[0].rows;
                            for (int i = wr.a + tid; i < wr.b; i += NT) dst[i] = __ldcg(mp.logits + i);
                        }
                    }
                }
            } else if (kind == PH_ATTN) {
                // only reached when the attention phase does not directly follow a QKV phase (never in the current tables)
                const int n_items = mp.n_kv * mp.attn_splits;
                if (cta < n_items) {
                    if (mp.head_dim == 128) attn_item<4, NT>(abuf, false, 0u, mp, P.g.k_cache, P.g.v_cache, cta, warp, lane, tid, step_pos, sflag);
                    else attn_item<2, NT>(abuf, false, 0u, mp, P.g.k_cache, P.g.v_cache, cta, warp, lane, tid, step_pos, sflag);
                }
            }
            if (kind == PH_GEMV && P.g.epi == EPI_QKV && ph + 2 < mp.n_phases) {
                // QKV -> [barrier] -> attention, fused: request this CTA's K/V pages of the old positions before the
                // barrier, run the attention item right after it, then skip the table's ATTN entry.
                const int n_items = mp.n_kv * mp.attn_splits;
                const __half* kc = P.g.k_cache;
                const __half* vc = P.g.v_cache;
                // (the staging buffer was last read two barriers ago; tid 0's own reads are ordered before these async writes)
                if (cta < n_items && tid == 0) attn_prefetch(abuf, mp, kc, vc, cta, step_pos);
                if (tr) tr[2] = gtime();
                prefetch_desc(ph + 2, dslot ^ 1);            // descriptor of the phase after attention (attn_output)
                ++nbar;
                grid_barrier<NT>(mp.bar_counter, bar_base + nbar * G, tid);
                if (tr) { tr[3] = gtime(); tr += 4; tr[0] = tr[1] = gtime(); }
                if (cta < n_items) {
                    if (mp.head_dim == 128) attn_item<4, NT>(abuf, true, aparity, mp, kc, vc, cta, warp, lane, tid, step_pos, sflag);
                    else attn_item<2, NT>(abuf, true, aparity, mp, kc, vc, cta, warp, lane, tid, step_pos, sflag);
                    // the barrier flipped phase only if something was staged for this item
                    {
                        const int L = step_pos + 1, npg = (L + KV_PAGE_TOKENS - 1) / KV_PAGE_TOKENS;
                        const int pps = (npg + mp.attn_splits - 1) / mp.attn_splits;
                        const int pg0 = (cta % mp.attn_splits) * pps;
                        if (min(npg, pg0 + pps) - pg0 > 0) aparity ^= 1;
                    }
                }
                if (tr) tr[2] = gtime();
                dslot ^= 1;
                ++ph;                                         // the ATTN entry is done
                ++nbar;
                grid_barrier<NT>(mp.bar_counter, bar_base + nbar * G, tid);
                if (tr) tr[3] = gtime();
                continue;
            }
            if (tr) tr[2] = gtime();
            // the next phase's descriptor travels while this CTA waits at the barrier
            if (ph + 1 < mp.n_phases) prefetch_desc(ph + 1, dslot ^ 1);
            dslot ^= 1;
            ++nbar;
            grid_barrier<NT>(mp.bar_counter, bar_base + nbar * G, tid);
            if (tr) tr[3] = gtime();
        }

        // ---- sampling / state advance (CTA 0), published to everyone by the barrier at the top of the next step
        if (cta == 0 && warp == 0) {
            if (mp.with_head) {
                float best = -INFINITY, sum = 0.f;
                int bi = 0x7fffffff;
                for (int c = lane; c < G; c += 32) {
                    const float m = __ldcg(mp.head_part + c * 4);
                    const int ix = __ldcg(reinterpret_cast<const int*>(mp.head_part) + c * 4 + 1);
                    if (m > best || (m == best && ix < bi)) { best = m; bi = ix; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                for (int c = lane; c < G; c += 32) {
                    const float m = __ldcg(mp.head_part + c * 4);
                    const float sc = __ldcg(mp.head_part + c * 4 + 2);
                    if (m > -INFINITY) sum += sc * expf(m - best);
                }
                sum = warp_sum(sum);
                if (lane == 0 && !__ldcg(&st->done)) {
                    const int oi = __ldcg(&st->out_idx);
                    if (oi < mp.max_out) {
                        mp.out_ids[oi] = bi;
                        mp.out_logprobs[oi] = -logf(sum);
                    }
                    st->token = bi;
                    st->pos = __ldcg(&st->pos) + 1;
                    st->out_idx = oi + 1;
                    if (!st->ignore_eos) {
                        for (int k = 0; k < st->n_stop; ++k)
                            if (st->stop_ids[k] == bi) st->done = 1;
                    }
                }
            } else if (lane == 0) {
                st->pos = __ldcg(&st->pos) + 1;
            }
        }
        if (cta == 0) named_bar_sync(1, NT);     // state written before CTA 0 starts the next embedding
    }
    if (cta == 0 && tid == 0) st->bar_base = bar_base + nbar * G;
}

}  // namespace

size_t mega_smem_bytes(int max_cols, int n_slots, int slot_bytes) {
    return (size_t)gemv_fixed_smem(max_cols) + 1024 + ATTN_SMEM_BYTES + (size_t)n_slots * slot_bytes;
}

namespace {
template <int ABITS, int NW>
cudaError_t mega_configure_one() {
    return cudaFuncSetAttribute(decode_mega_kernel<ABITS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
template <int ABITS, int NW>
cudaError_t mega_launch_one(const cudaLaunchConfig_t& cfg, const MegaParams& mp) {
    return cudaLaunchKernelEx(&cfg, decode_mega_kernel<ABITS, NW>, mp);
}
}  // namespace

cudaError_t mega_configure() {
    cudaError_t e = mega_configure_one<16, 8>();
    if (e == cudaSuccess) e = mega_configure_one<16, 12>();
    if (e == cudaSuccess) e = mega_configure_one<16, 16>();
    if (e == cudaSuccess) e = mega_configure_one<8, 8>();
    if (e == cudaSuccess) e = mega_configure_one<8, 12>();
    if (e == cudaSuccess) e = mega_configure_one<8, 16>();
    return e;
}

cudaError_t mega_launch(const MegaParams& mp, int abits, int nw, int n_ctas, cudaStream_t s) {
    if (!gemv_variant_ok(abits, nw)) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3((unsigned)gemv_threads(nw));
    cfg.dynamicSmemBytes = mega_smem_bytes(mp.max_cols, mp.n_tracks * mp.depth, mp.slot_bytes);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: the grid barrier depends on it
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (abits == 16) {
        if (nw == 8) return mega_launch_one<16, 8>(cfg, mp);
        if (nw == 12) return mega_launch_one<16, 12>(cfg, mp);
        return mega_launch_one<16, 16>(cfg, mp);
    }
    if (nw == 8) return mega_launch_one<8, 8>(cfg, mp);
    if (nw == 12) return mega_launch_one<8, 12>(cfg, mp);
    return mega_launch_one<8, 16>(cfg, mp);
}

}  // namespace gl
