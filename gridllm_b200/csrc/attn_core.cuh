// Shared pieces of the paged decode attention (stand-alone kernel in attention.cu, fused phase in decode_mega.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace gl {

// scores of one page: returns this lane's softmax weight w (for position lane >> 1) after updating (m_run, l_run) and
// rescaling o; kk = this lane's DPL dims of the 16 K rows
template <int DPL>
__device__ __forceinline__ float attn_page_scores(const uint2* kk, int npos, const float* q, float* o, float& m_run, float& l_run) {
    static_assert(KV_PAGE_TOKENS == 16, "the butterfly below is written for 16 positions per page");
    const int lane = threadIdx.x & 31;
    float a[KV_PAGE_TOKENS];
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&kk[j].x));
        float t = q[0] * k0.x + q[1] * k0.y;
        if (DPL == 4) {
            const float2 k1 = __half22float2(*reinterpret_cast<const __half2*>(&kk[j].y));
            t += q[DPL - 2] * k1.x + q[DPL - 1] * k1.y;
        }
        a[j] = t;
    }
    // after the step with xor w, a lane keeps the half of its values selected by its bit w and adds the partner's
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b4 ? a[i] : a[i + 8], keep = b4 ? a[i + 8] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b3 ? b[i] : b[i + 4], keep = b3 ? b[i + 4] : b[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b2 ? c[i] : c[i + 2], keep = b2 ? c[i + 2] : c[i];
        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    float s;
    {
        const float send = b1 ? d[0] : d[1], keep = b1 ? d[1] : d[0];
        s = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);          // score of position (lane >> 1), held by both lanes of the pair
    const int pj = lane >> 1;
    if (pj >= npos) s = -INFINITY;
    float m_t = s;
    m_t = fmaxf(m_t, __shfl_xor_sync(0xffffffffu, m_t, 2));
    m_t = fmaxf(m_t, __shfl_xor_sync(0xffffffffu, m_t, 4));
    m_t = fmaxf(m_t, __shfl_xor_sync(0xffffffffu, m_t, 8));
    m_t = fmaxf(m_t, __shfl_xor_sync(0xffffffffu, m_t, 16));
    const float m_new = fmaxf(m_run, m_t);             // npos >= 1, so m_new is finite
    const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
    const float w = (pj < npos) ? expf(s - m_new) : 0.f;
    float ws = w;
    ws += __shfl_xor_sync(0xffffffffu, ws, 2);
    ws += __shfl_xor_sync(0xffffffffu, ws, 4);
    ws += __shfl_xor_sync(0xffffffffu, ws, 8);
    ws += __shfl_xor_sync(0xffffffffu, ws, 16);        // every position once (the xor-1 partner holds the duplicate)
    l_run = l_run * corr + ws;
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) o[dd] *= corr;
    m_run = m_new;
    return w;
}

// o += w_j * V_j for one row: the weight of position j is broadcast from the lane pair that holds it
template <int DPL>
__device__ __forceinline__ void attn_pv_row(float w, int j, uint2 vrow, float* o) {
    const float wj = __shfl_sync(0xffffffffu, w, 2 * j);
    const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&vrow.x));
    o[0] += wj * v0.x; o[1] += wj * v0.y;
    if (DPL == 4) {
        const float2 v1 = __half22float2(*reinterpret_cast<const __half2*>(&vrow.y));
        o[DPL - 2] += wj * v1.x; o[DPL - 1] += wj * v1.y;
    }
}

// One 16-token page against one query head.  kk / vv hold this lane's DPL dims of the 16 K / V rows (8 B each for
// head_dim 128).  The 16 per-lane partial dots are reduced with a TRANSPOSING butterfly (8 + 4 + 2 + 1 + 1 = 16 shuffles
// instead of 16 x 5): afterwards lanes 2p and 2p+1 hold the score of position p, so the exponential is evaluated once
// per lane; max and sum take 4 shuffles each and the 16 weights are broadcast for the P V update (16 shuffles).
template <int DPL>
__device__ __forceinline__ void attn_page_math(const uint2* kk, const uint2* vv, int npos, const float* q, float* o, float& m_run, float& l_run) {
    const float w = attn_page_scores<DPL>(kk, npos, q, o, m_run, l_run);
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        if (j >= npos) break;                                      // warp-uniform; rows beyond npos may hold anything
        attn_pv_row<DPL>(w, j, vv[j], o);
    }
}

// the same with the page staged in shared memory: K rows are read for the scores, V rows only afterwards (fewer live
// registers, which is what lets an attention CTA share an SM with the GEMV kernels around it)
template <int DPL>
__device__ __forceinline__ void attn_page_math_smem(const __half* kb, const __half* vb, int npos, const float* q, float* o, float& m_run,
                                                    float& l_run) {
    constexpr int HD = DPL * 32;
    uint2 kk[KV_PAGE_TOKENS];
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        if (DPL == 4) kk[j] = *reinterpret_cast<const uint2*>(kb + j * HD);
        else kk[j] = make_uint2(*reinterpret_cast<const unsigned*>(kb + j * HD), 0u);
    }
    const float w = attn_page_scores<DPL>(kk, npos, q, o, m_run, l_run);
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        if (j >= npos) break;
        uint2 v;
        if (DPL == 4) v = *reinterpret_cast<const uint2*>(vb + j * HD);
        else v = make_uint2(*reinterpret_cast<const unsigned*>(vb + j * HD), 0u);
        attn_pv_row<DPL>(w, j, v, o);
    }
}


// Merge the split partials of one head (run by one warp; n_splits <= 32): out = sum_s w_s o_s / sum_s w_s l_s with
// w_s = exp(m_s - max m).  (m, l) of split `lane` and the partial outputs of 8 splits travel in one round trip.
// `stride` is the number of partial slots per head in the buffers, n_splits (<= stride) how many of them are in use.
template <int DPL, int MB = 16>      // MB partials per round trip
__device__ __forceinline__ void attn_merge_head(const float* part_o, const float* part_ml, float* attn_out, int head, int stride, int n_splits,
                                                int lane) {
    constexpr int HD = DPL * 32;
    float ms = -INFINITY, ls = 0.f;
    if (lane < n_splits) {
        ms = __ldcg(part_ml + ((size_t)head * stride + lane) * 2);
        ls = __ldcg(part_ml + ((size_t)head * stride + lane) * 2 + 1);
    }
    const float* pbase = part_o + (size_t)head * stride * HD + lane * DPL;
    float acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[d] = 0.f;
    float M = 0.f, wl = 0.f, den = 0.f;
    for (int s0 = 0; s0 < n_splits; s0 += MB) {
        float po[MB][DPL];
        // one vector load per partial (lane * DPL floats are 16- / 8-byte aligned: HD is a multiple of 32 * DPL)
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            if (s0 + i < n_splits) {
                if (DPL == 4) {
                    const float4 t = __ldcg(reinterpret_cast<const float4*>(pbase + (size_t)(s0 + i) * HD));
                    po[i][0] = t.x; po[i][1] = t.y; po[i][DPL - 2] = t.z; po[i][DPL - 1] = t.w;
                } else {
                    const float2 t = __ldcg(reinterpret_cast<const float2*>(pbase + (size_t)(s0 + i) * HD));
                    po[i][0] = t.x; po[i][1] = t.y;
                }
            } else {
#pragma unroll
                for (int d = 0; d < DPL; ++d) po[i][d] = 0.f;
            }
        }
        if (s0 == 0) {
            M = warp_max(ms);
            wl = (ms == -INFINITY) ? 0.f : expf(ms - M);
            den = warp_sum(wl * ls);
        }
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            const float w = __shfl_sync(0xffffffffu, wl, (s0 + i) & 31);
            if (s0 + i < n_splits) {
#pragma unroll
                for (int d = 0; d < DPL; ++d) acc[d] += w * po[i][d];
            }
        }
    }
    const float inv = 1.0f / den;
    float* out = attn_out + (size_t)head * HD + lane * DPL;
    if (DPL == 4) *reinterpret_cast<float4*>(out) = make_float4(acc[0] * inv, acc[1] * inv, acc[DPL - 2] * inv, acc[DPL - 1] * inv);
    else *reinterpret_cast<float2*>(out) = make_float2(acc[0] * inv, acc[1] * inv);
}

}  // namespace gl
