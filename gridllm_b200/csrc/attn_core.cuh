// Shared pieces of the paged decode attention (stand-alone kernel in attention.cu, fused phase in decode_mega.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace gl {

// One 16-token page against one query head.  kk / vv hold this lane's DPL dims of the 16 K / V rows (8 B each for
// head_dim 128); the 16 positions are 16 independent dot + shuffle-reduce chains, then one online-softmax update.
template <int DPL>
__device__ __forceinline__ void attn_page_math(const uint2* kk, const uint2* vv, int npos, const float* q, float* o, float& m_run, float& l_run) {
    float sc[KV_PAGE_TOKENS];
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&kk[j].x));
        float a = q[0] * k0.x + q[1] * k0.y;
        if (DPL == 4) {
            const float2 k1 = __half22float2(*reinterpret_cast<const __half2*>(&kk[j].y));
            a += q[DPL - 2] * k1.x + q[DPL - 1] * k1.y;
        }
        sc[j] = a;
    }
    float m_t = -INFINITY;
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        sc[j] = warp_sum(sc[j]);
        if (j < npos) m_t = fmaxf(m_t, sc[j]);
    }
    const float m_new = fmaxf(m_run, m_t);
    const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
    l_run *= corr;
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < KV_PAGE_TOKENS; ++j) {
        if (j < npos) {
            const float w = expf(sc[j] - m_new);
            l_run += w;
            const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&vv[j].x));
            o[0] += w * v0.x; o[1] += w * v0.y;
            if (DPL == 4) {
                const float2 v1 = __half22float2(*reinterpret_cast<const __half2*>(&vv[j].y));
                o[DPL - 2] += w * v1.x; o[DPL - 1] += w * v1.y;
            }
        }
    }
    m_run = m_new;
}


// Merge the split partials of one head (run by one warp; n_splits <= 32): out = sum_s w_s o_s / sum_s w_s l_s with
// w_s = exp(m_s - max m).  (m, l) of split `lane` and the partial outputs of 8 splits travel in one round trip.
// `stride` is the number of partial slots per head in the buffers, n_splits (<= stride) how many of them are in use.
template <int DPL>
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
    for (int s0 = 0; s0 < n_splits; s0 += 8) {
        float po[8][DPL];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int d = 0; d < DPL; ++d) po[i][d] = (s0 + i < n_splits) ? __ldcg(pbase + (size_t)(s0 + i) * HD + d) : 0.f;
        }
        if (s0 == 0) {
            M = warp_max(ms);
            wl = (ms == -INFINITY) ? 0.f : expf(ms - M);
            den = warp_sum(wl * ls);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float w = __shfl_sync(0xffffffffu, wl, (s0 + i) & 31);
            if (s0 + i < n_splits) {
#pragma unroll
                for (int d = 0; d < DPL; ++d) acc[d] += w * po[i][d];
            }
        }
    }
    const float inv = 1.0f / den;
    float* out = attn_out + (size_t)head * HD + lane * DPL;
#pragma unroll
    for (int d = 0; d < DPL; ++d) out[d] = acc[d] * inv;
}

}  // namespace gl
