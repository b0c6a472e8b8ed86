// Device-side building blocks of the weight-streaming GEMV, shared by the stand-alone kernel
// (gemv.cu) and the persistent decode-step kernel (decode_mega.cu).
//
//   producer lane : gemv_produce()  -- 1-D TMA bulk copies of whole rows into an mbarrier ring
//   consumer warps: gemv_prologue() -- (RMSNorm) + snap x to int8 planes -> this lane's XUnit registers
//                   gemv_consume()  -- dp4a block decode out of shared memory, warp-shuffle reduction,
//                                      fused epilogue (store / residual add / RoPE+KV append / SiLU*mul)
//
// For K <= 4096 (one warp spans a row) every consumer warp is autonomous: it waits for a stage, reduces its
// own rows, runs their epilogue from lane 0 with operands prefetched before the dot product, and releases
// the stage -- there is no CTA-wide barrier in the steady state.  For larger K (ffn_down) the 2/4/8 warps
// that share a row meet at a named barrier of just those warps.
#pragma once
#include "common.cuh"
#include "gguf_file.h"
#include "kernels.h"
#include "rowdot.h"

namespace gl {

constexpr int GEMV_MAX_WARPS = 16;      // consumer warps per CTA are a template parameter (8, 12 or 16)

// shared-memory carve-up (bytes from the start of dynamic smem)
constexpr int SM_BARS = 0;                 // full[8], empty[8] mbarriers
constexpr int SM_RED = 128;                // 32 floats: RMSNorm partials
constexpr int SM_RES = 256;                // 2 x 256 floats: cross-warp partial sums (K > 4096)
constexpr int SM_X = 256 + 2 * 256 * 4;    // x planes start (2304)

__host__ __device__ inline int gemv_x_bytes(int cols) { return 2 * cols + cols / 2; }   // hi, lo, sx, sm, s16
__host__ __device__ inline int gemv_fixed_smem(int cols) { return (SM_X + gemv_x_bytes(cols) + 127) & ~127; }

__host__ __device__ inline int warps_per_row(int cols) {
    const int w = (cols / UNIT_COLS + 31) / 32;
    int p = 1;
    while (p < w) p <<= 1;
    return p;       // 1,2,4,8
}

struct Ring {
    uint64_t* full;
    uint64_t* empty;
    uint8_t* slots;
    int n_slots;
    int slot_bytes;
    int st;
    uint32_t ph;
    __device__ __forceinline__ void advance() { if (++st == n_slots) { st = 0; ph ^= 1; } }
    __device__ __forceinline__ uint8_t* slot() const { return slots + (size_t)st * slot_bytes; }
};

struct WorkRange { int a, b; };
__device__ __forceinline__ WorkRange cta_range(int rows, int gran, int cta, int n_ctas) {
    const int units = rows / gran;
    WorkRange r;
    r.a = (int)(((long long)cta * units) / n_ctas) * gran;
    r.b = (int)(((long long)(cta + 1) * units) / n_ctas) * gran;
    return r;
}

// ---------------------------------------------------------------------------------------------------
// producer (one lane): stream this CTA's rows of every segment
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gemv_produce(const GemvParams& p, Ring& ring, int cta, int n_ctas) {
    const int gran = (p.epi == EPI_QKV) ? 2 : 1;
    const int nwork = p.pair ? 1 : p.nseg;
    for (int s = 0; s < nwork; ++s) {
        const GemvSeg sg = p.seg[s];
        const WorkRange wr = cta_range(sg.rows, gran, cta, n_ctas);
        for (int r0 = wr.a; r0 < wr.b; r0 += sg.rows_per_stage) {
            const int n = min(sg.rows_per_stage, wr.b - r0);
            const uint32_t bytes = (uint32_t)n * (uint32_t)sg.row_stride;
            mbar_wait(&ring.empty[ring.st], ring.ph ^ 1);
            uint8_t* dst = ring.slot();
            if (p.pair) {
                mbar_expect_tx(&ring.full[ring.st], 2 * bytes);
                tma_load_1d(dst, sg.w + (size_t)r0 * sg.row_stride, bytes, &ring.full[ring.st]);
                tma_load_1d(dst + bytes, p.seg[1].w + (size_t)r0 * sg.row_stride, bytes, &ring.full[ring.st]);
            } else {
                mbar_expect_tx(&ring.full[ring.st], bytes);
                tma_load_1d(dst, sg.w + (size_t)r0 * sg.row_stride, bytes, &ring.full[ring.st]);
            }
            ring.advance();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// consumer prologue: x -> (RMSNorm) -> int8 planes in smem -> XUnit registers of this lane
// All NCT consumer threads must call it (named barrier 1).
// ---------------------------------------------------------------------------------------------------
// Returns the scalar every row sum of this phase must be multiplied by: 1, or the RMSNorm factor
// rstd = 1/sqrt(mean(x^2)+eps).  The fixed-point integers v = rint(x*w / amax_blk(x*w) * RANGE) do not depend on
// rstd (it cancels), so the planes are built from x*w while the sum of squares is still being reduced, and rstd is
// applied once per output row.  One named barrier in total.
template <int ABITS, int NW>
__device__ __forceinline__ float gemv_prologue(const GemvParams& p, uint8_t* smem, int tid) {
    constexpr int NT = NW * 32;
    constexpr int NB = 8;                  // float4 loads in flight per thread (one L2 round trip per batch)
    const int K = p.cols;
    const int warp = tid >> 5, lane = tid & 31;
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    uint8_t* xhi = smem + SM_X;
    uint8_t* xlo = xhi + K;
    float* sx_arr = reinterpret_cast<float*>(xlo + K);
    float* sm_arr = sx_arr + K / 32;
    int* s16_arr = reinterpret_cast<int*>(sm_arr + K / 32);

    // Thread t owns float4 #(t + i*NT): a warp instruction covers 512 contiguous bytes (= four 32-column blocks, one
    // 128-column unit), eight lanes share a block.  Every read of data produced upstream uses ld.global.cg: this CTA
    // may have been resident (and its SM's L1 populated) before the producer of x finished.
    const int nf = K / 4;                  // K % 256 == 0, so liveness below is uniform over a warp
    const bool norm = p.norm_w != nullptr;
    const float4* x4 = reinterpret_cast<const float4*>(p.x);
    const float4* w4 = reinterpret_cast<const float4*>(p.norm_w);
    float ss = 0.f;
    for (int f0 = tid; f0 < nf; f0 += NB * NT) {
        float4 v[NB], wv[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int f = f0 + i * NT;
            if (f < nf) v[i] = __ldcg(x4 + f);
        }
        if (norm) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int f = f0 + i * NT;
                if (f < nf) wv[i] = w4[f];
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int f = f0 + i * NT;
            if (f < nf) {
                float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                if (norm) {
                    ss += e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3];
                    e[0] *= wv[i].x; e[1] *= wv[i].y; e[2] *= wv[i].z; e[3] *= wv[i].w;
                }
                float amax = fmaxf(fmaxf(fabsf(e[0]), fabsf(e[1])), fmaxf(fabsf(e[2]), fabsf(e[3])));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
                uint32_t h, l;
                int vs;
                snap4<ABITS>(e, snap_inv<ABITS>(amax), &h, &l, &vs);
                vs += __shfl_xor_sync(0xffffffffu, vs, 1);
                vs += __shfl_xor_sync(0xffffffffu, vs, 2);           // sum(v) of this 16-column group
                const int vs32 = vs + __shfl_xor_sync(0xffffffffu, vs, 4);
                const int blk = f >> 3, half = (f >> 2) & 1;
                const int u = blk >> 2;
                const int phys = (2 * (blk & 3) + half) ^ (u & 7);
                const int off = u * 128 + phys * 16 + (f & 3) * 4;
                *reinterpret_cast<uint32_t*>(xhi + off) = h;
                if (ABITS == 16) *reinterpret_cast<uint32_t*>(xlo + off) = l;
                if ((f & 3) == 0) s16_arr[2 * blk + half] = vs;
                if ((f & 7) == 0) {
                    const float sx = amax / (ABITS == 16 ? ACT16_RANGE : ACT8_RANGE);
                    sx_arr[blk] = sx;
                    sm_arr[blk] = sx * (float)vs32;
                }
            }
        }
    }
    if (norm) {
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
    }
    named_bar_sync(1, NT);
    float scale = 1.f;
    if (norm) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += red[w];
        scale = 1.0f / sqrtf(tot / (float)K + p.eps);
    }

    // NOTE: the planes (and red[]) are rewritten only by the NEXT prologue, which every caller separates from this
    // point by a CTA-wide barrier (end of kernel, or the grid barrier of the persistent kernel).
    return scale;
}

// reduce four per-lane partials over the warp with 6 shuffles.  On return, lane L holds the warp total of row
// ((L >> 3) & 3); i.e. lanes 0-7 -> row 0, 8-15 -> row 1, 16-23 -> row 2, 24-31 -> row 3.
__device__ __forceinline__ float warp_sum4(const float* o, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
    float ka = b4 ? o[2] : o[0], kb = b4 ? o[3] : o[1];
    const float sa = b4 ? o[0] : o[2], sb = b4 ? o[1] : o[3];
    ka += __shfl_xor_sync(0xffffffffu, sa, 16);
    kb += __shfl_xor_sync(0xffffffffu, sb, 16);
    float k = b3 ? kb : ka;
    const float snd = b3 ? ka : kb;
    k += __shfl_xor_sync(0xffffffffu, snd, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

struct EpiCtx {      // per-kernel constants of the QKV epilogue, loaded once
    int pos;
    int page;
};

// lane-0 epilogue of one item (1 or 2 rows)
__device__ __forceinline__ void gemv_epilogue_item(const GemvParams& p, int seg, int r, float v0, float v1, float pre0, float pre1,
                                                   const EpiCtx& ec) {
    if (p.epi == EPI_STORE) {
        p.out[r] = v0;
    } else if (p.epi == EPI_ADD) {
        p.out[r] = pre0 + v0;
    } else if (p.epi == EPI_SILU) {
        p.out[r] = (v0 / (1.0f + expf(-v0))) * v1;
    } else {   // EPI_QKV
        if (seg < 2) {
            const int d = r % p.head_dim;
            const float o0 = v0 * pre0 - v1 * pre1, o1 = v0 * pre1 + v1 * pre0;      // pre0 = cos, pre1 = sin
            if (seg == 0) {
                *reinterpret_cast<float2*>(p.out + r) = make_float2(o0, o1);
            } else {
                const int kvh = r / p.head_dim;
                const size_t off = (((size_t)ec.page * p.n_kv_heads + kvh) * KV_PAGE_TOKENS + (ec.pos % KV_PAGE_TOKENS)) * p.head_dim + d;
                *reinterpret_cast<__half2*>(p.k_cache + off) = __floats2half2_rn(o0, o1);
            }
        } else {
            const int kvh = r / p.head_dim, d = r % p.head_dim;
            const size_t off = (((size_t)ec.page * p.n_kv_heads + kvh) * KV_PAGE_TOKENS + (ec.pos % KV_PAGE_TOKENS)) * p.head_dim + d;
            p.v_cache[off] = __float2half_rn(v0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// consumer main loop over this CTA's stages.  All consumer warps call it.
//
// A warp works on a QUAD of rows per iteration (quad_dot_* in rowdot.h): activations come from the shared-memory
// planes, each x word is reused for four weight rows, 16 independent dp4a chains per lane.  After the 6-shuffle
// reduction lane 8*r holds row r of the quad and runs that row's epilogue (operands prefetched before the dots).
// Quad composition per epilogue kind:
//   STORE / ADD / V rows : rows 4q .. 4q+3 of the stage
//   RoPE pairs (q, k)    : same (pairs (0,1) and (2,3) are adjacent rows)
//   gate/up (SiLU*mul)   : gate rows 2q, 2q+1 and the matching up rows (stage holds n gate rows, then n up rows)
// ---------------------------------------------------------------------------------------------------
template <int ABITS, int NW>
__device__ __forceinline__ void gemv_consume(const GemvParams& p, Ring& ring, uint8_t* smem, int tid, float scale, int cta, int n_ctas) {
    const int K = p.cols;
    const int warp = tid >> 5, lane = tid & 31;
    const int nu = K / UNIT_COLS;
    const int wpr = warps_per_row(K);
    const int ngrp = NW / wpr;               // row groups working in parallel (warps beyond ngrp*wpr only hand stages back)
    const int grp = warp / wpr, wsub = warp % wpr;
    const int u = wsub * 32 + lane;
    const bool valid = u < nu;
    const bool in_grp = grp < ngrp;
    float* res = reinterpret_cast<float*>(smem + SM_RES);
    const int gran = (p.epi == EPI_QKV) ? 2 : 1;
    const int nwork = p.pair ? 1 : p.nseg;

    XPlanes xp;
    {
        uint8_t* xhi = smem + SM_X;
        const int uu = valid ? u : 0;
        xp.hi = xhi + uu * 128;
        xp.lo = xhi + K + uu * 128;
        const float* sx_arr = reinterpret_cast<const float*>(xhi + 2 * K);
        xp.sx = sx_arr + 4 * uu;
        xp.sm = sx_arr + K / 32 + 4 * uu;
        xp.s16 = reinterpret_cast<const int*>(sx_arr + 2 * (K / 32)) + 8 * uu;
        xp.sw = uu & 7;
    }
    EpiCtx ec{0, 0};
    if (p.epi == EPI_QKV) {
        ec.pos = __ldcg(&p.st->pos);
        ec.page = __ldcg(p.page_table + ec.pos / KV_PAGE_TOKENS);
    }
    const int myrow = lane >> 3;             // row of the quad this lane reports after the reduction
    const bool epi_lane = (lane & 7) == 0 && wsub == 0;
    int buf = 0;
    for (int s = 0; s < nwork; ++s) {
        const GemvSeg sg = p.seg[s];
        const WorkRange wr = cta_range(sg.rows, gran, cta, n_ctas);
        const bool pair_adj = (p.epi == EPI_QKV && s < 2);
        const bool pair_gu = p.pair != 0;
        for (int r0 = wr.a; r0 < wr.b; r0 += sg.rows_per_stage) {
            const int n = min(sg.rows_per_stage, wr.b - r0);
            const int nq = pair_gu ? (n + 1) / 2 : (n + 3) / 4;
            const uint8_t* base = ring.slot();
            mbar_wait(&ring.full[ring.st], ring.ph);
            for (int qi = in_grp ? grp : nq; qi < nq; qi += ngrp) {
                // stage-local rows of the quad (clamped: a ragged quad repeats its last row and ignores the result)
                int lr[4];
                if (pair_gu) {
                    const int g0 = 2 * qi, g1 = min(2 * qi + 1, n - 1);
                    lr[0] = g0; lr[1] = g1; lr[2] = n + g0; lr[3] = n + g1;
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) lr[r] = min(4 * qi + r, n - 1);
                }
                // this lane's row for the epilogue
                const int my_local = pair_gu ? (2 * qi + (myrow & 1)) : (4 * qi + myrow);
                const bool my_ok = epi_lane && my_local < n && (pair_gu ? myrow < 2 : (pair_adj ? (myrow & 1) == 0 : true));
                const int grow = r0 + my_local;                 // global row (gate row for gate/up)
                float pre0 = 0.f, pre1 = 0.f;
                if (my_ok) {
                    if (p.epi == EPI_ADD) pre0 = __ldcg(p.resid + grow);
                    else if (pair_adj) {
                        const int d2 = (grow % p.head_dim) >> 1;
                        pre0 = p.rope_cos[(size_t)ec.pos * (p.head_dim / 2) + d2];
                        pre1 = p.rope_sin[(size_t)ec.pos * (p.head_dim / 2) + d2];
                    }
                }
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                if (valid) {
                    const uint8_t* rp[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) rp[r] = base + (size_t)lr[r] * sg.row_stride;
                    if (sg.type == T_Q4_K) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) rp[r] += (size_t)(u >> 1) * 144;
                        quad_dot_q4k<ABITS>(rp, u & 1, xp, o);
                    } else if (sg.type == T_Q6_K) {
                        quad_dot_q6k<ABITS>(rp, K >> 8, u, xp, o);
                    } else {
                        quad_dot_q80<ABITS>(rp, K, u, xp, o);
                    }
                }
                float tot = warp_sum4(o, lane);
                if (wpr > 1) {
                    // K > 4096: wpr warps share the rows; their partials meet in shared memory
                    float* rbuf = res + buf * 256 + grp * 4 * wpr;
                    if ((lane & 7) == 0) rbuf[myrow * wpr + wsub] = tot;
                    named_bar_sync(2 + grp, 32 * wpr);
                    if (wsub == 0) {
                        float t = 0.f;
                        for (int j = 0; j < wpr; ++j) t += rbuf[myrow * wpr + j];
                        tot = t;
                    }
                    buf ^= 1;     // double buffer: the next quad's partials never race the readers of this one
                }
                // partner row of a pair: adjacent row (RoPE) or the up row (gate/up)
                const float other = __shfl_xor_sync(0xffffffffu, tot, pair_gu ? 16 : 8);
                if (my_ok) {
                    const float v0 = tot * scale, v1 = other * scale;      // RMSNorm factor of the fused prologue (1 if none)
                    gemv_epilogue_item(p, s, grow, v0, v1, pre0, pre1, ec);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[ring.st]);      // this warp is done with the stage's bytes
            ring.advance();
        }
    }
}

}  // namespace gl
