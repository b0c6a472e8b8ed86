// Per-lane block-decode arithmetic of the decode GEMV and the engine's HBM row layouts.
//
// Everything here is __host__ __device__ so that tests/hostcheck can run the exact lane
// program on the CPU (layout / bit-twiddling bugs are found without a GPU).  The CUDA kernels
// in gemv.cu call these same functions; nothing in the product path runs them on the host.
//
// A "unit" is the slice of one weight row that one lane owns: 128 consecutive columns
// (half a K-quant super-block, or four Q8_0 blocks).  The lane keeps the matching 128
// activations in registers as two int8 planes (hi, lo) + per-32-column scales for the whole
// kernel, so every weight byte staged in shared memory is touched exactly once.
//
// Engine row layouts in HBM (size-preserving permutations of the GGUF row):
//   Q4_K   : native GGUF (144-B super-blocks; block b at 144*b).  Conflict-free as is because
//            144 = 128 + 16 rotates consecutive blocks across the 16-B shared-memory slots.
//   Q6_K-T : per row of nb super-blocks / nu = 2*nb units:
//              [ql : 4 x nu x 16 B, chunk(i4,u) at (i4*nu+u)*16 ]   bytes h*64+16*i4.. of block b
//              [qh : 2 x nu x 16 B, chunk(i2,u) at (i2*nu+u)*16 ]   bytes h*32+16*i2.. of block b
//              [sc : nu x 8 B ]                                      scales[8h..8h+8) of block b
//              [d  : nb x 2 B ]                                      (u = 2*b + h)
//   Q8_0-T : per row of nu = cols/128 units (4 blocks each):
//              [qs : 8 x nu x 16 B, chunk(i,u) at (i*nu+u)*16 ] [d : nu x 8 B]
//   rows are padded to a multiple of 16 B (TMA bulk-copy granularity).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

namespace gl {

constexpr int UNIT_COLS = 128;
constexpr float ACT16_RANGE = 16256.0f;   // 127*128
constexpr float ACT8_RANGE = 127.0f;

struct alignas(16) U4 { uint32_t x, y, z, w; };

GL_HD int dp4a_s(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    return __dp4a((int)a, (int)b, c);
#else
    for (int i = 0; i < 4; ++i) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
#endif
}

// unsigned bytes (a) x signed bytes (b): lets the high nibbles stay in place (value 16*q) -- no shift
GL_HD int dp4a_us(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    for (int i = 0; i < 4; ++i) c += (int)(uint8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
#endif
}

GL_HD float bits_to_float(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

GL_HD float half_bits_to_float(uint16_t h) {
#if defined(__CUDA_ARCH__)
    return __half2float(__ushort_as_half(h));
#else
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF, out;
    if (exp == 0) {
        if (man == 0) out = sign;
        else { int e = -1; do { ++e; man <<= 1; } while (!(man & 0x400)); out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FF) << 13); }
    } else if (exp == 31) out = sign | 0x7F800000u | (man << 13);
    else out = sign | ((exp + 112) << 23) | (man << 13);
    float f; memcpy(&f, &out, 4); return f;
#endif
}

// One lane's activation slice (128 columns).  Unused members are dead-code-eliminated per type.
struct XUnit {
    uint32_t hi[32];   // int8x4 words, plane "hi" (v = 128*hi + lo when ABITS==16; v = hi when 8)
    uint32_t lo[32];
    float sx[4];       // fixed-point step of each 32-column block
    float sm[4];       // sx * sum(v) of each 32-column block      (Q4_K "mins" term)
    int s16[8];        // sum(v) of each 16-column group           (Q6_K -32 offset)
};

template <int ABITS> GL_HD int combine(int acc_hi, int acc_lo) { return ABITS == 16 ? acc_hi * 128 + acc_lo : acc_hi; }

// Snap 4 consecutive activations of a 32-column block to the fixed point: one hi word, one lo word, sum(v).
// inv = RANGE / amax of the WHOLE block (0 for an all-zero block).
// Spec (oracle/llama_oracle.py snap_i16 / snap_q8): sx = amax/RANGE, v = rint(x * (RANGE/amax)).
template <int ABITS>
GL_HD void snap4(const float* x, float inv, uint32_t* hi, uint32_t* lo, int* vsum) {
    uint32_t h = 0, l = 0;
    int s = 0;
    for (int b = 0; b < 4; ++b) {
#if defined(__CUDA_ARCH__)
        int v = __float2int_rn(x[b] * inv);
#else
        int v = (int)__builtin_rintf(x[b] * inv);
#endif
        s += v;
        if (ABITS == 16) {
            int lw = ((v + 64) & 127) - 64;
            int hh = (v - lw) >> 7;
            h |= (uint32_t)(hh & 0xFF) << (8 * b);
            l |= (uint32_t)(lw & 0xFF) << (8 * b);
        } else {
            h |= (uint32_t)(v & 0xFF) << (8 * b);
        }
    }
    *hi = h;
    *lo = l;
    *vsum = s;
}

template <int ABITS> GL_HD float snap_inv(float amax) {
    return amax > 0.f ? (ABITS == 16 ? ACT16_RANGE : ACT8_RANGE) / amax : 0.f;
}

// 16 consecutive activations (half of a 32-column block): 4 hi words, 4 lo words and sum(v).
template <int ABITS>
GL_HD void snap16(const float* x, float amax, uint32_t* hi4, uint32_t* lo4, int* vsum) {
    const float inv = snap_inv<ABITS>(amax);
    int s = 0;
    for (int w = 0; w < 4; ++w) {
        int vs;
        snap4<ABITS>(x + 4 * w, inv, hi4 + w, lo4 + w, &vs);
        s += vs;
    }
    *vsum = s;
}

// ---------------------------------------------------------------------------------------------
// Q4_K, native layout.  blk -> 144-B super-block; hb = which half (columns 128*hb .. +127).
// ---------------------------------------------------------------------------------------------
template <int ABITS>
GL_HD float unit_dot_q4k(const uint8_t* blk, int hb, const XUnit& x) {
    const U4 hdr = *reinterpret_cast<const U4*>(blk);
    const float d = half_bits_to_float((uint16_t)(hdr.x & 0xFFFF));
    const float dmin = half_bits_to_float((uint16_t)(hdr.x >> 16));
    const uint32_t s0 = hdr.y, s1 = hdr.z, s2 = hdr.w;
    // 6-bit scale / min of the unit's four 32-column sub-blocks, one per byte
    const uint32_t sc4 = hb ? ((s2 & 0x0F0F0F0Fu) | ((s0 >> 2) & 0x30303030u)) : (s0 & 0x3F3F3F3Fu);
    const uint32_t mn4 = hb ? (((s2 >> 4) & 0x0F0F0F0Fu) | ((s1 >> 2) & 0x30303030u)) : (s1 & 0x3F3F3F3Fu);
    const U4* q = reinterpret_cast<const U4*>(blk + 16 + 64 * hb);
    // all four 16-B chunks are requested before any arithmetic: one shared-memory latency per unit, not four
    const U4 qall[4] = {q[0], q[1], q[2], q[3]};
    float val = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {            // 32-byte chunk: low nibbles -> sub-block 2c, high -> 2c+1
        int ah = 0, al = 0, bh = 0, bl = 0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const U4 qq = qall[2 * c + v];
            const uint32_t w[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t lo4 = w[k] & 0x0F0F0F0Fu;
                const uint32_t hi4 = w[k] & 0xF0F0F0F0u;          // 16*q, consumed by the unsigned dp4a
                const int xa = 8 * (2 * c) + 4 * v + k, xb = 8 * (2 * c + 1) + 4 * v + k;
                ah = dp4a_s(lo4, x.hi[xa], ah);
                bh = dp4a_us(hi4, x.hi[xb], bh);
                if (ABITS == 16) {
                    al = dp4a_s(lo4, x.lo[xa], al);
                    bl = dp4a_us(hi4, x.lo[xb], bl);
                }
            }
        }
        const int sa = 2 * c, sb = 2 * c + 1;
        const float sca = (float)((sc4 >> (8 * sa)) & 0xFF), scb = (float)((sc4 >> (8 * sb)) & 0xFF) * 0.0625f;   // hi nibbles carry 16x
        const float mna = (float)((mn4 >> (8 * sa)) & 0xFF), mnb = (float)((mn4 >> (8 * sb)) & 0xFF);
        val += d * (sca * ((float)combine<ABITS>(ah, al) * x.sx[sa]) + scb * ((float)combine<ABITS>(bh, bl) * x.sx[sb]))
             - dmin * (mna * x.sm[sa] + mnb * x.sm[sb]);
    }
    return val;
}

// ---------------------------------------------------------------------------------------------
// Q6_K-T.  row -> start of the (transposed) row in shared memory; nb blocks, nu = 2*nb units.
// ---------------------------------------------------------------------------------------------
template <int ABITS>
GL_HD float unit_dot_q6k(const uint8_t* row, int nb, int u, const XUnit& x) {
    const int nu = 2 * nb;
    const uint8_t* qlp = row;
    const uint8_t* qhp = row + (size_t)nb * 128;
    const uint8_t* scp = row + (size_t)nb * 192 + (size_t)u * 8;
    const uint16_t dbits = *reinterpret_cast<const uint16_t*>(row + (size_t)nb * 208 + (size_t)(u >> 1) * 2);
    const float d = half_bits_to_float(dbits);
    const uint32_t scw[2] = {reinterpret_cast<const uint32_t*>(scp)[0], reinterpret_cast<const uint32_t*>(scp)[1]};
    U4 qh[2];
    qh[0] = *reinterpret_cast<const U4*>(qhp + ((size_t)0 * nu + u) * 16);
    qh[1] = *reinterpret_cast<const U4*>(qhp + ((size_t)1 * nu + u) * 16);
    U4 qlall[4];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) qlall[i4] = *reinterpret_cast<const U4*>(qlp + ((size_t)i4 * nu + u) * 16);
    float val = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        const U4 ql = qlall[i4];
        const uint32_t qlw[4] = {ql.x, ql.y, ql.z, ql.w};
        const U4 qhc = qh[i4 & 1];
        const uint32_t qhw[4] = {qhc.x, qhc.y, qhc.z, qhc.w};
        const int t = i4 >> 1;
        int ah = 0, al = 0, bh = 0, bl = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t a4 = (qlw[k] & 0x0F0F0F0Fu) | (((qhw[k] >> (2 * t)) & 0x03030303u) << 4);
            const uint32_t b4 = ((qlw[k] >> 4) & 0x0F0F0F0Fu) | (((qhw[k] >> (4 + 2 * t)) & 0x03030303u) << 4);
            const int xa = 4 * i4 + k, xb = 16 + 4 * i4 + k;
            ah = dp4a_s(a4, x.hi[xa], ah);
            bh = dp4a_s(b4, x.hi[xb], bh);
            if (ABITS == 16) {
                al = dp4a_s(a4, x.lo[xa], al);
                bl = dp4a_s(b4, x.lo[xb], bl);
            }
        }
        const int ga = i4, gb = 4 + i4;                       // 16-column groups inside the unit
        const float sca = (float)(int8_t)(scw[ga >> 2] >> (8 * (ga & 3)));
        const float scb = (float)(int8_t)(scw[gb >> 2] >> (8 * (gb & 3)));
        const int ia = combine<ABITS>(ah, al) - 32 * x.s16[ga];
        const int ib = combine<ABITS>(bh, bl) - 32 * x.s16[gb];
        val += sca * ((float)ia * x.sx[ga >> 1]) + scb * ((float)ib * x.sx[gb >> 1]);
    }
    return d * val;
}

// ---------------------------------------------------------------------------------------------
// Q8_0-T.  row -> transposed row; cols = K; nu = K/128.
// ---------------------------------------------------------------------------------------------
template <int ABITS>
GL_HD float unit_dot_q80(const uint8_t* row, int cols, int u, const XUnit& x) {
    const int nu = cols / UNIT_COLS;
    const uint32_t* dp = reinterpret_cast<const uint32_t*>(row + (size_t)cols + (size_t)u * 8);
    const uint32_t dw[2] = {dp[0], dp[1]};
    U4 qall[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qall[i] = *reinterpret_cast<const U4*>(row + ((size_t)i * nu + u) * 16);
    float val = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {            // 32-column block j of the unit
        int ah = 0, al = 0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int i = 2 * j + v;
            const U4 qq = qall[i];
            const uint32_t w[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ah = dp4a_s(w[k], x.hi[4 * i + k], ah);
                if (ABITS == 16) al = dp4a_s(w[k], x.lo[4 * i + k], al);
            }
        }
        const float dj = half_bits_to_float((uint16_t)(dw[j >> 1] >> (16 * (j & 1))));
        val += dj * ((float)combine<ABITS>(ah, al) * x.sx[j]);
    }
    return val;
}

// ---------------------------------------------------------------------------------------------
// Two rows at once.  Same arithmetic as the single-row functions, but the two rows advance word by word in
// lock-step so that every dp4a has 8-16 independent accumulators around it: the fixed-latency dependency
// stalls ("wait") that dominated the single-row schedule (profiles/r01_run5) are filled with the other row's work.
// ---------------------------------------------------------------------------------------------
template <int ABITS>
GL_HD void unit_dot2_q4k(const uint8_t* blk0, const uint8_t* blk1, int hb, const XUnit& x, float& out0, float& out1) {
    const uint8_t* blk[2] = {blk0, blk1};
    U4 hdr[2], qall[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        hdr[r] = *reinterpret_cast<const U4*>(blk[r]);
        const U4* q = reinterpret_cast<const U4*>(blk[r] + 16 + 64 * hb);
#pragma unroll
        for (int i = 0; i < 4; ++i) qall[r][i] = q[i];
    }
    float val[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        int ah[2] = {0, 0}, al[2] = {0, 0}, bh[2] = {0, 0}, bl[2] = {0, 0};
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const uint32_t w0[4] = {qall[0][2 * c + v].x, qall[0][2 * c + v].y, qall[0][2 * c + v].z, qall[0][2 * c + v].w};
            const uint32_t w1[4] = {qall[1][2 * c + v].x, qall[1][2 * c + v].y, qall[1][2 * c + v].z, qall[1][2 * c + v].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xa = 8 * (2 * c) + 4 * v + k, xb = 8 * (2 * c + 1) + 4 * v + k;
                const uint32_t lo0 = w0[k] & 0x0F0F0F0Fu, hi0 = w0[k] & 0xF0F0F0F0u;
                const uint32_t lo1 = w1[k] & 0x0F0F0F0Fu, hi1 = w1[k] & 0xF0F0F0F0u;
                ah[0] = dp4a_s(lo0, x.hi[xa], ah[0]);
                ah[1] = dp4a_s(lo1, x.hi[xa], ah[1]);
                bh[0] = dp4a_us(hi0, x.hi[xb], bh[0]);
                bh[1] = dp4a_us(hi1, x.hi[xb], bh[1]);
                if (ABITS == 16) {
                    al[0] = dp4a_s(lo0, x.lo[xa], al[0]);
                    al[1] = dp4a_s(lo1, x.lo[xa], al[1]);
                    bl[0] = dp4a_us(hi0, x.lo[xb], bl[0]);
                    bl[1] = dp4a_us(hi1, x.lo[xb], bl[1]);
                }
            }
        }
        const int sa = 2 * c, sb = 2 * c + 1;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t s0 = hdr[r].y, s1 = hdr[r].z, s2 = hdr[r].w;
            const uint32_t sc4 = hb ? ((s2 & 0x0F0F0F0Fu) | ((s0 >> 2) & 0x30303030u)) : (s0 & 0x3F3F3F3Fu);
            const uint32_t mn4 = hb ? (((s2 >> 4) & 0x0F0F0F0Fu) | ((s1 >> 2) & 0x30303030u)) : (s1 & 0x3F3F3F3Fu);
            const float d = half_bits_to_float((uint16_t)(hdr[r].x & 0xFFFF));
            const float dmin = half_bits_to_float((uint16_t)(hdr[r].x >> 16));
            const float sca = (float)((sc4 >> (8 * sa)) & 0xFF), scb = (float)((sc4 >> (8 * sb)) & 0xFF) * 0.0625f;
            const float mna = (float)((mn4 >> (8 * sa)) & 0xFF), mnb = (float)((mn4 >> (8 * sb)) & 0xFF);
            val[r] += d * (sca * ((float)combine<ABITS>(ah[r], al[r]) * x.sx[sa]) + scb * ((float)combine<ABITS>(bh[r], bl[r]) * x.sx[sb]))
                    - dmin * (mna * x.sm[sa] + mnb * x.sm[sb]);
        }
    }
    out0 = val[0];
    out1 = val[1];
}

template <int ABITS>
GL_HD void unit_dot2_q6k(const uint8_t* row0, const uint8_t* row1, int nb, int u, const XUnit& x, float& out0, float& out1) {
    const int nu = 2 * nb;
    const uint8_t* row[2] = {row0, row1};
    U4 ql[2][4], qh[2][2];
    uint32_t scw[2][2];
    float d[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint8_t* qlp = row[r];
        const uint8_t* qhp = row[r] + (size_t)nb * 128;
        const uint8_t* scp = row[r] + (size_t)nb * 192 + (size_t)u * 8;
        d[r] = half_bits_to_float(*reinterpret_cast<const uint16_t*>(row[r] + (size_t)nb * 208 + (size_t)(u >> 1) * 2));
        scw[r][0] = reinterpret_cast<const uint32_t*>(scp)[0];
        scw[r][1] = reinterpret_cast<const uint32_t*>(scp)[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) ql[r][i] = *reinterpret_cast<const U4*>(qlp + ((size_t)i * nu + u) * 16);
        qh[r][0] = *reinterpret_cast<const U4*>(qhp + ((size_t)0 * nu + u) * 16);
        qh[r][1] = *reinterpret_cast<const U4*>(qhp + ((size_t)1 * nu + u) * 16);
    }
    float val[2] = {0.f, 0.f};
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        const int t = i4 >> 1;
        int ah[2] = {0, 0}, al[2] = {0, 0}, bh[2] = {0, 0}, bl[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xa = 4 * i4 + k, xb = 16 + 4 * i4 + k;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t qlw = k == 0 ? ql[r][i4].x : k == 1 ? ql[r][i4].y : k == 2 ? ql[r][i4].z : ql[r][i4].w;
                const U4& qq = qh[r][i4 & 1];
                const uint32_t qhw = k == 0 ? qq.x : k == 1 ? qq.y : k == 2 ? qq.z : qq.w;
                const uint32_t a4 = (qlw & 0x0F0F0F0Fu) | (((qhw >> (2 * t)) & 0x03030303u) << 4);
                const uint32_t b4 = ((qlw >> 4) & 0x0F0F0F0Fu) | (((qhw >> (4 + 2 * t)) & 0x03030303u) << 4);
                ah[r] = dp4a_s(a4, x.hi[xa], ah[r]);
                bh[r] = dp4a_s(b4, x.hi[xb], bh[r]);
                if (ABITS == 16) {
                    al[r] = dp4a_s(a4, x.lo[xa], al[r]);
                    bl[r] = dp4a_s(b4, x.lo[xb], bl[r]);
                }
            }
        }
        const int ga = i4, gb = 4 + i4;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float sca = (float)(int8_t)(scw[r][ga >> 2] >> (8 * (ga & 3)));
            const float scb = (float)(int8_t)(scw[r][gb >> 2] >> (8 * (gb & 3)));
            const int ia = combine<ABITS>(ah[r], al[r]) - 32 * x.s16[ga];
            const int ib = combine<ABITS>(bh[r], bl[r]) - 32 * x.s16[gb];
            val[r] += sca * ((float)ia * x.sx[ga >> 1]) + scb * ((float)ib * x.sx[gb >> 1]);
        }
    }
    out0 = d[0] * val[0];
    out1 = d[1] * val[1];
}

// ---------------------------------------------------------------------------------------------
// Four rows at once, activations read from the (swizzled) shared-memory planes instead of registers.
//
// Register budget is what limited occupancy: an XUnit is 64-80 registers per lane.  Here a lane keeps only the
// x words of the 32-byte chunk it is working on (8-16 registers) and amortises each x load over FOUR weight rows,
// which also gives every dp4a 16 independent accumulator chains around it.  ~100 registers -> two CTAs (16
// consumer warps) per SM.
//
// XPlanes: pointers to this lane's unit inside the planes written by the prologue:
//   hi/lo  -> 128-byte area of the unit; 16-B chunk j lives at physical chunk (j ^ sw)
//   sx, sm -> per-32-column scale / scale*sum(v) of the unit (4 floats each); s16 -> per-16-column sums (8 ints)
// ---------------------------------------------------------------------------------------------
struct XPlanes {
    const uint8_t* hi;
    const uint8_t* lo;
    const float* sx;
    const float* sm;
    const int* s16;
    int sw;          // u & 7
};

GL_HD U4 xchunk(const uint8_t* plane, int j, int sw) { return *reinterpret_cast<const U4*>(plane + ((j ^ sw) << 4)); }

// rows[r] -> the 144-B super-block of row r that contains this unit (blk base), r = 0..3
template <int ABITS>
GL_HD void quad_dot_q4k(const uint8_t* const* blk, int hb, const XPlanes& xp, float* out) {
    float val[4] = {0.f, 0.f, 0.f, 0.f};
    U4 hdr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) hdr[r] = *reinterpret_cast<const U4*>(blk[r]);
    const U4 sx4 = *reinterpret_cast<const U4*>(xp.sx);
    const U4 sm4 = *reinterpret_cast<const U4*>(xp.sm);
    const float sxf[4] = {bits_to_float(sx4.x), bits_to_float(sx4.y), bits_to_float(sx4.z), bits_to_float(sx4.w)};
    const float smf[4] = {bits_to_float(sm4.x), bits_to_float(sm4.y), bits_to_float(sm4.z), bits_to_float(sm4.w)};
#pragma unroll
    for (int c = 0; c < 2; ++c) {            // 32-byte chunk pair: low nibbles -> sub-block 2c, high nibbles -> 2c+1
        int ah[4] = {0, 0, 0, 0}, al[4] = {0, 0, 0, 0}, bh[4] = {0, 0, 0, 0}, bl[4] = {0, 0, 0, 0};
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            // x words of sub-block 2c (chunk 2*(2c)+v) and sub-block 2c+1 (chunk 2*(2c+1)+v) of the unit
            const U4 xah = xchunk(xp.hi, 4 * c + v, xp.sw), xbh = xchunk(xp.hi, 4 * c + 2 + v, xp.sw);
            U4 xal = {0, 0, 0, 0}, xbl = {0, 0, 0, 0};
            if (ABITS == 16) { xal = xchunk(xp.lo, 4 * c + v, xp.sw); xbl = xchunk(xp.lo, 4 * c + 2 + v, xp.sw); }
            const uint32_t xa_h[4] = {xah.x, xah.y, xah.z, xah.w}, xb_h[4] = {xbh.x, xbh.y, xbh.z, xbh.w};
            const uint32_t xa_l[4] = {xal.x, xal.y, xal.z, xal.w}, xb_l[4] = {xbl.x, xbl.y, xbl.z, xbl.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const U4 qq = *reinterpret_cast<const U4*>(blk[r] + 16 + 64 * hb + 16 * (2 * c + v));
                const uint32_t w[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t lo4 = w[k] & 0x0F0F0F0Fu, hi4 = w[k] & 0xF0F0F0F0u;
                    ah[r] = dp4a_s(lo4, xa_h[k], ah[r]);
                    bh[r] = dp4a_us(hi4, xb_h[k], bh[r]);
                    if (ABITS == 16) {
                        al[r] = dp4a_s(lo4, xa_l[k], al[r]);
                        bl[r] = dp4a_us(hi4, xb_l[k], bl[r]);
                    }
                }
            }
        }
        const int sa = 2 * c, sb = 2 * c + 1;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t s0 = hdr[r].y, s1 = hdr[r].z, s2 = hdr[r].w;
            const uint32_t sc4 = hb ? ((s2 & 0x0F0F0F0Fu) | ((s0 >> 2) & 0x30303030u)) : (s0 & 0x3F3F3F3Fu);
            const uint32_t mn4 = hb ? (((s2 >> 4) & 0x0F0F0F0Fu) | ((s1 >> 2) & 0x30303030u)) : (s1 & 0x3F3F3F3Fu);
            const float d = half_bits_to_float((uint16_t)(hdr[r].x & 0xFFFF));
            const float dmin = half_bits_to_float((uint16_t)(hdr[r].x >> 16));
            const float sca = (float)((sc4 >> (8 * sa)) & 0xFF), scb = (float)((sc4 >> (8 * sb)) & 0xFF) * 0.0625f;
            const float mna = (float)((mn4 >> (8 * sa)) & 0xFF), mnb = (float)((mn4 >> (8 * sb)) & 0xFF);
            val[r] += d * (sca * ((float)combine<ABITS>(ah[r], al[r]) * sxf[sa]) + scb * ((float)combine<ABITS>(bh[r], bl[r]) * sxf[sb]))
                    - dmin * (mna * smf[sa] + mnb * smf[sb]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = val[r];
}

// rows[r] -> start of the Q6_K-T row r
template <int ABITS>
GL_HD void quad_dot_q6k(const uint8_t* const* row, int nb, int u, const XPlanes& xp, float* out) {
    const int nu = 2 * nb;
    float val[4] = {0.f, 0.f, 0.f, 0.f};
    const U4 sx4 = *reinterpret_cast<const U4*>(xp.sx);
    const float sxf[4] = {bits_to_float(sx4.x), bits_to_float(sx4.y), bits_to_float(sx4.z), bits_to_float(sx4.w)};
    const U4 g0 = *reinterpret_cast<const U4*>(xp.s16), g1 = *reinterpret_cast<const U4*>(xp.s16 + 4);
    const int s16[8] = {(int)g0.x, (int)g0.y, (int)g0.z, (int)g0.w, (int)g1.x, (int)g1.y, (int)g1.z, (int)g1.w};
    uint32_t scw[4][2];
    float d[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint8_t* scp = row[r] + (size_t)nb * 192 + (size_t)u * 8;
        scw[r][0] = reinterpret_cast<const uint32_t*>(scp)[0];
        scw[r][1] = reinterpret_cast<const uint32_t*>(scp)[1];
        d[r] = half_bits_to_float(*reinterpret_cast<const uint16_t*>(row[r] + (size_t)nb * 208 + (size_t)(u >> 1) * 2));
    }
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        const int t = i4 >> 1;
        // x words of the A elements (columns 16*i4..) and B elements (columns 64+16*i4..): chunk i4 and chunk 4+i4
        const U4 xah = xchunk(xp.hi, i4, xp.sw), xbh = xchunk(xp.hi, 4 + i4, xp.sw);
        U4 xal = {0, 0, 0, 0}, xbl = {0, 0, 0, 0};
        if (ABITS == 16) { xal = xchunk(xp.lo, i4, xp.sw); xbl = xchunk(xp.lo, 4 + i4, xp.sw); }
        const uint32_t xa_h[4] = {xah.x, xah.y, xah.z, xah.w}, xb_h[4] = {xbh.x, xbh.y, xbh.z, xbh.w};
        const uint32_t xa_l[4] = {xal.x, xal.y, xal.z, xal.w}, xb_l[4] = {xbl.x, xbl.y, xbl.z, xbl.w};
        const int ga = i4, gb = 4 + i4;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const U4 ql = *reinterpret_cast<const U4*>(row[r] + ((size_t)i4 * nu + u) * 16);
            const U4 qh = *reinterpret_cast<const U4*>(row[r] + (size_t)nb * 128 + ((size_t)(i4 & 1) * nu + u) * 16);
            const uint32_t qlw[4] = {ql.x, ql.y, ql.z, ql.w}, qhw[4] = {qh.x, qh.y, qh.z, qh.w};
            int ah = 0, al = 0, bh = 0, bl = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a4 = (qlw[k] & 0x0F0F0F0Fu) | (((qhw[k] >> (2 * t)) & 0x03030303u) << 4);
                const uint32_t b4 = ((qlw[k] >> 4) & 0x0F0F0F0Fu) | (((qhw[k] >> (4 + 2 * t)) & 0x03030303u) << 4);
                ah = dp4a_s(a4, xa_h[k], ah);
                bh = dp4a_s(b4, xb_h[k], bh);
                if (ABITS == 16) {
                    al = dp4a_s(a4, xa_l[k], al);
                    bl = dp4a_s(b4, xb_l[k], bl);
                }
            }
            const float sca = (float)(int8_t)(scw[r][ga >> 2] >> (8 * (ga & 3)));
            const float scb = (float)(int8_t)(scw[r][gb >> 2] >> (8 * (gb & 3)));
            const int ia = combine<ABITS>(ah, al) - 32 * s16[ga];
            const int ib = combine<ABITS>(bh, bl) - 32 * s16[gb];
            val[r] += sca * ((float)ia * sxf[ga >> 1]) + scb * ((float)ib * sxf[gb >> 1]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = d[r] * val[r];
}

// rows[r] -> start of the Q8_0-T row r
template <int ABITS>
GL_HD void quad_dot_q80(const uint8_t* const* row, int cols, int u, const XPlanes& xp, float* out) {
    const int nu = cols / UNIT_COLS;
    float val[4] = {0.f, 0.f, 0.f, 0.f};
    const U4 sx4 = *reinterpret_cast<const U4*>(xp.sx);
    const float sxf[4] = {bits_to_float(sx4.x), bits_to_float(sx4.y), bits_to_float(sx4.z), bits_to_float(sx4.w)};
    uint32_t dw[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t* dp = reinterpret_cast<const uint32_t*>(row[r] + (size_t)cols + (size_t)u * 8);
        dw[r][0] = dp[0];
        dw[r][1] = dp[1];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {            // 32-column block j of the unit = chunks 2j, 2j+1
        int ah[4] = {0, 0, 0, 0}, al[4] = {0, 0, 0, 0};
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int i = 2 * j + v;
            const U4 xh = xchunk(xp.hi, i, xp.sw);
            U4 xl = {0, 0, 0, 0};
            if (ABITS == 16) xl = xchunk(xp.lo, i, xp.sw);
            const uint32_t x_h[4] = {xh.x, xh.y, xh.z, xh.w}, x_l[4] = {xl.x, xl.y, xl.z, xl.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const U4 qq = *reinterpret_cast<const U4*>(row[r] + ((size_t)i * nu + u) * 16);
                const uint32_t w[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ah[r] = dp4a_s(w[k], x_h[k], ah[r]);
                    if (ABITS == 16) al[r] = dp4a_s(w[k], x_l[k], al[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float dj = half_bits_to_float((uint16_t)(dw[r][j >> 1] >> (16 * (j & 1))));
            val[r] += dj * ((float)combine<ABITS>(ah[r], al[r]) * sxf[j]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = val[r];
}

// ---------------------------------------------------------------------------------------------
// host-side row repackers (loader) -- GGUF row -> engine row.  dst has engine_row_stride bytes.
// ---------------------------------------------------------------------------------------------
inline size_t align16(size_t n) { return (n + 15) & ~(size_t)15; }

inline void repack_row_q6k(const uint8_t* src, uint8_t* dst, int nb) {
    const int nu = 2 * nb;
    for (int b = 0; b < nb; ++b) {
        const uint8_t* blk = src + (size_t)b * 210;
        for (int h = 0; h < 2; ++h) {
            const int u = 2 * b + h;
            for (int i4 = 0; i4 < 4; ++i4) memcpy(dst + ((size_t)i4 * nu + u) * 16, blk + h * 64 + 16 * i4, 16);
            for (int i2 = 0; i2 < 2; ++i2) memcpy(dst + (size_t)nb * 128 + ((size_t)i2 * nu + u) * 16, blk + 128 + h * 32 + 16 * i2, 16);
            memcpy(dst + (size_t)nb * 192 + (size_t)u * 8, blk + 192 + 8 * h, 8);
        }
        memcpy(dst + (size_t)nb * 208 + (size_t)b * 2, blk + 208, 2);
    }
}

inline void repack_row_q80(const uint8_t* src, uint8_t* dst, int cols) {
    const int nu = cols / UNIT_COLS;
    for (int u = 0; u < nu; ++u) {
        for (int j = 0; j < 4; ++j) {
            const uint8_t* blk = src + (size_t)(4 * u + j) * 34;
            memcpy(dst + ((size_t)(2 * j) * nu + u) * 16, blk + 2, 16);
            memcpy(dst + ((size_t)(2 * j + 1) * nu + u) * 16, blk + 18, 16);
            memcpy(dst + (size_t)cols + (size_t)u * 8 + 2 * j, blk, 2);
        }
    }
}

}  // namespace gl
