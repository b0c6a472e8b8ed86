// Per-lane block-decode arithmetic of the decode GEMV and the engine's HBM row layouts.
//
// Everything here is __host__ __device__ so that tests/hostcheck can run the exact lane
// program on the CPU (layout / bit-twiddling bugs are found without a GPU).  The CUDA kernels
// call these same functions; nothing in the product path runs them on the host.
//
// Vocabulary
//   block   : 256 consecutive columns of one weight row (one K-quant super-block, or 8 Q8_0 blocks)
//   unit    : 128 consecutive columns = half a block: the slice of a row one lane owns
//   K-seg   : a run of <= 16 blocks (<= 32 units) of a row -- what one warp covers in one pass.  Rows are
//             cut into nks equal K-segments (ksplit()); a segment of a row is one contiguous byte range
//   item    : R rows (R = 4, or 2 when four do not fit a ring slot) x all K-segments; the unit of work of
//             one consumer warp.  One ring slot holds the R row-segments of one (item, K-segment).
//
// Engine layouts in HBM (per-row permutations of the GGUF rows).  A matrix is stored as
// [tile of R rows][K-segment][row in tile][segment], so the R row-segments of one ring slot are one contiguous
// bulk copy.  With n = units per K-segment, a segment (padded to 16 B, the TMA granularity) is:
//   Q4_K   : native GGUF (144-B super-blocks), a segment is n/2 consecutive super-blocks.  Conflict-free
//            as is because 144 = 128 + 16 rotates consecutive blocks across the 16-B shared-memory slots.
//   Q6_K-T : [ql : 4 x n x 16 B, chunk(i4,u) at (i4*n+u)*16 ]    bytes h*64+16*i4.. of block u>>1, h = u&1
//            [qh : 2 x n x 16 B, chunk(i2,u) at (i2*n+u)*16 ]    bytes h*32+16*i2..
//            [sc : n x 8 B ]                                       scales[8h..8h+8)
//            [d  : n/2 x 2 B ]
//   Q8_0-T : [qs : 8 x n x 16 B, chunk(i,u) at (i*n+u)*16 ] [d : n x 8 B]       (4 blocks of 32 per unit)
//   so that lane u reading chunk i is part of a contiguous 512-B warp access (GGUF's 210-B / 34-B blocks
//   are only 2-byte aligned and cannot be read with wide loads).
//
// Activations: 15-bit (or 8-bit) fixed point per 32-column block, x ~= sx * (128*hi + lo), kept in shared
// memory as int8 planes ("XPlanes") so that every product is an exact integer dp4a.
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

// four int8 lanes -> one word (two's complement bytes, lane 0 in the low byte)
GL_HD uint32_t pack4_s8(int a0, int a1, int a2, int a3) {
#if defined(__CUDA_ARCH__)
    const uint32_t p01 = __byte_perm((uint32_t)a0, (uint32_t)a1, 0x0040);
    const uint32_t p23 = __byte_perm((uint32_t)a2, (uint32_t)a3, 0x0040);
    return __byte_perm(p01, p23, 0x5410);
#else
    return ((uint32_t)a0 & 0xFF) | (((uint32_t)a1 & 0xFF) << 8) | (((uint32_t)a2 & 0xFF) << 16) | (((uint32_t)a3 & 0xFF) << 24);
#endif
}

// Snap 4 consecutive activations of a 32-column block to the fixed point: one hi word, one lo word, sum(v).
// inv = RANGE / amax of the WHOLE block (0 for an all-zero block).
// Spec (oracle/llama_oracle.py snap_i16 / snap_q8): sx = amax/RANGE, v = rint(x * (RANGE/amax)); the 15-bit value is
// split as v = 128*hi + lo with lo in [-64, 63]  (hi = floor((v + 64) / 128)).
template <int ABITS>
GL_HD void snap4(const float* x, float inv, uint32_t* hi, uint32_t* lo, int* vsum) {
    int v[4], h[4], l[4];
    for (int b = 0; b < 4; ++b) {
#if defined(__CUDA_ARCH__)
        v[b] = __float2int_rn(x[b] * inv);
#else
        v[b] = (int)__builtin_rintf(x[b] * inv);
#endif
        if (ABITS == 16) {
            h[b] = (v[b] + 64) >> 7;
            l[b] = v[b] - (h[b] << 7);
        } else {
            h[b] = v[b];
            l[b] = 0;
        }
    }
    *hi = pack4_s8(h[0], h[1], h[2], h[3]);
    *lo = ABITS == 16 ? pack4_s8(l[0], l[1], l[2], l[3]) : 0u;
    *vsum = (v[0] + v[1]) + (v[2] + v[3]);
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
// K-segmentation and engine row geometry
// ---------------------------------------------------------------------------------------------
constexpr int MAX_KSEG_BLOCKS = 16;       // 32 units: one lane per unit
constexpr int MAX_KSEGS = 8;              // K <= 32768

struct KSplit { int nks; int seg_nb; };   // K-segments per row, 256-column blocks per segment

// cols % 256 == 0.  Equal segments of <= 16 blocks; nks == 0 when the row cannot be cut that way.
GL_HD KSplit ksplit(int cols) {
    KSplit k{0, 0};
    if (cols <= 0 || (cols & 255)) return k;
    const int nb = cols >> 8;
    int nks = (nb + MAX_KSEG_BLOCKS - 1) / MAX_KSEG_BLOCKS;
    while (nks <= MAX_KSEGS && (nb % nks)) ++nks;
    if (nks > MAX_KSEGS) return k;
    k.nks = nks;
    k.seg_nb = nb / nks;
    return k;
}

// ggml type ids (gguf_file.h): Q8_0 = 8, Q4_K = 12, Q6_K = 14.  Bytes of 256 columns.
GL_HD int quant_block_bytes(int type) { return type == 12 ? 144 : type == 14 ? 210 : type == 8 ? 272 : 0; }
GL_HD int kseg_bytes(int type, int seg_nb) { return (seg_nb * quant_block_bytes(type) + 15) & ~15; }
// Matrix layout: rows are grouped into TILES of tile_rows rows (the rows of one GEMV item); inside a tile the
// K-segments are the outer index, so the tile_rows segments a ring slot holds are ONE contiguous byte range:
//      [tile][K-segment][row in tile][segment bytes]
// With one K-segment (K <= 4096) this is plain row-major whatever tile_rows is.
GL_HD size_t engine_seg_offset(int seg_bytes, int nks, int tile_rows, int r, int ks) {
    return (size_t)(r / tile_rows) * ((size_t)tile_rows * nks * seg_bytes) + (size_t)ks * tile_rows * seg_bytes +
           (size_t)(r % tile_rows) * seg_bytes;
}
GL_HD size_t engine_matrix_bytes(int type, int rows, int cols, int tile_rows) {
    const KSplit k = ksplit(cols);
    return (size_t)((rows + tile_rows - 1) / tile_rows) * tile_rows * k.nks * kseg_bytes(type, k.seg_nb);
}
// rows of one non-paired item of a type (gemv_plan): the tile size a matrix is stored with unless it is a gate/up pair
GL_HD int item_rows(int type) { return type == 12 ? 4 : 2; }

// ---------------------------------------------------------------------------------------------
// Activation planes in shared memory (written by the GEMV prologue)
//   hi/lo  -> 128-byte area of the unit; 16-B chunk j lives at physical chunk (j ^ sw)
//   sx, sm -> per-32-column scale / scale*sum(v) of the unit (4 floats each); s16 -> per-16-column sums (8 ints)
// ---------------------------------------------------------------------------------------------
struct XPlanes {
    const uint8_t* hi;
    const uint8_t* lo;
    const float* sx;
    const float* sm;
    const int* s16;
    int sw;          // global unit index & 7
};

GL_HD U4 xchunk(const uint8_t* plane, int j, int sw) { return *reinterpret_cast<const U4*>(plane + ((j ^ sw) << 4)); }
GL_HD U4 ld16(const uint8_t* p) { return *reinterpret_cast<const U4*>(p); }

// ---------------------------------------------------------------------------------------------
// Item dot products: R rows at once against one unit (128 columns) of x.
//
//   seg  -> the slot: R row-segments back to back, row j at seg + j*seg_bytes
//   l    -> this lane's unit inside the K-segment (0 .. n-1)
//   xp   -> the matching unit of the activation planes
//   acc  -> R running row sums (accumulated, not overwritten)
//
// Each x word is loaded once and used for all R rows; with R = 4 every dp4a has 16 independent accumulator
// chains around it, which is what hides the fixed 4-cycle dependent-issue latency.
// ---------------------------------------------------------------------------------------------
template <int ABITS, int R>
GL_HD void item_dot_q4k(const uint8_t* seg, int seg_bytes, int l, const XPlanes& xp, float* acc) {
    const int hb = l & 1;
    const uint8_t* blk = seg + (size_t)(l >> 1) * 144;
    U4 hdr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) hdr[r] = ld16(blk + (size_t)r * seg_bytes);
    const U4 sx4 = ld16(reinterpret_cast<const uint8_t*>(xp.sx));
    const U4 sm4 = ld16(reinterpret_cast<const uint8_t*>(xp.sm));
    const float sxf[4] = {bits_to_float(sx4.x), bits_to_float(sx4.y), bits_to_float(sx4.z), bits_to_float(sx4.w)};
    const float smf[4] = {bits_to_float(sm4.x), bits_to_float(sm4.y), bits_to_float(sm4.z), bits_to_float(sm4.w)};
    // 6-bit scale / min of the unit's four 32-column sub-blocks, one per byte
    uint32_t sc4[R], mn4[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t s0 = hdr[r].y, s1 = hdr[r].z, s2 = hdr[r].w;
        sc4[r] = hb ? ((s2 & 0x0F0F0F0Fu) | ((s0 >> 2) & 0x30303030u)) : (s0 & 0x3F3F3F3Fu);
        mn4[r] = hb ? (((s2 >> 4) & 0x0F0F0F0Fu) | ((s1 >> 2) & 0x30303030u)) : (s1 & 0x3F3F3F3Fu);
    }
    float vd[R], vm[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { vd[r] = 0.f; vm[r] = 0.f; }
    const uint8_t* q0 = blk + 16 + 64 * hb;
#pragma unroll
    for (int c = 0; c < 2; ++c) {            // 32-byte chunk pair: low nibbles -> sub-block 2c, high nibbles -> 2c+1
        int ah[R], al[R], bh[R], bl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { ah[r] = al[r] = bh[r] = bl[r] = 0; }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            // x words of sub-block 2c (chunk 2*(2c)+v) and sub-block 2c+1 (chunk 2*(2c+1)+v) of the unit
            const U4 xah = xchunk(xp.hi, 4 * c + v, xp.sw), xbh = xchunk(xp.hi, 4 * c + 2 + v, xp.sw);
            U4 xal = {0, 0, 0, 0}, xbl = {0, 0, 0, 0};
            if (ABITS == 16) { xal = xchunk(xp.lo, 4 * c + v, xp.sw); xbl = xchunk(xp.lo, 4 * c + 2 + v, xp.sw); }
            const uint32_t xa_h[4] = {xah.x, xah.y, xah.z, xah.w}, xb_h[4] = {xbh.x, xbh.y, xbh.z, xbh.w};
            const uint32_t xa_l[4] = {xal.x, xal.y, xal.z, xal.w}, xb_l[4] = {xbl.x, xbl.y, xbl.z, xbl.w};
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const U4 qq = ld16(q0 + (size_t)r * seg_bytes + 16 * (2 * c + v));
                const uint32_t w[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t lo4 = w[k] & 0x0F0F0F0Fu, hi4 = w[k] & 0xF0F0F0F0u;     // hi4 = 16*q, unsigned dp4a
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
        const float sxa = sxf[sa], sxb = sxf[sb] * 0.0625f;        // high nibbles carry 16x
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float sca = (float)((sc4[r] >> (8 * sa)) & 0xFF), scb = (float)((sc4[r] >> (8 * sb)) & 0xFF);
            const float mna = (float)((mn4[r] >> (8 * sa)) & 0xFF), mnb = (float)((mn4[r] >> (8 * sb)) & 0xFF);
            vd[r] += sca * ((float)combine<ABITS>(ah[r], al[r]) * sxa) + scb * ((float)combine<ABITS>(bh[r], bl[r]) * sxb);
            vm[r] += mna * smf[sa] + mnb * smf[sb];
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float d = half_bits_to_float((uint16_t)(hdr[r].x & 0xFFFF));
        const float dmin = half_bits_to_float((uint16_t)(hdr[r].x >> 16));
        acc[r] += d * vd[r] - dmin * vm[r];
    }
}

// n = units in the K-segment (2 * seg_nb)
template <int ABITS, int R>
GL_HD void item_dot_q6k(const uint8_t* seg, int seg_bytes, int n, int l, const XPlanes& xp, float* acc) {
    const U4 sx4 = ld16(reinterpret_cast<const uint8_t*>(xp.sx));
    const float sxf[4] = {bits_to_float(sx4.x), bits_to_float(sx4.y), bits_to_float(sx4.z), bits_to_float(sx4.w)};
    const U4 g0 = ld16(reinterpret_cast<const uint8_t*>(xp.s16)), g1 = ld16(reinterpret_cast<const uint8_t*>(xp.s16 + 4));
    const int s16[8] = {(int)g0.x, (int)g0.y, (int)g0.z, (int)g0.w, (int)g1.x, (int)g1.y, (int)g1.z, (int)g1.w};
    uint32_t scw[R][2];
    float d[R], val[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint8_t* row = seg + (size_t)r * seg_bytes;
        const uint32_t* scp = reinterpret_cast<const uint32_t*>(row + (size_t)n * 96 + (size_t)l * 8);
        scw[r][0] = scp[0];
        scw[r][1] = scp[1];
        d[r] = half_bits_to_float(*reinterpret_cast<const uint16_t*>(row + (size_t)n * 104 + (size_t)(l >> 1) * 2));
        val[r] = 0.f;
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
        const int ga = i4, gb = 4 + i4;       // 16-column groups inside the unit
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint8_t* row = seg + (size_t)r * seg_bytes;
            const U4 ql = ld16(row + ((size_t)i4 * n + l) * 16);
            const U4 qh = ld16(row + (size_t)n * 64 + ((size_t)(i4 & 1) * n + l) * 16);
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
    for (int r = 0; r < R; ++r) acc[r] += d[r] * val[r];
}

template <int ABITS, int R>
GL_HD void item_dot_q80(const uint8_t* seg, int seg_bytes, int n, int l, const XPlanes& xp, float* acc) {
    const U4 sx4 = ld16(reinterpret_cast<const uint8_t*>(xp.sx));
    const float sxf[4] = {bits_to_float(sx4.x), bits_to_float(sx4.y), bits_to_float(sx4.z), bits_to_float(sx4.w)};
    uint32_t dw[R][2];
    float val[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t* dp = reinterpret_cast<const uint32_t*>(seg + (size_t)r * seg_bytes + (size_t)n * 128 + (size_t)l * 8);
        dw[r][0] = dp[0];
        dw[r][1] = dp[1];
        val[r] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {            // 32-column block j of the unit = chunks 2j, 2j+1
        int ah[R], al[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { ah[r] = 0; al[r] = 0; }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int i = 2 * j + v;
            const U4 xh = xchunk(xp.hi, i, xp.sw);
            U4 xl = {0, 0, 0, 0};
            if (ABITS == 16) xl = xchunk(xp.lo, i, xp.sw);
            const uint32_t x_h[4] = {xh.x, xh.y, xh.z, xh.w}, x_l[4] = {xl.x, xl.y, xl.z, xl.w};
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const U4 qq = ld16(seg + (size_t)r * seg_bytes + ((size_t)i * n + l) * 16);
                const uint32_t w[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ah[r] = dp4a_s(w[k], x_h[k], ah[r]);
                    if (ABITS == 16) al[r] = dp4a_s(w[k], x_l[k], al[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float dj = half_bits_to_float((uint16_t)(dw[r][j >> 1] >> (16 * (j & 1))));
            val[r] += dj * ((float)combine<ABITS>(ah[r], al[r]) * sxf[j]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] += val[r];
}

// ---------------------------------------------------------------------------------------------
// host-side repackers (loader) -- GGUF rows -> engine matrix
// ---------------------------------------------------------------------------------------------
inline size_t align16(size_t n) { return (n + 15) & ~(size_t)15; }

// K-segment ks of one GGUF row -> engine segment (kseg_bytes() bytes at dst)
inline void repack_seg_q4k(const uint8_t* src_row, uint8_t* dst, int cols, int ks) {
    const KSplit k = ksplit(cols);
    memcpy(dst, src_row + (size_t)ks * k.seg_nb * 144, (size_t)k.seg_nb * 144);
}

inline void repack_seg_q6k(const uint8_t* src_row, uint8_t* dst, int cols, int ks) {
    const KSplit k = ksplit(cols);
    const int n = 2 * k.seg_nb, sb = kseg_bytes(14, k.seg_nb);
    memset(dst, 0, (size_t)sb);
    for (int bl = 0; bl < k.seg_nb; ++bl) {
        const uint8_t* blk = src_row + (size_t)(ks * k.seg_nb + bl) * 210;
        for (int h = 0; h < 2; ++h) {
            const int u = 2 * bl + h;
            for (int i4 = 0; i4 < 4; ++i4) memcpy(dst + ((size_t)i4 * n + u) * 16, blk + h * 64 + 16 * i4, 16);
            for (int i2 = 0; i2 < 2; ++i2) memcpy(dst + (size_t)n * 64 + ((size_t)i2 * n + u) * 16, blk + 128 + h * 32 + 16 * i2, 16);
            memcpy(dst + (size_t)n * 96 + (size_t)u * 8, blk + 192 + 8 * h, 8);
        }
        memcpy(dst + (size_t)n * 104 + (size_t)bl * 2, blk + 208, 2);
    }
}

inline void repack_seg_q80(const uint8_t* src_row, uint8_t* dst, int cols, int ks) {
    const KSplit k = ksplit(cols);
    const int n = 2 * k.seg_nb;
    for (int u = 0; u < n; ++u) {
        const int ug = ks * n + u;
        for (int j = 0; j < 4; ++j) {
            const uint8_t* blk = src_row + (size_t)(4 * ug + j) * 34;
            memcpy(dst + ((size_t)(2 * j) * n + u) * 16, blk + 2, 16);
            memcpy(dst + ((size_t)(2 * j + 1) * n + u) * 16, blk + 18, 16);
            memcpy(dst + (size_t)n * 128 + (size_t)u * 8 + 2 * j, blk, 2);
        }
    }
}

// one GGUF row -> its segments inside the engine matrix `mat` (engine_matrix_bytes() bytes, zero-initialised)
inline void repack_row(int type, const uint8_t* src_row, uint8_t* mat, int cols, int tile_rows, int r) {
    const KSplit k = ksplit(cols);
    const int sb = kseg_bytes(type, k.seg_nb);
    for (int ks = 0; ks < k.nks; ++ks) {
        uint8_t* dst = mat + engine_seg_offset(sb, k.nks, tile_rows, r, ks);
        if (type == 12) repack_seg_q4k(src_row, dst, cols, ks);
        else if (type == 14) repack_seg_q6k(src_row, dst, cols, ks);
        else repack_seg_q80(src_row, dst, cols, ks);
    }
}

// element c of a NATIVE GGUF Q4_K row
GL_HD float dequant_native_q4k(const uint8_t* row, int c) {
    const uint8_t* b = row + (size_t)(c >> 8) * 144;
    const int e = c & 255, sub = e >> 5, l = e & 31;
    const float d = half_bits_to_float(*reinterpret_cast<const uint16_t*>(b));
    const float dmin = half_bits_to_float(*reinterpret_cast<const uint16_t*>(b + 2));
    const uint8_t* sc = b + 4;
    int s, m;
    if (sub < 4) { s = sc[sub] & 63; m = sc[4 + sub] & 63; }
    else { s = (sc[4 + sub] & 0xF) | ((sc[sub - 4] >> 6) << 4); m = (sc[4 + sub] >> 4) | ((sc[sub] >> 6) << 4); }
    const uint8_t qb = b[16 + (sub >> 1) * 32 + l];
    const int q = (sub & 1) ? (qb >> 4) : (qb & 0xF);
    return d * (float)s * (float)q - dmin * (float)m;
}

// element (r, c) of an ENGINE-layout quantised matrix (load-time 16-bit copy for the batched prefill; host checks)
GL_HD float dequant_engine_quant(const uint8_t* mat, int type, int cols, int tile_rows, int r, int c) {
    const KSplit ks = ksplit(cols);
    const int n = 2 * ks.seg_nb;
    const int ug = c >> 7, w = c & 127;
    const uint8_t* seg = mat + engine_seg_offset(kseg_bytes(type, ks.seg_nb), ks.nks, tile_rows, r, ug / n);
    const int u = ug % n;
    if (type == 12) return dequant_native_q4k(seg, (u << 7) + w);      // a Q4_K segment is a run of native super-blocks
    if (type == 14) {
        const int i = w & 63, s = w >> 6, j = w & 31, t = w >> 5;
        const int qlv = (seg[((size_t)(i >> 4) * n + u) * 16 + (i & 15)] >> (4 * s)) & 0xF;
        const int qhv = (seg[(size_t)n * 64 + ((size_t)(j >> 4) * n + u) * 16 + (j & 15)] >> (2 * t)) & 3;
        const float d = half_bits_to_float(*reinterpret_cast<const uint16_t*>(seg + (size_t)n * 104 + (size_t)(u >> 1) * 2));
        const float sc = (float)(int8_t)seg[(size_t)n * 96 + (size_t)u * 8 + (w >> 4)];
        return d * sc * (float)((qlv | (qhv << 4)) - 32);
    }
    if (type == 8) {
        const float d = half_bits_to_float(*reinterpret_cast<const uint16_t*>(seg + (size_t)n * 128 + (size_t)u * 8 + 2 * (w >> 5)));
        return d * (float)(int8_t)seg[((size_t)(w >> 4) * n + u) * 16 + (w & 15)];
    }
    return 0.f;
}

}  // namespace gl
