// Batched decode GEMM on QUANTISED weights: C[b][n] = sum_k act[b][k] * dequant(W)[n][k] for the B <= 64 sequences of one batched
// decode step (continuous batching).  The weights stay in their GGUF bit budget in HBM (Q4_K 4.5 /
// Q6_K 6.5625 bits per weight) and are read ONCE per step for all sequences; they become fp16 tensor-core operands tile by
// tile inside the kernel.  Reference side: the decode loop of several concurrent requests inside Ollama behind
// OllamaService.generate*Response (/root/reference/client/src/services/OllamaService.ts:142-145, 235-237) once the worker holds
// more than one job (WorkerClientService.ts:500-505; MAX_CONCURRENT_JOBS_PER_WORKER, server/src/config/index.ts:31).
//
// Bound: HBM.  A step streams the model's 4.6 GB once; at B = 32 the tensor work is 128 x 32 x 64 MACs per 16 KB of weights
// (16 cycles of tcgen05 against ~800 cycles of HBM time per qtile and SM), so the tensor pipe idles and the CUDA cores'
// job -- ~2.3 lane-operations per weight to unpack nibbles into fp16 -- is what has to keep up with the memory pipe.
//
// Shape of the kernel (persistent, one CTA per SM, 14 warps, hand-written PTX; layouts and the per-thread unpack program in
// qgemm_layout.h, which the CPU suite runs bit for bit):
//   the weights of a GEMM are a stream of QTILES (128 rows x 256 columns, one contiguous 18 / 26 KB range each); the U = tiles x
//   K-blocks qtiles are dealt to the G CTAs as contiguous ranges [c U / G, (c+1) U / G) ("stream-K": perfect balance for every
//   shape -- 32 tiles x 16 blocks on 148 SMs as well as 1002 x 16 -- and every CTA streams ONE contiguous byte range);
//   warp 0 / one lane : producer -- per qtile one 1-D TMA bulk copy of the raw bytes plus four 2-D TMA boxes of the activations
//                       [B x 64] fp16 (L2-resident) into a 2-3 stage ring;
//   warps 2..9        : unpack -- thread (row, half) turns 128 columns of its row into sixteen 16-byte chunks of the four
//                       128-byte-swizzled operand tiles [128 x 64] fp16 of the qtile (conflict-free 128-bit loads and stores),
//                       fence.proxy.async, mbarrier arrive;
//   warp 1 / one lane : MMA issuer -- per K-step four tcgen05.mma.cta_group::1.kind::f16 (M 128 = weight rows, N = B, K 16) on
//                       shared-memory descriptors, accumulator [128 lanes x B columns] fp32 in TENSOR MEMORY, double-buffered
//                       across output tiles; tcgen05.commit hands operand tiles / stages / accumulators on;
//   warps 10..13      : epilogue -- tcgen05.ld, then either the fused epilogue (fp32 store, residual add, SiLU*mul -> fp16) or,
//                       for an output tile whose K range is shared with neighbouring CTAs, a partial to scratch + atomic
//                       ticket; the last contributor adds the partials IN CTA ORDER (deterministic) and runs the epilogue.
// SASS to look for: UTCHMMA (tcgen05.mma), UBLKCP (1-D bulk copy), UTMALDG (2-D TMA), LDTM (tcgen05.ld), UTCBAR (commit).
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "qgemm.h"
#include "qgemm_layout.h"

namespace gl {

namespace {

constexpr int QG_THREADS = 15 * 32;
constexpr int QG_A_SLOTS = 6;                        // operand-tile ring: 1.5 qtiles, so the unpack warps never wait for the MMAs of the qtile before
constexpr int QG_ACT_UNITS = 2;                      // activation ring: the four [NB x 64] boxes of two qtiles
constexpr size_t QG_SMEM = 227 * 1024;               // the whole opt-in shared memory; the raw ring takes what the other two leave

template <int NB> struct QCfg {
    static constexpr int B_TILE = NB * 128;          // activations [NB rows x 64] fp16, 128-byte swizzle
    static constexpr int ACT_BYTES = QG_ACT_UNITS * 4 * B_TILE;
    static constexpr int A_BYTES = QG_A_SLOTS * QG_A_TILE_BYTES;
    static constexpr int BAR_BYTES = 512;
    static constexpr int RAW_BUDGET = (int)QG_SMEM - 1024 /* alignment */ - BAR_BYTES - ACT_BYTES - A_BYTES;
    static constexpr int TMEM_COLS = 2 * NB < 32 ? 32 : 2 * NB;
};
constexpr int QG_MAX_RAW_STAGES = 6;

struct QParams {
    CUtensorMap tb;              // activations: dims {K, rows_alloc}, box {64, NB}, 128-byte swizzle
    const uint8_t* w;
    const uint64_t* tile_off;    // per-tile tables: only when the GEMM's tiles are not "type0 up to tile_split, then type1"
    const uint8_t* tile_type;
    unsigned* counters;
    float* partial;              // [grid][2][NB * 128]
    void* c;
    int ldc, epi, n, n_tiles, nkb;
    int type0, type1, tile_split;          // tiles [0, tile_split) are type0, the rest type1 (Q | K | V with a Q6_K V); tile_split = n_tiles when uniform
    unsigned long long off_split;          // byte offset of tile tile_split
    int raw_stride, raw_stages;            // raw ring geometry: stride = the largest qtile of this GEMM (18 432 or 27 648)
};

__device__ __forceinline__ void tma_load_2d_q(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void q_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void q_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void q_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void q_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle, rows of 64 fp16 (8-row atoms of 1024 B): the descriptor prefill_tc5.cu uses
__device__ __forceinline__ uint64_t q_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// 32 lanes x 16 columns of fp32: thread t of the warp receives row (lane base + t), 16 consecutive columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the range of qtile units CTA c owns, and who shares an output tile with whom
__device__ __forceinline__ long long q_range_start(int c, long long U, int G) { return (long long)c * U / G; }
// largest c with range_start(c) <= x
__device__ __forceinline__ int q_owner_of(long long x, long long U, int G) { return (int)(((x + 1) * G + U - 1) / U) - 1; }

// type and byte offset of a tile: arithmetic for the common shapes (no dependent global loads in front of the first TMA)
__device__ __forceinline__ int q_tile_type(const QParams& p, int tile) {
    if (p.tile_type != nullptr) return __ldg(p.tile_type + tile);
    return tile < p.tile_split ? p.type0 : p.type1;
}
__device__ __forceinline__ unsigned long long q_tile_off(const QParams& p, int tile) {
    if (p.tile_off != nullptr) return __ldg(p.tile_off + tile);
    return tile < p.tile_split ? (unsigned long long)tile * p.nkb * (unsigned)qg_qtile_bytes(p.type0)
                               : p.off_split + (unsigned long long)(tile - p.tile_split) * p.nkb * (unsigned)qg_qtile_bytes(p.type1);
}

template <int NB>
__device__ __forceinline__ void q_epilogue(const QParams& p, int n, int lane, const float* v) {
    if (p.epi == GEMM_EPI_SILU) {
        // weight rows are interleaved [8 gate | 8 up] at load: lane l of a 16-lane group holds gate (l < 8) or up (l >= 8) of
        // hidden column (n / 16) * 8 + l % 8
        __half* out = reinterpret_cast<__half*>(p.c);
        const int hcol = (n >> 4) * 8 + (n & 7);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float up = __shfl_xor_sync(0xffffffffu, v[b], 8);
            if ((lane & 8) == 0 && n < p.n) {
                const float g = v[b];
                out[(size_t)b * p.ldc + hcol] = __float2half_rn((g / (1.0f + expf(-g))) * up);
            }
        }
        return;
    }
    if (n >= p.n) return;
    float* out = reinterpret_cast<float*>(p.c) + n;
    if (p.epi == GEMM_EPI_ADD_F32) {
#pragma unroll
        for (int b = 0; b < NB; ++b) out[(size_t)b * p.ldc] += v[b];
    } else {
#pragma unroll
        for (int b = 0; b < NB; ++b) out[(size_t)b * p.ldc] = v[b];
    }
}

template <int NB>
__global__ void __launch_bounds__(QG_THREADS, 1) qgemm_kernel(const __grid_constant__ QParams p) {
    using Cfg = QCfg<NB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;                                                   // [6][128 x 64 fp16]        unpack warps -> tensor core
    uint8_t* act_ring = a_ring + Cfg::A_BYTES;                                // [2][4][NB x 64 fp16]      TMA (L2)     -> tensor core
    uint8_t* raw_ring = act_ring + Cfg::ACT_BYTES;                            // [raw_stages][raw_stride]  TMA (HBM)    -> unpack warps
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (QG_SMEM - 1024 - Cfg::BAR_BYTES));
    uint64_t* raw_full = bars;                               // [6]  producer (tx bytes)  -> unpack warps
    uint64_t* raw_empty = raw_full + QG_MAX_RAW_STAGES;      // [6]  8 unpack warps       -> producer
    uint64_t* act_full = raw_empty + QG_MAX_RAW_STAGES;      // [2]  producer (tx bytes)  -> MMA issuer
    uint64_t* act_empty = act_full + QG_ACT_UNITS;           // [2]  commit               -> activation producer
    uint64_t* a_full = act_empty + QG_ACT_UNITS;             // [6]  4 unpack warps       -> MMA issuer
    uint64_t* a_empty = a_full + QG_A_SLOTS;                 // [6]  commit               -> unpack warps
    uint64_t* acc_full = a_empty + QG_A_SLOTS;               // [2]  commit               -> epilogue
    uint64_t* acc_empty = acc_full + 2;                      // [2]  4 epilogue warps     -> MMA issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int* flag = reinterpret_cast<int*>(tmem_slot + 1);                        // "this CTA finishes the shared tile"

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const long long U = (long long)p.n_tiles * p.nkb;
    const long long u0 = q_range_start(cta, U, G), u1 = q_range_start(cta + 1, U, G);
    const int R = p.raw_stages;

    if (warp == 14 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tb) : "memory");
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < QG_MAX_RAW_STAGES; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 8); }
        for (int i = 0; i < QG_ACT_UNITS; ++i) { mbar_init(&act_full[i], 1); mbar_init(&act_empty[i], 1); }
        for (int i = 0; i < QG_A_SLOTS; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    q_fence_before();
    __syncthreads();
    q_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== weight producer: one 1-D bulk copy per qtile, as far ahead as the raw ring is deep =====
            for (long long u = u0; u < u1; ++u) {
                const int i = (int)(u - u0), s = i % R;
                const uint32_t ph = (uint32_t)(i / R) & 1u;
                const int tile = (int)(u / p.nkb), kb = (int)(u % p.nkb);
                const uint32_t qb = (uint32_t)qg_qtile_bytes(q_tile_type(p, tile));
                mbar_wait(&raw_empty[s], ph ^ 1u);
                mbar_expect_tx(&raw_full[s], qb);
                tma_load_1d(raw_ring + (size_t)s * p.raw_stride, p.w + q_tile_off(p, tile) + (size_t)kb * qb, qb, &raw_full[s]);
            }
        }
    } else if (warp == 14) {
        if (lane == 0) {
            // ===== activation producer: the four [NB x 64] boxes of a qtile's K range (L2-resident), two qtiles deep =====
            for (long long u = u0; u < u1; ++u) {
                const int i = (int)(u - u0), s = i % QG_ACT_UNITS;
                const uint32_t ph = (uint32_t)(i / QG_ACT_UNITS) & 1u;
                const int kb = (int)(u % p.nkb);
                uint8_t* st = act_ring + (size_t)s * 4 * Cfg::B_TILE;
                mbar_wait(&act_empty[s], ph ^ 1u);
                mbar_expect_tx(&act_full[s], 4u * Cfg::B_TILE);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tma_load_2d_q(st + kk * Cfg::B_TILE, &p.tb, kb * QG_COLS + kk * QG_KSTEP, 0, &act_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            // instruction descriptor: D = F32, A / B = F16, both K-major, N = NB, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int seg = 0;
            for (long long u = u0; u < u1; ++seg) {
                const int kb_lo = (int)(u % p.nkb);
                const int n_kb = (int)min((long long)(p.nkb - kb_lo), u1 - u);
                const int buf = seg & 1;
                mbar_wait(&acc_empty[buf], ((uint32_t)(seg >> 1) & 1u) ^ 1u);
                q_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * NB);
                for (int j = 0; j < n_kb; ++j, ++u) {
                    const int i = (int)(u - u0), sa = i % QG_ACT_UNITS;
                    mbar_wait(&act_full[sa], (uint32_t)(i / QG_ACT_UNITS) & 1u);      // the activation boxes of this qtile have landed
                    q_fence_after();
                    const uint32_t sb = smem_u32(act_ring + (size_t)sa * 4 * Cfg::B_TILE);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int g = 4 * i + kk, slot = g % QG_A_SLOTS;
                        mbar_wait(&a_full[slot], (uint32_t)(g / QG_A_SLOTS) & 1u);    // the unpack warps have written this operand tile
                        q_fence_after();
                        const uint64_t adesc = q_desc_sw128(smem_u32(a_ring + (size_t)slot * QG_A_TILE_BYTES));
                        const uint64_t bdesc = q_desc_sw128(sb + (uint32_t)(kk * Cfg::B_TILE));
#pragma unroll
                        for (int k = 0; k < QG_KSTEP / 16; ++k)
                            q_mma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (j | kk | k) ? 1u : 0u);
                        q_commit(&a_empty[slot]);                     // the slot may be overwritten once these MMAs have read it
                    }
                    q_commit(&act_empty[sa]);                         // ... and the qtile's activation boxes
                    if (j == n_kb - 1) q_commit(&acc_full[buf]);
                }
            }
        }
    } else if (warp < 10) {
        // ===== unpack warps: thread (row r, half h) =====
        const int t = (warp - 2) * 32 + lane, r = t & 127, h = t >> 7;
        for (long long u = u0; u < u1; ++u) {
            const int i = (int)(u - u0), s = i % R;
            const uint32_t ph = (uint32_t)(i / R) & 1u;
            const int type = q_tile_type(p, (int)(u / p.nkb));
            const uint8_t* raw = raw_ring + (size_t)s * p.raw_stride;
            mbar_wait(&raw_full[s], ph);
            qg_dequant_thread(
                type, raw, r, h, [&](int kk) { return a_ring + (size_t)((4 * i + kk) % QG_A_SLOTS) * QG_A_TILE_BYTES; },
                [&](int kk) {                                                                  // the MMAs that last read this slot are done
                    const int g = 4 * i + kk;
                    mbar_wait(&a_empty[g % QG_A_SLOTS], ((uint32_t)(g / QG_A_SLOTS) & 1u) ^ 1u);
                },
                [&](int kk) {
                    fence_proxy_async();                                                      // generic-proxy stores -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[(4 * i + kk) % QG_A_SLOTS]);
                });
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[s]);                                        // this warp has read all it needs of the raw bytes
        }
    } else {
        // ===== epilogue warps (10..13) =====
        const int q = warp & 3;                                       // TMEM lane quarter this warp may access
        const int nl = q * 32 + lane;                                 // row inside the tile
        int seg = 0;
        for (long long u = u0; u < u1; ++seg) {
            const int tile = (int)(u / p.nkb), kb_lo = (int)(u % p.nkb);
            const int n_kb = (int)min((long long)(p.nkb - kb_lo), u1 - u);
            u += n_kb;
            const int buf = seg & 1;
            mbar_wait(&acc_full[buf], (uint32_t)(seg >> 1) & 1u);
            q_fence_after();
            float v[NB];
#pragma unroll
            for (int c = 0; c < NB / 16; ++c) tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NB + c * 16), v + c * 16);
            q_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);              // the accumulator half is in registers: the next tile may start
            const int n = tile * QG_ROWS + nl;
            if (n_kb == p.nkb) {                                      // the whole K range of this tile is ours
                q_epilogue<NB>(p, n, lane, v);
                continue;
            }
            // shared tile: partial -> scratch, ticket; the last contributor sums the partials in CTA order
            const long long t0 = (long long)tile * p.nkb;
            const int c_first = q_owner_of(t0, U, G), c_last = q_owner_of(t0 + p.nkb - 1, U, G);
            float* mine = p.partial + ((size_t)cta * 2 + (cta == c_first ? 1 : 0)) * (NB * QG_ROWS);
#pragma unroll
            for (int b = 0; b < NB; ++b) mine[b * QG_ROWS + nl] = v[b];
            __threadfence();
            named_bar_sync(1, 128);
            if (warp == 10 && lane == 0) {
                unsigned ticket;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(p.counters + tile) : "memory");
                const int last = ticket == (unsigned)(c_last - c_first);
                if (last) p.counters[tile] = 0;                       // ready for the next launch
                *flag = last;
            }
            named_bar_sync(1, 128);
            const int finish = *flag;
            named_bar_sync(1, 128);                                   // everyone has read the flag before the next shared tile rewrites it
            if (!finish) continue;
            // sum in CTA order (deterministic); two contributors per round trip: their loads are independent
#pragma unroll
            for (int b = 0; b < NB; ++b) v[b] = 0.f;
            int c = c_first;
            for (; c + 1 <= c_last; c += 2) {
                const float* pa = p.partial + ((size_t)c * 2 + (c == c_first ? 1 : 0)) * (NB * QG_ROWS) + nl;
                const float* pb = p.partial + ((size_t)(c + 1) * 2) * (NB * QG_ROWS) + nl;
                float ta[NB], tb2[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b) { ta[b] = __ldcg(pa + b * QG_ROWS); tb2[b] = __ldcg(pb + b * QG_ROWS); }
#pragma unroll
                for (int b = 0; b < NB; ++b) v[b] = (v[b] + ta[b]) + tb2[b];
            }
            if (c <= c_last) {
                const float* pa = p.partial + ((size_t)c * 2 + (c == c_first ? 1 : 0)) * (NB * QG_ROWS) + nl;
#pragma unroll
                for (int b = 0; b < NB; ++b) v[b] += __ldcg(pa + b * QG_ROWS);
            }
            q_epilogue<NB>(p, n, lane, v);
        }
    }
    q_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
}

// ---- load time: native GGUF rows -> QG qtile stream -----------------------------------------------------------------------
struct PackParams {
    QGemmSource src[3];
    int nsrc, mode, nkb, n_rows;
    const uint64_t* tile_off;
    uint8_t* dst;
};
__global__ void __launch_bounds__(128) qgemm_pack_kernel(const PackParams p) {
    const int row = blockIdx.x * 128 + threadIdx.x;       // output row of the GEMM
    const int kb = blockIdx.y;
    if (row >= p.n_rows) return;
    int si = 0, srow = row;
    if (p.mode == 1) {                                     // [8 gate | 8 up] groups
        si = (row & 15) >> 3;
        srow = (row >> 4) * 8 + (row & 7);
    } else {
        while (si < p.nsrc - 1 && srow >= p.src[si].rows) { srow -= p.src[si].rows; ++si; }
    }
    const int type = p.src[si].type;
    const int bb = type == 12 ? 144 : 210;
    const uint8_t* blk = p.src[si].w + ((size_t)srow * p.nkb + kb) * bb;
    const int tile = row >> 7;
    qg_pack_block(type, blk, p.dst + p.tile_off[tile] + (size_t)kb * qg_qtile_bytes(type), row & 127);
}

typedef CUresult (*EncodeTiledFnQ)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnQ encode_fn_q() {
    static EncodeTiledFnQ fn = []() -> EncodeTiledFnQ {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFnQ>(f);
    }();
    return fn;
}

template <int NB>
cudaError_t launch_nb(QParams& qp, int grid, cudaStream_t s) {
    qp.raw_stages = std::min(QG_MAX_RAW_STAGES, QCfg<NB>::RAW_BUDGET / qp.raw_stride);
    if (qp.raw_stages < 2) return cudaErrorInvalidValue;
    qgemm_kernel<NB><<<grid, QG_THREADS, QG_SMEM, s>>>(qp);
    return cudaGetLastError();
}

}  // namespace

size_t qgemm_partial_floats(int nb) { return (size_t)QGEMM_MAX_GRID * 2 * nb * QG_ROWS; }
bool qgemm_batch_ok(int nb) { return nb == 16 || nb == 32 || nb == 64; }

cudaError_t qgemm_configure() {
    cudaError_t e = cudaFuncSetAttribute(qgemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QG_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qgemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QG_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qgemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QG_SMEM);
    return e;
}

cudaError_t qgemm_pack_launch(const QGemmSource* src, int nsrc, int mode, int k, uint8_t* dst, uint64_t* tile_off_host, uint8_t* tile_type_host,
                              cudaStream_t s) {
    if (nsrc < 1 || nsrc > 3 || (mode == 1 && nsrc != 2) || k % QG_COLS) return cudaErrorInvalidValue;
    int n_rows = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (!qg_type_ok(src[i].type) || src[i].rows % (mode == 1 ? 64 : QG_ROWS)) return cudaErrorInvalidValue;
        n_rows += src[i].rows;
    }
    if (mode == 1 && (src[0].rows != src[1].rows || src[0].type != src[1].type)) return cudaErrorInvalidValue;
    const int nkb = k / QG_COLS, n_tiles = n_rows / QG_ROWS;
    uint64_t off = 0;
    for (int t = 0; t < n_tiles; ++t) {
        int type;
        if (mode == 1) type = src[0].type;
        else {
            int row = t * QG_ROWS, si = 0;
            while (si < nsrc - 1 && row >= src[si].rows) { row -= src[si].rows; ++si; }
            type = src[si].type;
        }
        tile_off_host[t] = off;
        tile_type_host[t] = (uint8_t)type;
        off += (uint64_t)nkb * qg_qtile_bytes(type);
    }
    uint64_t* d_off = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_off, (size_t)n_tiles * 8);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(d_off, tile_off_host, (size_t)n_tiles * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        PackParams pp{};
        for (int i = 0; i < nsrc; ++i) pp.src[i] = src[i];
        pp.nsrc = nsrc; pp.mode = mode; pp.nkb = nkb; pp.n_rows = n_rows; pp.tile_off = d_off; pp.dst = dst;
        qgemm_pack_kernel<<<dim3((unsigned)n_tiles, (unsigned)nkb), 128, 0, s>>>(pp);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_off);
    return e;
}

cudaError_t qgemm_launch(const QGemmWeights& wt, const __half* act, int act_rows_alloc, int nb, void* c, int ldc, int epi, float* partial,
                         int n_sm, cudaStream_t s) {
    if (!qgemm_batch_ok(nb) || act_rows_alloc < nb || !wt.w || (wt.k % QG_COLS) || ((uintptr_t)act % 16)) return cudaErrorInvalidValue;
    if (epi != GEMM_EPI_F32 && epi != GEMM_EPI_ADD_F32 && epi != GEMM_EPI_SILU) return cudaErrorInvalidValue;
    EncodeTiledFnQ fn = encode_fn_q();
    if (!fn) return cudaErrorInvalidValue;
    QParams qp{};
    const cuuint64_t dims[2] = {(cuuint64_t)wt.k, (cuuint64_t)act_rows_alloc};
    const cuuint64_t strides[1] = {(cuuint64_t)wt.k * 2};
    const cuuint32_t box[2] = {(cuuint32_t)QG_KSTEP, (cuuint32_t)nb};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&qp.tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(act), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    qp.w = wt.w; qp.counters = wt.counters; qp.partial = partial;
    // tile addressing: arithmetic when the tiles are "type0, then type1" (every Llama GEMM: uniform, or Q | K | V with another V
    // type); the per-tile tables only for anything else
    qp.type0 = wt.type0; qp.type1 = wt.type1; qp.tile_split = wt.tile_split; qp.off_split = wt.off_split;
    qp.tile_off = wt.two_segment ? nullptr : wt.tile_off;
    qp.tile_type = wt.two_segment ? nullptr : wt.tile_type;
    qp.raw_stride = wt.has_q6k ? QG_Q6K_BYTES + 768 : QG_Q4K_BYTES;      // 27 648 / 18 432: multiples of 1024
    qp.c = c; qp.ldc = ldc; qp.epi = epi; qp.n = wt.n; qp.n_tiles = wt.n_tiles; qp.nkb = wt.nkb;
    const long long U = (long long)wt.n_tiles * wt.nkb;
    const int grid = (int)std::min<long long>(std::min(n_sm, QGEMM_MAX_GRID), U);
    if (nb == 16) return launch_nb<16>(qp, grid, s);
    if (nb == 32) return launch_nb<32>(qp, grid, s);
    return launch_nb<64>(qp, grid, s);
}

}  // namespace gl
