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

constexpr int QG_THREADS = 14 * 32;
constexpr int QG_RAW_STRIDE = 27648;                 // >= 26 880, multiple of 1024
constexpr int QG_A_SLOTS = 4;                        // operand-tile ring = the four K-steps of one qtile

template <int NB> struct QCfg {
    static constexpr int STAGES = NB <= 32 ? 3 : 2;
    static constexpr int B_TILE = NB * 128;          // activations [NB rows x 64] fp16, 128-byte swizzle
    static constexpr int STAGE_BYTES = QG_RAW_STRIDE + 4 * B_TILE;
    static constexpr int TMEM_COLS = 2 * NB < 32 ? 32 : 2 * NB;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + (size_t)QG_A_SLOTS * QG_A_TILE_BYTES + 1024 /* alignment */ + 512 /* barriers */;
};

struct QParams {
    CUtensorMap tb;              // activations: dims {K, rows_alloc}, box {64, NB}, 128-byte swizzle
    const uint8_t* w;
    const uint64_t* tile_off;
    const uint8_t* tile_type;
    unsigned* counters;
    float* partial;              // [grid][2][NB * 128]
    void* c;
    int ldc, epi, n, n_tiles, nkb;
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
    uint8_t* stages = smem;                                                   // [STAGES][raw | 4 activation tiles]
    uint8_t* a_ring = smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES;          // [4][128 x 64 fp16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + (size_t)QG_A_SLOTS * QG_A_TILE_BYTES);
    uint64_t* stage_full = bars;                       // [STAGES]  producer (tx bytes)     -> unpack warps, MMA issuer
    uint64_t* stage_empty = stage_full + Cfg::STAGES;  // [STAGES]  8 unpack warps + 1 commit -> producer
    uint64_t* a_full = stage_empty + Cfg::STAGES;      // [4]       4 unpack warps            -> MMA issuer
    uint64_t* a_empty = a_full + QG_A_SLOTS;           // [4]       commit                    -> unpack warps
    uint64_t* acc_full = a_empty + QG_A_SLOTS;         // [2]       commit                    -> epilogue
    uint64_t* acc_empty = acc_full + 2;                // [2]       4 epilogue warps          -> MMA issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int* flag = reinterpret_cast<int*>(tmem_slot + 1);                        // "this CTA finishes the shared tile"

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const long long U = (long long)p.n_tiles * p.nkb;
    const long long u0 = q_range_start(cta, U, G), u1 = q_range_start(cta + 1, U, G);

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tb) : "memory");
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&stage_full[i], 1); mbar_init(&stage_empty[i], 9); }
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
            // ===== producer =====
            for (long long u = u0; u < u1; ++u) {
                const int i = (int)(u - u0), s = i % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(i / Cfg::STAGES) & 1u;
                const int tile = (int)(u / p.nkb), kb = (int)(u % p.nkb);
                const uint32_t qb = (uint32_t)qg_qtile_bytes(p.tile_type[tile]);
                uint8_t* st = stages + (size_t)s * Cfg::STAGE_BYTES;
                mbar_wait(&stage_empty[s], ph ^ 1u);
                mbar_expect_tx(&stage_full[s], qb + 4u * Cfg::B_TILE);
                tma_load_1d(st, p.w + p.tile_off[tile] + (size_t)kb * qb, qb, &stage_full[s]);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tma_load_2d_q(st + QG_RAW_STRIDE + kk * Cfg::B_TILE, &p.tb, kb * QG_COLS + kk * QG_KSTEP, 0, &stage_full[s]);
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
                    const int i = (int)(u - u0), s = i % Cfg::STAGES;
                    const uint32_t ph = (uint32_t)(i / Cfg::STAGES) & 1u;
                    mbar_wait(&stage_full[s], ph);                   // the activation tiles of this stage have landed
                    q_fence_after();
                    const uint32_t sb = smem_u32(stages + (size_t)s * Cfg::STAGE_BYTES + QG_RAW_STRIDE);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        mbar_wait(&a_full[kk], (uint32_t)i & 1u);    // the unpack warps have written operand tile kk of this qtile
                        q_fence_after();
                        const uint64_t adesc = q_desc_sw128(smem_u32(a_ring + (size_t)kk * QG_A_TILE_BYTES));
                        const uint64_t bdesc = q_desc_sw128(sb + (uint32_t)(kk * Cfg::B_TILE));
#pragma unroll
                        for (int k = 0; k < QG_KSTEP / 16; ++k)
                            q_mma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (j | kk | k) ? 1u : 0u);
                        q_commit(&a_empty[kk]);                       // operand tile kk may be overwritten once these MMAs have read it
                    }
                    q_commit(&stage_empty[s]);                        // ... and the stage's activation tiles
                    if (j == n_kb - 1) q_commit(&acc_full[buf]);
                }
            }
        }
    } else if (warp < 10) {
        // ===== unpack warps: thread (row r, half h) =====
        const int t = (warp - 2) * 32 + lane, r = t & 127, h = t >> 7;
        for (long long u = u0; u < u1; ++u) {
            const int i = (int)(u - u0), s = i % Cfg::STAGES;
            const uint32_t ph = (uint32_t)(i / Cfg::STAGES) & 1u;
            const int type = p.tile_type[(int)(u / p.nkb)];
            const uint8_t* raw = stages + (size_t)s * Cfg::STAGE_BYTES;
            mbar_wait(&stage_full[s], ph);
            qg_dequant_thread(
                type, raw, r, h, [&](int kk) { return a_ring + (size_t)kk * QG_A_TILE_BYTES; },
                [&](int kk) { mbar_wait(&a_empty[kk], ((uint32_t)i & 1u) ^ 1u); },          // the MMAs of the previous qtile have read slot kk
                [&](int kk) {
                    fence_proxy_async();                                                      // generic-proxy stores -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[kk]);
                });
            __syncwarp();
            if (lane == 0) mbar_arrive(&stage_empty[s]);                                      // this warp has read all it needs of the raw bytes
        }
    } else {
        // ===== epilogue warps =====
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
#pragma unroll
            for (int b = 0; b < NB; ++b) v[b] = 0.f;
            for (int c = c_first; c <= c_last; ++c) {
                const float* part = p.partial + ((size_t)c * 2 + (c == c_first ? 1 : 0)) * (NB * QG_ROWS);
#pragma unroll
                for (int b = 0; b < NB; ++b) v[b] += __ldcg(part + b * QG_ROWS + nl);
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
cudaError_t launch_nb(const QParams& qp, int grid, cudaStream_t s) {
    qgemm_kernel<NB><<<grid, QG_THREADS, QCfg<NB>::SMEM, s>>>(qp);
    return cudaGetLastError();
}

}  // namespace

size_t qgemm_partial_floats(int nb) { return (size_t)QGEMM_MAX_GRID * 2 * nb * QG_ROWS; }
bool qgemm_batch_ok(int nb) { return nb == 16 || nb == 32 || nb == 64; }

cudaError_t qgemm_configure() {
    cudaError_t e = cudaFuncSetAttribute(qgemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QCfg<16>::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qgemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QCfg<32>::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qgemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QCfg<64>::SMEM);
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
    qp.w = wt.w; qp.tile_off = wt.tile_off; qp.tile_type = wt.tile_type; qp.counters = wt.counters; qp.partial = partial;
    qp.c = c; qp.ldc = ldc; qp.epi = epi; qp.n = wt.n; qp.n_tiles = wt.n_tiles; qp.nkb = wt.nkb;
    const long long U = (long long)wt.n_tiles * wt.nkb;
    const int grid = (int)std::min<long long>(std::min(n_sm, QGEMM_MAX_GRID), U);
    if (nb == 16) return launch_nb<16>(qp, grid, s);
    if (nb == 32) return launch_nb<32>(qp, grid, s);
    return launch_nb<64>(qp, grid, s);
}

}  // namespace gl
