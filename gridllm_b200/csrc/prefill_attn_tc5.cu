// Prompt attention on the 5th-generation tensor cores (head dim 128): S = Q K^T and O += P V as tcgen05.mma with the
// scores, the probabilities and the output accumulator in TENSOR MEMORY -- the tcgen05 counterpart of prefill_attn.cu's
// mma.sync kernel, which ncu shows bound by the legacy tensor path (61 % of a pipe a quarter as wide).  Same place on the
// path: the prompt-evaluation phase behind OllamaService.generate*Response / generateEmbedding
// (/root/reference/client/src/services/OllamaService.ts:142-145, 235-237, 633-636).
//
// One CTA = 128 query rows of one head of one sequence, six warps, everything asynchronous:
//   warp 0 / one lane : TMA producer -- the Q tile once, then per KV tile of 128 rows the K tile [kv][hd] and the V^T tile
//                       [hd][kv] (both K-major B operands: no transposes), two 64 KB stages;
//   warp 1 / one lane : MMA issuer   -- S[j & 1] = Q K_j^T (8 x M128 N128 K16, A / B from shared memory), then
//                       O += P_(j-1) V_(j-1) (8 x M128 N128 K16 with A = P FROM TENSOR MEMORY, the ".ts" form): the products of
//                       tile j + 1 and j - 1 run while the softmax warps work on tile j;
//   warps 2..5        : softmax      -- thread = query row (no shuffles): tcgen05.ld of the row's scores, causal mask on the
//                       diagonal tile, running maximum in the exp2 domain, P = exp2(s - m) rounded to fp16 and written over the
//                       first half of the same S buffer (tcgen05.st), row sum over the rounded values; the output accumulator is
//                       rescaled in tensor memory only when a row's maximum moved (warp-uniform test), after the previous P V
//                       product has completed (o_ready);  at the end O / sum -> fp16 rows.
// Tensor memory: S0 / P0 columns 0..127, S1 / P1 128..255, O 256..383 (512 allocated).  tcgen05.mma instructions of one thread
// execute in issue order, which is what makes re-using an S buffer two tiles later safe without another barrier: the product
// that read P_(j-2) out of it was issued before the product that overwrites it.
// A pack of sequences is one launch (blockIdx.z = sequence, tiles relative to the sequence's first row), as in prefill_attn.cu.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "prefill.h"

namespace gl {

namespace {

constexpr int TA_BM = 128, TA_BN = 128, TA_HD = 128, TA_THREADS = 192;
constexpr int TA_TILE_BYTES = 128 * 64 * 2;              // one 64-column half of a 128-row operand tile
constexpr int TA_Q_BYTES = 2 * TA_TILE_BYTES;            // 32 KB
constexpr int TA_STAGE_BYTES = 4 * TA_TILE_BYTES;        // K (2 halves) + V^T (2 halves) = 64 KB
constexpr int TA_STAGES = 2;
constexpr size_t TA_SMEM = TA_Q_BYTES + (size_t)TA_STAGES * TA_STAGE_BYTES + 1024 /* alignment */ + 256 /* barriers */;
constexpr int TA_TMEM_COLS = 512;
constexpr uint32_t TA_COL_S = 0, TA_COL_O = 256;

struct TaParams {
    CUtensorMap tq;      // q  [rows][qd]:     dims {qd, rows},    box {64, 128}
    CUtensorMap tk;      // k  [rows][kvd]:    dims {kvd, rows},   box {64, 128}
    CUtensorMap tv;      // vt [kvd][vt_ld]:   dims {vt_ld, kvd},  box {64, 128}
    __half* out;         // [rows][qd]
    int qd, grp;
    float scale_log2;
    PrefillSegs segs;
};

__device__ __forceinline__ void ta_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void ta_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ta_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool ta_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void ta_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ta_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A from tensor memory: 128 lanes x 8 columns of two fp16 per K = 16
__device__ __forceinline__ void ta_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle, rows of 64 fp16 (8-row atoms of 1024 B)
__device__ __forceinline__ uint64_t ta_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// (no wait: several loads are issued back to back, then ta_wait_ld())
__device__ __forceinline__ void ta_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ta_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ta_st32(uint32_t taddr, const uint32_t* w) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]),
          "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]),
          "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31])
        : "memory");
}
__device__ __forceinline__ void ta_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ta_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(TA_THREADS, 1) flash_tc5_kernel(const __grid_constant__ TaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sq = smem;
    uint8_t* skv = smem + TA_Q_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(skv + (size_t)TA_STAGES * TA_STAGE_BYTES);
    uint64_t* q_full = bars;                 // [1]
    uint64_t* kv_full = bars + 1;            // [2]
    uint64_t* kv_empty = bars + 3;           // [2]
    uint64_t* s_full = bars + 5;             // [2] scores of tile j are in S[j & 1]
    uint64_t* p_full = bars + 7;             // [2] probabilities of tile j are in S[j & 1] (and O is rescaled): 4 warps arrive
    uint64_t* o_ready = bars + 9;            // [1] P V of tile j has completed (phase j)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seq = blockIdx.z, h = blockIdx.x;
    const int len = p.segs.len[seq], r0 = p.segs.start[seq];
    const int n_qt = (len + TA_BM - 1) / TA_BM;
    if ((int)blockIdx.y >= n_qt) return;
    const int qt = n_qt - 1 - (int)blockIdx.y;        // the longest tiles of a sequence first
    const int m0 = qt * TA_BM;
    const int kvh = h / p.grp;
    const int n_kt = qt + 1;                          // KV tiles 0 .. qt (BN = BM: the last one is the diagonal tile)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tq) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tk) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tv) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
        }
        mbar_init(o_ready, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TA_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ta_fence_before();
    __syncthreads();
    ta_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            mbar_expect_tx(q_full, TA_Q_BYTES);
            ta_tma_2d(sq, &p.tq, h * TA_HD, r0 + m0, q_full);
            ta_tma_2d(sq + TA_TILE_BYTES, &p.tq, h * TA_HD + 64, r0 + m0, q_full);
            for (int j = 0; j < n_kt; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_empty[st], (((uint32_t)(j >> 1)) & 1u) ^ 1u);
                uint8_t* sk = skv + (size_t)st * TA_STAGE_BYTES;
                mbar_expect_tx(&kv_full[st], TA_STAGE_BYTES);
                const int kv0 = r0 + j * TA_BN;
                ta_tma_2d(sk, &p.tk, kvh * TA_HD, kv0, &kv_full[st]);
                ta_tma_2d(sk + TA_TILE_BYTES, &p.tk, kvh * TA_HD + 64, kv0, &kv_full[st]);
                ta_tma_2d(sk + 2 * TA_TILE_BYTES, &p.tv, kv0, kvh * TA_HD, &kv_full[st]);
                ta_tma_2d(sk + 3 * TA_TILE_BYTES, &p.tv, kv0 + 64, kvh * TA_HD, &kv_full[st]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop, one elected lane issues (uniform operands: see prefill_tc5.cu) =====
        // instruction descriptor: D = F32, A / B = F16, both K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t sq_u = __shfl_sync(0xffffffffu, smem_u32(sq), 0);
        const uint32_t skv_u = __shfl_sync(0xffffffffu, smem_u32(skv), 0);
        mbar_wait(q_full, 0);
        for (int j = 0; j <= n_kt; ++j) {
            if (j < n_kt) {
                // S[j & 1] = Q K_j^T
                const int st = j & 1;
                mbar_wait(&kv_full[st], ((uint32_t)(j >> 1)) & 1u);
                ta_fence_after();
                const uint32_t sk = skv_u + (uint32_t)(st * TA_STAGE_BYTES);
                const uint32_t tmem_s = tb + TA_COL_S + (uint32_t)(st * 128);
                if (ta_elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < TA_HD / 16; ++kk) {
                        const uint64_t ad = ta_desc_sw128(sq_u + (uint32_t)((kk >> 2) * TA_TILE_BYTES)) + (uint64_t)(2 * (kk & 3));
                        const uint64_t bd = ta_desc_sw128(sk + (uint32_t)((kk >> 2) * TA_TILE_BYTES)) + (uint64_t)(2 * (kk & 3));
                        ta_mma_ss(tmem_s, ad, bd, idesc, kk ? 1u : 0u);
                    }
                    ta_commit(&s_full[st]);
                }
                __syncwarp();
            }
            if (j >= 1) {
                // O += P_(j-1) V_(j-1): A = the probabilities the softmax warps wrote over S[(j-1) & 1]
                const int i = j - 1, st = i & 1;
                mbar_wait(&p_full[st], ((uint32_t)(i >> 1)) & 1u);
                ta_fence_after();
                const uint32_t sv = skv_u + (uint32_t)(st * TA_STAGE_BYTES + 2 * TA_TILE_BYTES);
                const uint32_t tmem_p = tb + TA_COL_S + (uint32_t)(st * 128);
                if (ta_elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < TA_BN / 16; ++kk) {
                        const uint64_t bd = ta_desc_sw128(sv + (uint32_t)((kk >> 2) * TA_TILE_BYTES)) + (uint64_t)(2 * (kk & 3));
                        ta_mma_ts(tb + TA_COL_O, tmem_p + (uint32_t)(8 * kk), bd, idesc, (i | kk) ? 1u : 0u);
                    }
                    ta_commit(&kv_empty[st]);             // the stage's K and V^T tiles have been read
                    ta_commit(o_ready);                   // ... and O holds tiles 0 .. i
                }
                __syncwarp();
            }
        }
    } else {
        // ===== softmax warps: thread = query row =====
        const int qtr = warp & 3;                              // the TMEM lane quarter this warp may access
        const int row = m0 + qtr * 32 + lane;                  // relative to the sequence
        const uint32_t lane_base = tmem_base + ((uint32_t)(qtr * 32) << 16);
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kt; ++j) {
            const int st = j & 1;
            const uint32_t ts = lane_base + TA_COL_S + (uint32_t)(st * 128);
            mbar_wait(&s_full[st], ((uint32_t)(j >> 1)) & 1u);
            ta_fence_after();
            const bool diag = (j == n_kt - 1);
            const int col0 = j * TA_BN;
            // the row's 128 scores are read from tensor memory ONCE (its read port moves 64 B per cycle: two passes over the tile
            // cost as many cycles as the tile's 16 MMAs) and stay in registers for the maximum and the exponentials
            uint32_t v[128];
            ta_ld32(ts, v);
            ta_ld32(ts + 32u, v + 32);
            ta_ld32(ts + 64u, v + 64);
            ta_ld32(ts + 96u, v + 96);
            ta_wait_ld();
            float mx = -INFINITY;
#pragma unroll
            for (int e = 0; e < 128; ++e) {
                const float s = __uint_as_float(v[e]);
                if (!diag || col0 + e <= row) mx = fmaxf(mx, s);
            }
            // column 0 of the first tile is never masked: the maximum is finite from the first tile on
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float corr = ta_exp2(m_run - m_new);       // first tile: exp2(-inf) = 0
            // the accumulator holds tiles 0 .. j-1 once the previous P V product has completed; rescale it where a maximum moved
            if (j > 0) {
                mbar_wait(o_ready, ((uint32_t)(j - 1)) & 1u);
                ta_fence_after();
                if (__any_sync(0xffffffffu, m_new > m_run)) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t ov[32];
                        ta_ld32(lane_base + TA_COL_O + (uint32_t)(c * 32), ov);
                        ta_wait_ld();
#pragma unroll
                        for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * corr);
                        ta_st32(lane_base + TA_COL_O + (uint32_t)(c * 32), ov);
                    }
                }
            }
            m_run = m_new;
            // P = exp2(s * scale - m) -> fp16 pairs, written over the first 64 columns of the same buffer (every score is in
            // registers by now)
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t w[32];
#pragma unroll
                for (int e = 0; e < 64; e += 2) {
                    const int cc = col0 + c * 64 + e;
                    float p0 = ta_exp2(fmaf(__uint_as_float(v[c * 64 + e]), p.scale_log2, -m_new));
                    float p1 = ta_exp2(fmaf(__uint_as_float(v[c * 64 + e + 1]), p.scale_log2, -m_new));
                    if (diag) {
                        if (cc > row) p0 = 0.f;
                        if (cc + 1 > row) p1 = 0.f;
                    }
                    const __half2 ph = __floats2half2_rn(p0, p1);
                    const float2 pf = __half22float2(ph);
                    sum += pf.x + pf.y;                       // the row sum is taken over the ROUNDED probabilities
                    w[e / 2] = *reinterpret_cast<const uint32_t*>(&ph);
                }
                ta_st32(ts + (uint32_t)(c * 32), w);
            }
            l_run = l_run * corr + sum;
            ta_wait_st();
            ta_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[st]);
        }
        // ---- O / sum -> fp16 row (rows beyond the sequence inside its last tile: zeros) ----
        mbar_wait(o_ready, ((uint32_t)(n_kt - 1)) & 1u);
        ta_fence_after();
        const float inv = (row < len && l_run > 0.f) ? 1.0f / l_run : 0.f;
        __half* orow = p.out + (size_t)(r0 + row) * p.qd + (size_t)h * TA_HD;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            ta_ld32(lane_base + TA_COL_O + (uint32_t)(c * 32), v);
            ta_wait_ld();
            __align__(16) __half2 o[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __floats2half2_rn(__uint_as_float(v[2 * e]) * inv, __uint_as_float(v[2 * e + 1]) * inv);
#pragma unroll
            for (int e = 0; e < 4; ++e) reinterpret_cast<uint4*>(orow + c * 32)[e] = reinterpret_cast<const uint4*>(o)[e];
        }
    }
    ta_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TA_TMEM_COLS) : "memory");
}

typedef CUresult (*TaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TaEncodeFn ta_encode_fn() {
    static TaEncodeFn fn = []() -> TaEncodeFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<TaEncodeFn>(f);
    }();
    return fn;
}
// rows x cols fp16, row stride ld elements; box = 128 rows x 64 elements, 128-byte swizzle, rows / columns out of range read as zeros
bool ta_make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld) {
    TaEncodeFn fn = ta_encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

cudaError_t flash_tc5_configure() {
    return cudaFuncSetAttribute(flash_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TA_SMEM);
}

bool flash_tc5_supported(int hd) { return hd == TA_HD; }

// rows_alloc = rows of q / k (and columns of vt) that exist in memory
cudaError_t flash_tc5_launch(const __half* q, const __half* k, const __half* vt, __half* out, const PrefillSegs& segs, int n_head, int n_kv, int hd,
                             int vt_ld, int rows_alloc, float scale, cudaStream_t s) {
    if (segs.n < 1 || segs.n > PF_MAX_SEGS || hd != TA_HD || n_kv < 1 || n_head % n_kv || (vt_ld & 7) || rows_alloc < 128 || vt_ld < rows_alloc) return cudaErrorInvalidValue;
    int max_len = 0;
    for (int i = 0; i < segs.n; ++i) {
        if (segs.len[i] < 1 || (segs.start[i] & 127)) return cudaErrorInvalidValue;
        max_len = segs.len[i] > max_len ? segs.len[i] : max_len;
    }
    TaParams p{};
    const int qd = n_head * hd, kvd = n_kv * hd;
    if (!ta_make_map(&p.tq, q, rows_alloc, qd, qd) || !ta_make_map(&p.tk, k, rows_alloc, kvd, kvd) || !ta_make_map(&p.tv, vt, kvd, rows_alloc, vt_ld))
        return cudaErrorInvalidValue;
    p.out = out;
    p.qd = qd;
    p.grp = n_head / n_kv;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.segs = segs;
    const dim3 grid((unsigned)n_head, (unsigned)((max_len + TA_BM - 1) / TA_BM), (unsigned)segs.n);
    flash_tc5_kernel<<<grid, TA_THREADS, TA_SMEM, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gl
