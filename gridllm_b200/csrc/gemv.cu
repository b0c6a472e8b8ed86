// Decode GEMV: y = dequant(W)[rows x K] * x[K], the kernel that bounds Llama decode on B200.
//
// Replaces the quantised mat-vec that, in the reference, runs inside the Ollama daemon behind
// OllamaService.generateResponse / generateStreamResponse
// (/root/reference/client/src/services/OllamaService.ts:142-145, 235-237).
//
// Design.  HBM-bound: every weight byte is used once per token, so the
// kernel is a byte-streaming pipeline, not a GEMM.
//   * persistent grid, one CTA per SM; each CTA owns a contiguous slice of every weight segment's
//     rows (balanced to within 2 rows), so there is no tail wave and no atomics;
//   * one producer warp streams whole rows with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP)
//     into an mbarrier-guarded ring of shared-memory stages (weights are static, so the ring is
//     filled BEFORE griddepcontrol.wait: under programmatic dependent launch the next kernel's
//     weights are already in flight while the previous kernel drains);
//   * eight consumer warps decode blocks straight out of shared memory with dp4a against the
//     activation vector held in registers as two int8 planes (15-bit fixed point per 32-column
//     block; rowdot.h), fp32 accumulate, warp-shuffle reduction;
//   * fused prologue: RMSNorm + activation snap; fused epilogues: residual add, RoPE + KV-page
//     append, SiLU*mul.
#include "common.cuh"
#include "kernels.h"
#include "rowdot.h"
#include "gguf_file.h"

namespace gl {

namespace {

constexpr int NCW = GEMV_CONSUMER_WARPS;
constexpr int NCT = NCW * 32;          // consumer threads
constexpr int MAXP = GEMV_MAX_PASSES;
constexpr int SM_BARS = 0;             // full[8], empty[8]
constexpr int SM_RED = 128;            // 32 floats
constexpr int SM_RES = 256;            // res[2][256] floats
constexpr int SM_X = 256 + 2 * 256 * 4;   // 2304

__host__ __device__ inline int warps_per_row(int cols) {
    int nu = cols / UNIT_COLS;
    int w = (nu + 31) / 32;
    int p = 1;
    while (p < w) p <<= 1;
    return p;       // 1,2,4,8
}

struct WorkRange { int a, b; };
__device__ __forceinline__ WorkRange cta_range(int rows, int gran) {
    const int units = rows / gran;
    const int G = gridDim.x, c = blockIdx.x;
    WorkRange r;
    r.a = (int)(((long long)c * units) / G) * gran;
    r.b = (int)(((long long)(c + 1) * units) / G) * gran;
    return r;
}

template <int ABITS>
__global__ void __launch_bounds__(GEMV_THREADS, 1) gemv_kernel(const __grid_constant__ GemvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM_BARS);
    uint64_t* empty = full + GEMV_MAX_STAGES;
    float* red = reinterpret_cast<float*>(smem + SM_RED);
    float* res = reinterpret_cast<float*>(smem + SM_RES);
    const int K = p.cols;
    uint8_t* xhi = smem + SM_X;
    uint8_t* xlo = xhi + K;
    float* sx_arr = reinterpret_cast<float*>(xlo + K);
    float* sm_arr = sx_arr + K / 32;
    int* s16_arr = reinterpret_cast<int*>(sm_arr + K / 32);
    uint8_t* stages = smem + ((SM_X + 2 * K + K / 2 + 127) & ~127);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.n_stages;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();

    const int gran = (p.epi == EPI_QKV) ? 2 : 1;
    const int nwork = p.pair ? 1 : p.nseg;

    if (warp == NCW) {
        // ===================== producer: stream weight rows, no dependency on upstream kernels ===
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int s = 0; s < nwork; ++s) {
                const GemvSeg sg = p.seg[s];
                const WorkRange wr = cta_range(sg.rows, gran);
                for (int r0 = wr.a; r0 < wr.b; r0 += sg.rows_per_stage) {
                    const int n = min(sg.rows_per_stage, wr.b - r0);
                    const uint32_t bytes = (uint32_t)n * (uint32_t)sg.row_stride;
                    mbar_wait(&empty[st], ph ^ 1);
                    uint8_t* dst = stages + (size_t)st * p.stage_bytes;
                    if (p.pair) {
                        mbar_expect_tx(&full[st], 2 * bytes);
                        tma_load_1d(dst, sg.w + (size_t)r0 * sg.row_stride, bytes, &full[st]);
                        tma_load_1d(dst + bytes, p.seg[1].w + (size_t)r0 * sg.row_stride, bytes, &full[st]);
                    } else {
                        mbar_expect_tx(&full[st], bytes);
                        tma_load_1d(dst, sg.w + (size_t)r0 * sg.row_stride, bytes, &full[st]);
                    }
                    if (++st == S) { st = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ========================= consumers ==========================================================
    pdl_wait();   // x (and the residual / KV pages we write) belong to the upstream kernel
    // NOTE: every read of data produced by an upstream kernel uses ld.global.cg (__ldcg): under PDL
    // this CTA may have been resident (and its SM's L1 populated) before the producer finished.

    // ---- fused prologue: (RMSNorm) + snap x to the int8 planes --------------------------------
    float rstd = 1.f;
    if (p.norm_w != nullptr) {
        float ss = 0.f;
        for (int i = tid * 4; i < K; i += NCT * 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(p.x + i));
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        named_bar_sync(1, NCT);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < NCW; ++w) tot += red[w];
        rstd = 1.0f / sqrtf(tot / (float)K + p.eps);
    }
    {
        const int half = tid & 1;
        const int nblk = K / 32;
        for (int blk = tid >> 1; blk < ((nblk + NCT / 2 - 1) / (NCT / 2)) * (NCT / 2); blk += NCT / 2) {
            const bool live = blk < nblk;
            float v[16];
            float amax = 0.f;
            if (live) {
                const float* xb = p.x + blk * 32 + half * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = __ldcg(reinterpret_cast<const float4*>(xb + 4 * q));
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                if (p.norm_w != nullptr) {
                    const float* wb = p.norm_w + blk * 32 + half * 16;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 wv = *reinterpret_cast<const float4*>(wb + 4 * q);
                        v[4 * q] = (v[4 * q] * rstd) * wv.x;
                        v[4 * q + 1] = (v[4 * q + 1] * rstd) * wv.y;
                        v[4 * q + 2] = (v[4 * q + 2] * rstd) * wv.z;
                        v[4 * q + 3] = (v[4 * q + 3] * rstd) * wv.w;
                    }
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) amax = fmaxf(amax, fabsf(v[q]));
            }
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            uint32_t h4[4], l4[4];
            int vs = 0;
            if (live) snap16<ABITS>(v, amax, h4, l4, &vs);
            const int vs_other = __shfl_xor_sync(0xffffffffu, vs, 1);
            if (live) {
                const int u = blk >> 2;
                const int j = 2 * (blk & 3) + half;
                const int phys = j ^ (u & 7);
                *reinterpret_cast<uint4*>(xhi + u * 128 + phys * 16) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
                if (ABITS == 16) *reinterpret_cast<uint4*>(xlo + u * 128 + phys * 16) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
                s16_arr[2 * blk + half] = vs;
                if (half == 0) {
                    const float sx = amax / (ABITS == 16 ? ACT16_RANGE : ACT8_RANGE);
                    sx_arr[blk] = sx;
                    sm_arr[blk] = sx * (float)(vs + vs_other);
                }
            }
        }
    }
    named_bar_sync(1, NCT);

    // ---- this lane's slice of x into registers --------------------------------------------------
    const int nu = K / UNIT_COLS;
    const int wpr = warps_per_row(K);
    const int rpp = NCW / wpr;
    const int row_slot = warp / wpr;
    const int wsub = warp % wpr;
    const int u = wsub * 32 + lane;
    const bool valid = u < nu;
    XUnit xr;
    if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int phys = j ^ (u & 7);
            const uint4 h = *reinterpret_cast<const uint4*>(xhi + u * 128 + phys * 16);
            xr.hi[4 * j] = h.x; xr.hi[4 * j + 1] = h.y; xr.hi[4 * j + 2] = h.z; xr.hi[4 * j + 3] = h.w;
            if (ABITS == 16) {
                const uint4 l = *reinterpret_cast<const uint4*>(xlo + u * 128 + phys * 16);
                xr.lo[4 * j] = l.x; xr.lo[4 * j + 1] = l.y; xr.lo[4 * j + 2] = l.z; xr.lo[4 * j + 3] = l.w;
            }
        }
        const float4 a = *reinterpret_cast<const float4*>(sx_arr + 4 * u);
        const float4 b = *reinterpret_cast<const float4*>(sm_arr + 4 * u);
        xr.sx[0] = a.x; xr.sx[1] = a.y; xr.sx[2] = a.z; xr.sx[3] = a.w;
        xr.sm[0] = b.x; xr.sm[1] = b.y; xr.sm[2] = b.z; xr.sm[3] = b.w;
        const int4 c0 = *reinterpret_cast<const int4*>(s16_arr + 8 * u);
        const int4 c1 = *reinterpret_cast<const int4*>(s16_arr + 8 * u + 4);
        xr.s16[0] = c0.x; xr.s16[1] = c0.y; xr.s16[2] = c0.z; xr.s16[3] = c0.w;
        xr.s16[4] = c1.x; xr.s16[5] = c1.y; xr.s16[6] = c1.z; xr.s16[7] = c1.w;
    }

    // ---- main loop over this CTA's stages --------------------------------------------------------
    int st = 0;
    uint32_t ph = 0;
    int buf = 0;
    for (int s = 0; s < nwork; ++s) {
        const GemvSeg sg = p.seg[s];
        const WorkRange wr = cta_range(sg.rows, gran);
        for (int r0 = wr.a; r0 < wr.b; r0 += sg.rows_per_stage) {
            const int n = min(sg.rows_per_stage, wr.b - r0);
            const int ntot = p.pair ? 2 * n : n;
            const uint8_t* base = stages + (size_t)st * p.stage_bytes;
            mbar_wait(&full[st], ph);
            float* rb = res + buf * 256;
            const int npass = (ntot + rpp - 1) / rpp;
            switch (sg.type) {
                case T_Q4_K:
#pragma unroll 1
                    for (int ps = 0; ps < npass; ++ps) {
                        const int r = ps * rpp + row_slot;
                        float a = 0.f;
                        if (r < ntot && valid) a = unit_dot_q4k<ABITS>(base + (size_t)r * sg.row_stride + (size_t)(u >> 1) * 144, u & 1, xr);
                        a = warp_sum(a);
                        if (lane == 0 && r < ntot) rb[r * wpr + wsub] = a;
                    }
                    break;
                case T_Q6_K:
#pragma unroll 1
                    for (int ps = 0; ps < npass; ++ps) {
                        const int r = ps * rpp + row_slot;
                        float a = 0.f;
                        if (r < ntot && valid) a = unit_dot_q6k<ABITS>(base + (size_t)r * sg.row_stride, K >> 8, u, xr);
                        a = warp_sum(a);
                        if (lane == 0 && r < ntot) rb[r * wpr + wsub] = a;
                    }
                    break;
                default:
#pragma unroll 1
                    for (int ps = 0; ps < npass; ++ps) {
                        const int r = ps * rpp + row_slot;
                        float a = 0.f;
                        if (r < ntot && valid) a = unit_dot_q80<ABITS>(base + (size_t)r * sg.row_stride, K, u, xr);
                        a = warp_sum(a);
                        if (lane == 0 && r < ntot) rb[r * wpr + wsub] = a;
                    }
                    break;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);      // stage bytes are no longer needed by this warp
            if (++st == S) { st = 0; ph ^= 1; }
            named_bar_sync(1, NCT);
            buf ^= 1;

            // ---- epilogue -----------------------------------------------------------------------
            if (p.epi == EPI_QKV && s < 2) {
                if (tid < n / 2) {
                    float v0 = 0.f, v1 = 0.f;
                    for (int j = 0; j < wpr; ++j) { v0 += rb[(2 * tid) * wpr + j]; v1 += rb[(2 * tid + 1) * wpr + j]; }
                    const int r = r0 + 2 * tid;
                    const int pos = __ldcg(&p.st->pos);
                    const int d = r % p.head_dim;
                    const float c = p.rope_cos[(size_t)pos * (p.head_dim / 2) + d / 2];
                    const float sn = p.rope_sin[(size_t)pos * (p.head_dim / 2) + d / 2];
                    const float o0 = v0 * c - v1 * sn, o1 = v0 * sn + v1 * c;
                    if (s == 0) {
                        p.out[r] = o0;
                        p.out[r + 1] = o1;
                    } else {
                        const int kvh = r / p.head_dim;
                        const int page = __ldcg(p.page_table + pos / KV_PAGE_TOKENS);
                        const size_t off = (((size_t)page * p.n_kv_heads + kvh) * KV_PAGE_TOKENS + (pos % KV_PAGE_TOKENS)) * p.head_dim + d;
                        *reinterpret_cast<__half2*>(p.k_cache + off) = __floats2half2_rn(o0, o1);
                    }
                }
            } else if (tid < n) {
                float v = 0.f;
                for (int j = 0; j < wpr; ++j) v += rb[tid * wpr + j];
                const int r = r0 + tid;
                if (p.epi == EPI_STORE) {
                    p.out[r] = v;
                } else if (p.epi == EPI_ADD) {
                    p.out[r] = __ldcg(p.resid + r) + v;
                } else if (p.epi == EPI_SILU) {
                    float up = 0.f;
                    for (int j = 0; j < wpr; ++j) up += rb[(n + tid) * wpr + j];
                    p.out[r] = (v / (1.0f + expf(-v))) * up;
                } else {   // EPI_QKV, V segment
                    const int pos = __ldcg(&p.st->pos);
                    const int kvh = r / p.head_dim, d = r % p.head_dim;
                    const int page = __ldcg(p.page_table + pos / KV_PAGE_TOKENS);
                    const size_t off = (((size_t)page * p.n_kv_heads + kvh) * KV_PAGE_TOKENS + (pos % KV_PAGE_TOKENS)) * p.head_dim + d;
                    p.v_cache[off] = __float2half_rn(v);
                }
            }
        }
    }
}

}  // namespace

size_t gemv_smem_bytes(int cols, int n_stages, int stage_bytes) {
    size_t fixed = (SM_X + 2 * (size_t)cols + cols / 2 + 127) & ~(size_t)127;
    return fixed + (size_t)n_stages * stage_bytes;
}

bool gemv_plan(GemvParams& p) {
    if (p.cols % UNIT_COLS || p.cols <= 0 || p.cols > 32768) return false;
    if (p.n_stages < 2 || p.n_stages > GEMV_MAX_STAGES || (p.stage_bytes & 127)) return false;
    const int wpr = warps_per_row(p.cols);
    const int rpp = NCW / wpr;
    const int max_rows = rpp * MAXP;
    for (int s = 0; s < p.nseg; ++s) {
        GemvSeg& sg = p.seg[s];
        if (sg.type != T_Q4_K && sg.type != T_Q6_K && sg.type != T_Q8_0) return false;
        if ((sg.type == T_Q4_K || sg.type == T_Q6_K) && (p.cols % 256)) return false;
        if (sg.row_stride % 16 || ((uintptr_t)sg.w & 15)) return false;
        const int mult = p.pair ? 2 : 1;
        int r = p.stage_bytes / (sg.row_stride * mult);
        if (r > max_rows / mult) r = max_rows / mult;
        const int granule = rpp / mult > 0 ? rpp / mult : 1;
        if (r >= granule) r = r / granule * granule;
        if (p.epi == EPI_QKV) r &= ~1;
        if (r < ((p.epi == EPI_QKV) ? 2 : 1)) return false;
        sg.rows_per_stage = r;
        if (p.epi == EPI_QKV && (sg.rows & 1)) return false;
    }
    if (p.pair) {
        if (p.nseg != 2 || p.seg[0].type != p.seg[1].type || p.seg[0].rows != p.seg[1].rows ||
            p.seg[0].row_stride != p.seg[1].row_stride) return false;
        p.seg[1].rows_per_stage = p.seg[0].rows_per_stage;
    }
    return true;
}

cudaError_t gemv_configure() {
    cudaError_t e = cudaFuncSetAttribute(gemv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t gemv_launch(const GemvParams& p, int abits, int n_ctas, bool pdl, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3(GEMV_THREADS);
    cfg.dynamicSmemBytes = gemv_smem_bytes(p.cols, p.n_stages, p.stage_bytes);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    if (abits == 16) return cudaLaunchKernelEx(&cfg, gemv_kernel<16>, p);
    return cudaLaunchKernelEx(&cfg, gemv_kernel<8>, p);
}

// ------------------------------------------------------------------------------------------------
// fp-weight GEMV (F32 / F16 / BF16): warp per row, 16-B loads, fp32 FMA.  Not the headline path.
// ------------------------------------------------------------------------------------------------
namespace {
template <int TYPE>
__global__ void __launch_bounds__(256) gemv_fp_kernel(const uint8_t* __restrict__ w, int rows, int cols,
                                                      const float* __restrict__ x, float* __restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        float acc = 0.f;
        if (TYPE == T_F32) {
            const float* wr = reinterpret_cast<const float*>(w) + (size_t)r * cols;
            for (int c = lane * 4; c < cols; c += 128) {
                const float4 a = *reinterpret_cast<const float4*>(wr + c);
                const float4 b = *reinterpret_cast<const float4*>(x + c);
                acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
            }
        } else {
            const uint16_t* wr = reinterpret_cast<const uint16_t*>(w) + (size_t)r * cols;
            for (int c = lane * 8; c < cols; c += 256) {
                const uint4 a = *reinterpret_cast<const uint4*>(wr + c);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float f0, f1;
                    if (TYPE == T_F16) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw[q]));
                        f0 = f.x; f1 = f.y;
                    } else {
                        f0 = __uint_as_float(aw[q] << 16);
                        f1 = __uint_as_float(aw[q] & 0xFFFF0000u);
                    }
                    acc += f0 * x[c + 2 * q] + f1 * x[c + 2 * q + 1];
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) y[r] = acc;
    }
}
}  // namespace

cudaError_t gemv_fp_launch(const void* w, int type, int rows, int cols, const float* x, float* y, cudaStream_t s) {
    if (cols % 8) return cudaErrorInvalidValue;
    const int blocks = min((rows + 7) / 8, 148 * 8);
    const uint8_t* wb = static_cast<const uint8_t*>(w);
    switch (type) {
        case T_F32: gemv_fp_kernel<T_F32><<<blocks, 256, 0, s>>>(wb, rows, cols, x, y); break;
        case T_F16: gemv_fp_kernel<T_F16><<<blocks, 256, 0, s>>>(wb, rows, cols, x, y); break;
        case T_BF16: gemv_fp_kernel<T_BF16><<<blocks, 256, 0, s>>>(wb, rows, cols, x, y); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace gl
