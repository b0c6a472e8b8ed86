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
#include "gemv_core.cuh"

namespace gl {

namespace {

template <int ABITS, int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) gemv_kernel(const __grid_constant__ GemvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Ring ring;
    ring.full = reinterpret_cast<uint64_t*>(smem + SM_BARS);
    ring.empty = ring.full + GEMV_MAX_STAGES;
    ring.slots = smem + gemv_fixed_smem(p.cols);
    ring.n_slots = p.n_stages;
    ring.slot_bytes = p.stage_bytes;
    ring.st = 0;
    ring.ph = 0;
    if (tid == 0) {
        for (int i = 0; i < p.n_stages; ++i) {
            mbar_init(&ring.full[i], 1);
            mbar_init(&ring.empty[i], NW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();

    if (warp == NW) {
        // producer: weights are static, so streaming starts before the upstream kernel has finished
        if (lane == 0) gemv_produce(p, ring, blockIdx.x, gridDim.x);
        return;
    }
    pdl_wait();   // x (and the residual / KV pages we write) belong to the upstream kernel
    const float scale = gemv_prologue<ABITS, NW>(p, smem, tid);
    gemv_consume<ABITS, NW>(p, ring, smem, tid, scale, blockIdx.x, gridDim.x);
}

}  // namespace

size_t gemv_smem_bytes(int cols, int n_stages, int stage_bytes) {
    return (size_t)gemv_fixed_smem(cols) + (size_t)n_stages * stage_bytes;
}

bool gemv_plan(GemvParams& p, int consumer_warps) {
    if (p.cols % UNIT_COLS || p.cols <= 0 || p.cols > 32768) return false;
    if (p.n_stages < 2 || p.n_stages > GEMV_MAX_STAGES || (p.stage_bytes & 127)) return false;
    const int wpr = warps_per_row(p.cols);
    const int ngrp = consumer_warps / wpr;
    if (ngrp < 1) return false;
    for (int s = 0; s < p.nseg; ++s) {
        GemvSeg& sg = p.seg[s];
        if (sg.type != T_Q4_K && sg.type != T_Q6_K && sg.type != T_Q8_0) return false;
        if ((sg.type == T_Q4_K || sg.type == T_Q6_K) && (p.cols % 256)) return false;
        if (sg.row_stride % 16 || ((uintptr_t)sg.w & 15)) return false;
        const int mult = p.pair ? 2 : 1;
        const bool pair_adj = (p.epi == EPI_QKV && s < 2);
        int r = p.stage_bytes / (sg.row_stride * mult);
        if (r > 128) r = 128;
        // whole quads per stage: 4 rows (2 gate + 2 up rows in pair mode), and where it fits one quad per row group
        const int quad = p.pair ? 2 : 4;
        if (r >= quad * ngrp) r = r / (quad * ngrp) * (quad * ngrp);
        else if (r >= quad) r = r / quad * quad;
        if (pair_adj || p.epi == EPI_QKV) r &= ~1;
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

bool gemv_variant_ok(int abits, int nw) { return (abits == 16 || abits == 8) && nw == 8; }

cudaError_t gemv_configure() {
    cudaError_t e = cudaFuncSetAttribute(gemv_kernel<16, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemv_kernel<8, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemv_kernel<16, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemv_kernel<8, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    return e;
}

// ctas_per_sm selects the register budget the kernel was compiled for (1: unconstrained, 2: <= 112 registers)
cudaError_t gemv_launch(const GemvParams& p, int abits, int nw, int n_ctas, int ctas_per_sm, bool pdl, cudaStream_t s) {
    if (!gemv_variant_ok(abits, nw)) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3((unsigned)gemv_threads(nw));
    cfg.dynamicSmemBytes = gemv_smem_bytes(p.cols, p.n_stages, p.stage_bytes);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    if (ctas_per_sm >= 2)
        return abits == 16 ? cudaLaunchKernelEx(&cfg, gemv_kernel<16, 8, 2>, p) : cudaLaunchKernelEx(&cfg, gemv_kernel<8, 8, 2>, p);
    return abits == 16 ? cudaLaunchKernelEx(&cfg, gemv_kernel<16, 8, 1>, p) : cudaLaunchKernelEx(&cfg, gemv_kernel<8, 8, 1>, p);
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
