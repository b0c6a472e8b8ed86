// Continuous batching inside one engine: the small kernels of a batched decode step.
//
// Reference side: the worker drops a second assignment while busy (/root/reference/client/src/services/
// WorkerClientService.ts:500-505) and the server hands out one job per worker (MAX_CONCURRENT_JOBS_PER_WORKER,
// server/src/config/index.ts:31).  With that limit raised, B open sequences share ONE decode step here: the linear layers
// read the weights once for all of them (tensor-core GEMMs, engine_batch.cu), and the per-sequence pieces -- embedding
// gather, RoPE + KV append at each row's own position, paged attention over each row's own pages, sampling -- are the
// kernels below, one grid dimension over the rows of the step.
//
// The step lives in a CUDA graph per batch-size bucket, so WHO is in the batch is device-resident state: BatchCtl maps
// rows to sequence slots and is rewritten by the host only when the composition changes; rows >= n_rows leave at once.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace gl {

constexpr int MAX_BATCH = 128;          // sequence slots per engine = rows of one GEMM M tile

struct BatchCtl {
    int n_rows;                          // rows of this step (<= bucket <= MAX_BATCH)
    int row_slot[MAX_BATCH];             // row -> sequence slot (index into the StepState / page-table arrays)
};

// what the host reads back per row after a step (one 16-byte record per row, one D2H copy per step)
struct BatchOut {
    int token;                           // the token drawn this step (also the next step's input)
    float logprob;
    int done;                            // 1: the token drawn was a stop token (it is not part of the output)
    int pos;                             // position the NEXT step will process
};

struct BatchAttnParams {
    const float* q;                      // [rows][n_head][head_dim] fp32, rotated (null when the kernel rotates itself: qkv below)
    // fused RoPE + KV append (tensor-core kernel only): the raw QKV rows of the step.  The CTA rotates its own query heads; the
    // warp that owns the page of the newest position rotates that K row, converts the V row, patches its shared-memory copy of
    // the page and writes both rows to the cache.  One kernel and one dependent launch less per layer.
    const float* qkv;                    // [rows][ld_qkv] fp32: q | k | v, or null
    int ld_qkv;
    const float* cos_t;                  // [pos][head_dim / 2]
    const float* sin_t;
    const __half* k_cache;               // layer base, [page][kv][16][hd]
    const __half* v_cache;
    const int* tables;                   // [MAX_BATCH slots][table_stride] logical -> physical page
    int table_stride;
    const StepState* st;                 // [slots]
    const BatchCtl* ctl;
    __half* out16;                       // [rows][n_head * head_dim] fp16: the A operand of the attn_output GEMM
    float* part_o;                       // [rows][n_head][n_splits][head_dim]
    float* part_ml;                      // [rows][n_head][n_splits][2]
    unsigned* counters;                  // [rows][n_kv], zero between launches
    int n_head, n_kv_heads, head_dim, n_splits;
    float scale;
};

bool batch_pdl_enabled();              // GL_BATCH_PDL (default on): the per-layer kernels of the step use programmatic dependent launch
// ids[r] = token of row r's sequence (0 for rows beyond n_rows, so that padded rows stay finite)
cudaError_t batch_gather_tokens_launch(const BatchCtl* ctl, const StepState* st, int* ids, int bucket, cudaStream_t s);
// QKV rows fp32 [bucket x (qd + 2 kvd)]: RoPE at each row's own position, q out fp32, K / V appended to the row's own pages
cudaError_t batch_rope_kv_launch(const float* qkv, int bucket, const BatchCtl* ctl, const StepState* st, const int* tables, int table_stride,
                                 int n_head, int n_kv, int hd, const float* cos_t, const float* sin_t, float* q_out, __half* k_cache,
                                 __half* v_cache, cudaStream_t s);
cudaError_t batch_attn_configure();      // opt in to the tensor-core attention kernel's dynamic shared memory (once per device)
cudaError_t batch_attn_launch(const BatchAttnParams& p, int bucket, cudaStream_t s);
bool batch_attn_fuses_rope(int head_dim);    // tensor-core kernel in use (GL_BATCH_ATTN_MMA != 0) and GL_BATCH_FUSE_ROPE != 0
// greedy rows (temperature 0): argmax + log-softmax of the winner over logits [bucket][n_vocab]; advances the row's StepState
// and writes out_ids / out_lps [slot][max_out] exactly like the single-sequence sampler.  Rows with temperature > 0 are skipped
// (the host launches the seeded top-k sampler on them).
constexpr int BATCH_SAMPLE_PARTS = 8;                                   // CTAs per row
constexpr int BATCH_SAMPLE_ROW_FLOATS = 3 * BATCH_SAMPLE_PARTS + 1;     // scratch per row: per part (max, argmax, sum exp) + a ticket; zero at first
cudaError_t batch_sample_greedy_launch(const float* logits, int n_vocab, int bucket, const BatchCtl* ctl, StepState* st, int* out_ids,
                                       float* out_lps, int max_out, float* scratch /*[MAX_BATCH][BATCH_SAMPLE_ROW_FLOATS]*/, cudaStream_t s);
// out[r] = {token, logprob, done, pos} of row r after its sampler ran
cudaError_t batch_collect_launch(const BatchCtl* ctl, const StepState* st, const float* out_lps, int max_out, BatchOut* out, int bucket,
                                 cudaStream_t s);
// fp32 rows -> RMSNorm * w -> fp16 rows (rows beyond `rows` are left alone: they only ever hold finite values)
cudaError_t batch_rmsnorm_launch(const float* x, const float* w, int rows, int n, float eps, __half* y, cudaStream_t s);

}  // namespace gl
