// Persistent decode-step kernel (decode_mega.cu): phase table + launch interface.
#pragma once
#include "kernels.h"

namespace gl {

enum MegaPhaseKind : int { PH_GEMV = 1, PH_ATTN = 2 };
enum MegaPhaseFlags : int { PHF_HEAD = 1 };   // lm_head: also produce per-CTA softmax statistics

struct alignas(16) MegaPhase {
    int kind;
    int flags;
    int pad[2];
    GemvParams g;          // PH_GEMV: the whole work description; PH_ATTN: k_cache / v_cache of the layer
};

// The producer lane walks one ProdDesc (kernels.h) per GEMV phase.  They live in the kernel-parameter constant bank
// so that it never waits on a global-memory round trip when it crosses a phase boundary.
constexpr int MEGA_MAX_GEMV_PHASES = 400;     // 80 layers x 4 + head; 400 x 64 B = 25.6 KB of the 32 KB parameter space

struct MegaParams {
    const MegaPhase* phases;   // device memory, n_phases entries (one token)
    int n_phases;
    int n_steps;               // tokens per launch
    int with_head;             // 1: lm_head + greedy sample each step; 0: sequential-prefill step (pos += 1)
    StepState* st;
    unsigned* bar_counter;     // grid barrier, monotonic (StepState::bar_base carries the epoch)
    const int* prompt_ids;
    // embedding gather (token_embd in native GGUF layout)
    const uint8_t* embd_w;
    int embd_type, embd_row_bytes, n_embd;
    float* x;                  // residual stream [n_embd]
    // attention
    const float* q;
    float* attn_out;
    float* part_o;
    float* part_ml;
    unsigned* attn_counters;
    const int* page_table;
    int n_head, n_kv, head_dim, attn_splits;
    float attn_scale;
    // sampling
    const float* logits;
    float* head_part;          // [n_ctas][4]: max, argmax (int bits), sum exp
    int* out_ids;
    float* out_logprobs;
    float* logits_keep;
    int max_out;
    // ring
    int n_tracks, depth, slot_bytes, max_cols;   // ring: n_tracks x depth slots (gemv_core.cuh)
    unsigned long long* trace; // optional [n_ctas][n_phases+1][4] globaltimer stamps of the LAST step (GL_MEGA_TRACE=1)
    int n_prod;                // GEMV phases per token (entries of prod[] in use)
    ProdDesc prod[MEGA_MAX_GEMV_PHASES];
};

size_t mega_smem_bytes(int max_cols, int n_slots, int slot_bytes);
cudaError_t mega_configure();
cudaError_t mega_launch(const MegaParams& mp, int abits, int consumer_warps, int n_ctas, cudaStream_t s);

}  // namespace gl
