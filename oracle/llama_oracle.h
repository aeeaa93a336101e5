/*
 * llama_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * PARITY UNPINNED against the reference binary: crowdllama's worker delegates the whole model
 * step to github.com/ollama/ollama v0.9.6 (/root/reference/go.mod:12, call site
 * /root/reference/pkg/crowdllama/api.go:108-160), which is not vendored, cannot be built here
 * (no Go toolchain, no network) and whose outputs no reference test pins
 * (/root/reference/test/integration_test.go:60-115 stubs it with an echo).  This file therefore
 * restates the PUBLIC llama-family algorithm (as realised by HF transformers
 * LlamaForCausalLM / llama.cpp's llama arch) and is pinned against transformers on CPU through
 * tests/golden/ (see tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libclengine.so) never links or calls it.
 *
 * Numerics contract "cl-llama v1" (shared with the CUDA engine):
 *   weights bf16; residual stream fp32; every GEMV/GEMM input vector rounded to bf16
 *   (round-to-nearest-even), fp32 accumulation, fp32 outputs; RMSNorm, RoPE, softmax, SiLU in
 *   fp32; K (post-RoPE), V and q (post-RoPE) rounded to bf16; logits fp32; greedy = argmax with
 *   lowest-index tie-break.
 */
#ifndef LLAMA_ORACLE_H_
#define LLAMA_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oc_config {
  int32_t n_layers, d_model, n_heads, n_kv_heads, head_dim, d_ff, vocab_size, max_seq_len;
  float rope_theta, rms_eps;
  /* "llama3" rotary scaling (HF transformers modeling_rope_utils.py _compute_llama3_parameters); rope_factor <= 1: none */
  float rope_factor, rope_low_freq_factor, rope_high_freq_factor;
  int32_t rope_original_max_pos;
} oc_config;

/* tensor kinds of the synthetic weight generator: key = layer * 16 + kind */
enum {
  OC_EMBED = 0, OC_LM_HEAD = 1, OC_FINAL_NORM = 2, OC_ATTN_NORM = 3, OC_WQ = 4, OC_WK = 5,
  OC_WV = 6, OC_WO = 7, OC_FFN_NORM = 8, OC_WGATE = 9, OC_WUP = 10, OC_WDOWN = 11
};
#define OC_LINEAR_SCALE 1.35e-4f          /* std(weight) ~= 0.02 */
#define OC_NORM_SCALE (1.0f / 4096.0f)    /* gain = 1 + s * OC_NORM_SCALE, s in [-510, 510] */

typedef struct oc_model oc_model;
typedef struct oc_seq oc_seq;

uint16_t oc_bf16_from_f32(float f);       /* round-to-nearest-even */
float oc_f32_from_bf16(uint16_t h);
float oc_bf16_round(float f);
/* counter-based generator: integer sum of 4 hash bytes minus 510, in [-510, 510] */
int32_t oc_synth_int(uint64_t seed, int32_t key, uint64_t index);
void oc_synth_bf16(uint64_t seed, int32_t key, uint64_t first, uint64_t n, float scale, uint16_t* out);

oc_model* oc_model_create(const oc_config* cfg);
void oc_model_destroy(oc_model* m);
void oc_model_fill_synthetic(oc_model* m, uint64_t seed);
/* layer ignored for EMBED / LM_HEAD / FINAL_NORM.  data is bf16 (norm gains too). */
int oc_model_set_tensor(oc_model* m, int32_t layer, int32_t kind, const uint16_t* data, int64_t n);
/* 1 (default): cl-llama v1 bf16 rounding points.  0: pure fp32 activations / KV (HF pin). */
void oc_model_set_act_rounding(oc_model* m, int32_t mode);
void oc_set_num_threads(int32_t n);
int32_t oc_get_num_threads(void);

oc_seq* oc_seq_create(oc_model* m, int32_t max_len);
void oc_seq_destroy(oc_seq* s);
int32_t oc_seq_len(const oc_seq* s);
void oc_seq_truncate(oc_seq* s, int32_t len);
/* timing-only helper: pretend `len` tokens are cached (fills K/V with a cheap pattern) */
void oc_seq_fake_fill(oc_seq* s, int32_t len);

/* Append n tokens.  all_logits==0: logits[vocab] of the last position; else logits[n][vocab]. */
int oc_forward(oc_model* m, oc_seq* s, const int32_t* ids, int32_t n, float* logits, int32_t all_logits);
/* Append n tokens layer by layer (weight rows reused over tiles of tokens): bit-identical to oc_forward with
 * all_logits == 0, but affordable for prompts of thousands of tokens.  logits[vocab] of the last position (may be NULL). */
int oc_prefill_block(oc_model* m, oc_seq* s, const int32_t* ids, int32_t n, float* logits);
/* cached K (which = 0) / V (which = 1) of `layer`, tokens t0..t0+n-1: out[n][n_kv*head_dim] */
int oc_seq_kv(const oc_seq* s, int32_t layer, int32_t which, int32_t t0, int32_t n, float* out);
/* residual stream of the last processed token after `layer` layers (0..n_layers) */
int oc_debug_hidden(oc_model* m, int32_t layer, float* out);
int oc_debug_vec(oc_model* m, int32_t which, float* out);
int32_t oc_argmax(const float* logits, int32_t n);
/* greedy loop: feeds first_id, then each argmax; ids_out[n_steps]; margins_out (may be NULL)
 * receives top1-top2 logit gaps */
int oc_greedy(oc_model* m, oc_seq* s, int32_t first_id, int32_t n_steps, int32_t* ids_out, float* margins_out);

/* sampler mirror (same counter-based RNG as the engine) */
typedef struct oc_sampling {
  float temperature; int32_t top_k; float top_p; float repeat_penalty; int32_t repeat_last_n;
  uint64_t seed;
} oc_sampling;
int32_t oc_sample(const float* logits, int32_t vocab, const oc_sampling* sp, const int32_t* history,
                  int32_t n_history, uint64_t step);

/* single ops (kernel-level parity) */
void oc_op_gemv(const uint16_t* w, const float* x, float* y, int32_t n_rows, int32_t k);
void oc_op_rmsnorm(const float* h, const float* gain, float eps, int32_t n, int32_t round_bf16, float* out);
void oc_op_rope(float* v, int32_t n_heads, int32_t head_dim, int32_t pos, float theta);
/* q [n_heads*hd] (roped, rounded), kc/vc [ctx][n_kv][hd] float; out [n_heads*hd] */
void oc_op_attention(const float* q, const float* kc, const float* vc, int32_t ctx, int32_t n_heads,
                     int32_t n_kv, int32_t head_dim, float* out);
/* Y[t][n] = sum_k X[t][k] W[n][k] (bf16 inputs, fp32 accumulate) */
void oc_op_gemm(const uint16_t* x, const uint16_t* w, float* y, int32_t t, int32_t n, int32_t k);

#ifdef __cplusplus
}
#endif
#endif
