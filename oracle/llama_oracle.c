/*
 * llama_oracle.c — CPU ORACLE for the llama-family token step.  TEST INFRASTRUCTURE ONLY.
 * PARITY UNPINNED against the reference binary (see llama_oracle.h for why); pinned against
 * HF transformers on CPU via tests/golden/.
 *
 * What it restates (the arithmetic behind /root/reference/pkg/crowdllama/api.go:108-160, i.e.
 * the decode loop of github.com/ollama/ollama v0.9.6, /root/reference/go.mod:12, as the public
 * llama architecture): embedding gather -> L x [RMSNorm -> q,k,v projections -> RoPE
 * (rotate-half pairing i <-> i + d/2, the HF layout) -> KV append -> causal GQA attention ->
 * o projection + residual -> RMSNorm -> SwiGLU MLP + residual] -> final RMSNorm -> LM head ->
 * sampler.
 *
 * Build: see oracle/Makefile (gcc -O3 -mavx2 -mfma -fopenmp).
 */
#include "llama_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#ifdef __AVX2__
#include <immintrin.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* bf16 helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */
uint16_t oc_bf16_from_f32(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* quiet NaN */
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}
float oc_f32_from_bf16(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
float oc_bf16_round(float f) { return oc_f32_from_bf16(oc_bf16_from_f32(f)); }

/* ------------------------------------------------------------------------------------------ */
/* counter-based synthetic weights (bit-identical in C, numpy and the CUDA generator)          */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
int32_t oc_synth_int(uint64_t seed, int32_t key, uint64_t index) {
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + (((uint64_t)(uint32_t)key << 40) | index);
  uint32_t r = (uint32_t)mix64(x);
  return (int32_t)((r & 0xff) + ((r >> 8) & 0xff) + ((r >> 16) & 0xff) + (r >> 24)) - 510;
}
void oc_synth_bf16(uint64_t seed, int32_t key, uint64_t first, uint64_t n, float scale, uint16_t* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i)
    out[i] = oc_bf16_from_f32((float)oc_synth_int(seed, key, first + (uint64_t)i) * scale);
}

/* ------------------------------------------------------------------------------------------ */
/* model                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct oc_layer {
  float* attn_norm; float* ffn_norm;
  uint16_t *wq, *wk, *wv, *wo, *wg, *wu, *wd;
} oc_layer;

struct oc_model {
  oc_config c;
  int act_rounding;
  uint16_t* embed; uint16_t* lm_head; float* final_norm;
  oc_layer* layers;
  float* rope;            /* [max_seq_len][head_dim/2][2] cos,sin */
  float* dbg_hidden;      /* [n_layers+1][d_model] of the last processed token */
  /* scratch */
  float *h, *xn, *q, *k, *v, *att, *g, *u, *act, *scores, *tmp;
};

struct oc_seq {
  oc_model* m;
  int32_t max_len, len;
  float* kc; float* vc;   /* [n_layers][max_len][n_kv][hd] */
};

static void* xcalloc(size_t n, size_t sz) {
  void* p = calloc(n ? n : 1, sz);
  if (!p) { fprintf(stderr, "oracle: out of memory (%zu x %zu)\n", n, sz); abort(); }
  return p;
}

void oc_set_num_threads(int32_t n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int32_t oc_get_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Inverse frequency of dimension pair i.  Plain: theta^(-2i/hd).  With "llama3" scaling (Llama-3.1 / 3.2; HF transformers
 * modeling_rope_utils.py, _compute_llama3_parameters): wavelengths beyond original_max_pos / low_freq_factor are
 * stretched by `factor`, those below original_max_pos / high_freq_factor are kept, the band in between is blended. */
static double rope_inv_freq(const oc_config* c, int i) {
  const double pi = 3.14159265358979323846;
  double inv = pow((double)c->rope_theta, -2.0 * (double)i / (double)c->head_dim);
  if (!(c->rope_factor > 1.0f) || c->rope_original_max_pos <= 0) return inv;
  const double factor = c->rope_factor, lo = c->rope_low_freq_factor, hi = c->rope_high_freq_factor, old = c->rope_original_max_pos;
  const double wavelen = 2.0 * pi / inv;
  if (wavelen > old / lo) return inv / factor;
  if (wavelen < old / hi) return inv;
  const double smooth = (old / wavelen - lo) / (hi - lo);
  return (1.0 - smooth) * inv / factor + smooth * inv;
}

/* cos/sin table computed in double on the host, rounded to fp32.  The engine builds the same
 * table with the same expression on the host and uploads it, so both sides share it exactly. */
static void build_rope(float* tab, const oc_config* c) {
  int half = c->head_dim / 2, max_len = c->max_seq_len;
  for (int p = 0; p < max_len; ++p)
    for (int i = 0; i < half; ++i) {
      double inv = rope_inv_freq(c, i);
      double a = (double)p * inv;
      tab[((size_t)p * half + i) * 2 + 0] = (float)cos(a);
      tab[((size_t)p * half + i) * 2 + 1] = (float)sin(a);
    }
}

oc_model* oc_model_create(const oc_config* cfg) {
  oc_model* m = (oc_model*)xcalloc(1, sizeof(oc_model));
  m->c = *cfg;
  m->act_rounding = 1;
  const oc_config* c = &m->c;
  size_t d = c->d_model, qd = (size_t)c->n_heads * c->head_dim, kvd = (size_t)c->n_kv_heads * c->head_dim;
  m->embed = (uint16_t*)xcalloc((size_t)c->vocab_size * d, 2);
  m->lm_head = (uint16_t*)xcalloc((size_t)c->vocab_size * d, 2);
  m->final_norm = (float*)xcalloc(d, 4);
  m->layers = (oc_layer*)xcalloc(c->n_layers, sizeof(oc_layer));
  for (int l = 0; l < c->n_layers; ++l) {
    oc_layer* L = &m->layers[l];
    L->attn_norm = (float*)xcalloc(d, 4);
    L->ffn_norm = (float*)xcalloc(d, 4);
    L->wq = (uint16_t*)xcalloc(qd * d, 2);
    L->wk = (uint16_t*)xcalloc(kvd * d, 2);
    L->wv = (uint16_t*)xcalloc(kvd * d, 2);
    L->wo = (uint16_t*)xcalloc(d * qd, 2);
    L->wg = (uint16_t*)xcalloc((size_t)c->d_ff * d, 2);
    L->wu = (uint16_t*)xcalloc((size_t)c->d_ff * d, 2);
    L->wd = (uint16_t*)xcalloc(d * (size_t)c->d_ff, 2);
  }
  m->rope = (float*)xcalloc((size_t)c->max_seq_len * (c->head_dim / 2) * 2, 4);
  build_rope(m->rope, c);
  m->dbg_hidden = (float*)xcalloc((size_t)(c->n_layers + 1) * d, 4);
  m->h = (float*)xcalloc(d, 4); m->xn = (float*)xcalloc(d, 4);
  m->q = (float*)xcalloc(qd, 4); m->k = (float*)xcalloc(kvd, 4); m->v = (float*)xcalloc(kvd, 4);
  m->att = (float*)xcalloc(qd, 4);
  m->g = (float*)xcalloc(c->d_ff, 4); m->u = (float*)xcalloc(c->d_ff, 4); m->act = (float*)xcalloc(c->d_ff, 4);
  m->scores = (float*)xcalloc((size_t)c->n_heads * c->max_seq_len, 4);
  m->tmp = (float*)xcalloc(d, 4);
  return m;
}

void oc_model_destroy(oc_model* m) {
  if (!m) return;
  for (int l = 0; l < m->c.n_layers; ++l) {
    oc_layer* L = &m->layers[l];
    free(L->attn_norm); free(L->ffn_norm); free(L->wq); free(L->wk); free(L->wv); free(L->wo);
    free(L->wg); free(L->wu); free(L->wd);
  }
  free(m->layers); free(m->embed); free(m->lm_head); free(m->final_norm); free(m->rope);
  free(m->dbg_hidden); free(m->h); free(m->xn); free(m->q); free(m->k); free(m->v); free(m->att);
  free(m->g); free(m->u); free(m->act); free(m->scores); free(m->tmp);
  free(m);
}

void oc_model_set_act_rounding(oc_model* m, int32_t mode) { m->act_rounding = mode ? 1 : 0; }

static void synth_norm(uint64_t seed, int key, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = 1.0f + (float)oc_synth_int(seed, key, (uint64_t)i) * OC_NORM_SCALE;
}

void oc_model_fill_synthetic(oc_model* m, uint64_t seed) {
  const oc_config* c = &m->c;
  uint64_t d = c->d_model, qd = (uint64_t)c->n_heads * c->head_dim, kvd = (uint64_t)c->n_kv_heads * c->head_dim;
  oc_synth_bf16(seed, OC_EMBED, 0, (uint64_t)c->vocab_size * d, OC_LINEAR_SCALE, m->embed);
  oc_synth_bf16(seed, OC_LM_HEAD, 0, (uint64_t)c->vocab_size * d, OC_LINEAR_SCALE, m->lm_head);
  synth_norm(seed, OC_FINAL_NORM, (int)d, m->final_norm);
  for (int l = 0; l < c->n_layers; ++l) {
    oc_layer* L = &m->layers[l];
    int b = l * 16;
    synth_norm(seed, b + OC_ATTN_NORM, (int)d, L->attn_norm);
    synth_norm(seed, b + OC_FFN_NORM, (int)d, L->ffn_norm);
    oc_synth_bf16(seed, b + OC_WQ, 0, qd * d, OC_LINEAR_SCALE, L->wq);
    oc_synth_bf16(seed, b + OC_WK, 0, kvd * d, OC_LINEAR_SCALE, L->wk);
    oc_synth_bf16(seed, b + OC_WV, 0, kvd * d, OC_LINEAR_SCALE, L->wv);
    oc_synth_bf16(seed, b + OC_WO, 0, d * qd, OC_LINEAR_SCALE, L->wo);
    oc_synth_bf16(seed, b + OC_WGATE, 0, (uint64_t)c->d_ff * d, OC_LINEAR_SCALE, L->wg);
    oc_synth_bf16(seed, b + OC_WUP, 0, (uint64_t)c->d_ff * d, OC_LINEAR_SCALE, L->wu);
    oc_synth_bf16(seed, b + OC_WDOWN, 0, d * (uint64_t)c->d_ff, OC_LINEAR_SCALE, L->wd);
  }
}

int oc_model_set_tensor(oc_model* m, int32_t layer, int32_t kind, const uint16_t* data, int64_t n) {
  const oc_config* c = &m->c;
  int64_t d = c->d_model, qd = (int64_t)c->n_heads * c->head_dim, kvd = (int64_t)c->n_kv_heads * c->head_dim;
  uint16_t* dst16 = NULL; float* dstf = NULL; int64_t want = 0;
  if (kind == OC_EMBED) { dst16 = m->embed; want = (int64_t)c->vocab_size * d; }
  else if (kind == OC_LM_HEAD) { dst16 = m->lm_head; want = (int64_t)c->vocab_size * d; }
  else if (kind == OC_FINAL_NORM) { dstf = m->final_norm; want = d; }
  else {
    if (layer < 0 || layer >= c->n_layers) return -1;
    oc_layer* L = &m->layers[layer];
    switch (kind) {
      case OC_ATTN_NORM: dstf = L->attn_norm; want = d; break;
      case OC_FFN_NORM: dstf = L->ffn_norm; want = d; break;
      case OC_WQ: dst16 = L->wq; want = qd * d; break;
      case OC_WK: dst16 = L->wk; want = kvd * d; break;
      case OC_WV: dst16 = L->wv; want = kvd * d; break;
      case OC_WO: dst16 = L->wo; want = d * qd; break;
      case OC_WGATE: dst16 = L->wg; want = (int64_t)c->d_ff * d; break;
      case OC_WUP: dst16 = L->wu; want = (int64_t)c->d_ff * d; break;
      case OC_WDOWN: dst16 = L->wd; want = d * (int64_t)c->d_ff; break;
      default: return -1;
    }
  }
  if (n != want) return -2;
  if (dst16) memcpy(dst16, data, (size_t)n * 2);
  else for (int64_t i = 0; i < n; ++i) dstf[i] = oc_f32_from_bf16(data[i]);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* kernels                                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* dot(bf16 row, fp32 x): 16 lane accumulators (lane j takes elements == j mod 16), fused
 * multiply-add, fixed reduction tree -> identical result for the AVX2 and the scalar build
 * and for any thread count. */
static inline float dot_bf16(const uint16_t* w, const float* x, int k) {
  int k16 = k & ~15;
#ifdef __AVX2__
  __m256 a0 = _mm256_setzero_ps(), a1 = _mm256_setzero_ps();
  for (int i = 0; i < k16; i += 16) {
    __m128i r0 = _mm_loadu_si128((const __m128i*)(w + i));
    __m128i r1 = _mm_loadu_si128((const __m128i*)(w + i + 8));
    __m256 w0 = _mm256_castsi256_ps(_mm256_slli_epi32(_mm256_cvtepu16_epi32(r0), 16));
    __m256 w1 = _mm256_castsi256_ps(_mm256_slli_epi32(_mm256_cvtepu16_epi32(r1), 16));
    a0 = _mm256_fmadd_ps(w0, _mm256_loadu_ps(x + i), a0);
    a1 = _mm256_fmadd_ps(w1, _mm256_loadu_ps(x + i + 8), a1);
  }
  float lane[16];
  _mm256_storeu_ps(lane, a0);
  _mm256_storeu_ps(lane + 8, a1);
#else
  float lane[16] = {0};
  for (int i = 0; i < k16; i += 16)
    for (int j = 0; j < 16; ++j) lane[j] = fmaf(oc_f32_from_bf16(w[i + j]), x[i + j], lane[j]);
#endif
  for (int s = 8; s >= 1; s >>= 1)
    for (int j = 0; j < s; ++j) lane[j] += lane[j + s];
  float acc = lane[0];
  for (int i = k16; i < k; ++i) acc = fmaf(oc_f32_from_bf16(w[i]), x[i], acc);
  return acc;
}

void oc_op_gemv(const uint16_t* w, const float* x, float* y, int32_t n_rows, int32_t k) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n_rows; ++r) y[r] = dot_bf16(w + (size_t)r * k, x, k);
}

void oc_op_gemm(const uint16_t* x, const uint16_t* w, float* y, int32_t t, int32_t n, int32_t k) {
  float* xf = (float*)xcalloc((size_t)t * k, 4);
  for (size_t i = 0; i < (size_t)t * k; ++i) xf[i] = oc_f32_from_bf16(x[i]);
#pragma omp parallel for schedule(static) collapse(2)
  for (int ti = 0; ti < t; ++ti)
    for (int r = 0; r < n; ++r) y[(size_t)ti * n + r] = dot_bf16(w + (size_t)r * k, xf + (size_t)ti * k, k);
  free(xf);
}

void oc_op_rmsnorm(const float* h, const float* gain, float eps, int32_t n, int32_t round_bf16, float* out) {
  double ss = 0.0;
  for (int i = 0; i < n; ++i) ss += (double)h[i] * (double)h[i];
  float inv = 1.0f / sqrtf((float)(ss / (double)n) + eps);
  for (int i = 0; i < n; ++i) {
    float v = h[i] * inv * gain[i];
    out[i] = round_bf16 ? oc_bf16_round(v) : v;
  }
}

static void rope_with_table(float* v, int n_heads, int hd, const float* tab_pos) {
  int half = hd / 2;
  for (int hh = 0; hh < n_heads; ++hh) {
    float* p = v + (size_t)hh * hd;
    for (int i = 0; i < half; ++i) {
      float c = tab_pos[2 * i], s = tab_pos[2 * i + 1];
      float a = p[i], b = p[i + half];
      p[i] = a * c - b * s;
      p[i + half] = b * c + a * s;
    }
  }
}

void oc_op_rope(float* v, int32_t n_heads, int32_t head_dim, int32_t pos, float theta) {
  float* tab = (float*)xcalloc((size_t)head_dim, 4);
  int half = head_dim / 2;
  for (int i = 0; i < half; ++i) {
    double inv = pow((double)theta, -2.0 * (double)i / (double)head_dim);
    double a = (double)pos * inv;
    tab[2 * i] = (float)cos(a);
    tab[2 * i + 1] = (float)sin(a);
  }
  rope_with_table(v, n_heads, head_dim, tab);
  free(tab);
}

/* causal GQA attention of ONE query position against ctx cached positions */
static void attention_core(const float* q, const float* kc, const float* vc, size_t tok_stride, int ctx,
                           int n_heads, int n_kv, int hd, float* scores, float* out) {
  int rep = n_heads / n_kv;
  float scale = 1.0f / sqrtf((float)hd);
#pragma omp parallel for schedule(static)
  for (int hh = 0; hh < n_heads; ++hh) {
    int kvh = hh / rep;
    const float* qh = q + (size_t)hh * hd;
    float* sc = scores + (size_t)hh * ctx;
    float mx = -INFINITY;
    for (int t = 0; t < ctx; ++t) {
      const float* kt = kc + (size_t)t * tok_stride + (size_t)kvh * hd;
      float acc = 0.f;
      for (int i = 0; i < hd; ++i) acc = fmaf(qh[i], kt[i], acc);
      acc *= scale;
      sc[t] = acc;
      if (acc > mx) mx = acc;
    }
    double den = 0.0;
    for (int t = 0; t < ctx; ++t) { float e = expf(sc[t] - mx); sc[t] = e; den += (double)e; }
    float inv = (float)(1.0 / den);
    float* oh = out + (size_t)hh * hd;
    for (int i = 0; i < hd; ++i) oh[i] = 0.f;
    for (int t = 0; t < ctx; ++t) {
      const float* vt = vc + (size_t)t * tok_stride + (size_t)kvh * hd;
      float p = sc[t] * inv;
      for (int i = 0; i < hd; ++i) oh[i] = fmaf(p, vt[i], oh[i]);
    }
  }
}

void oc_op_attention(const float* q, const float* kc, const float* vc, int32_t ctx, int32_t n_heads,
                     int32_t n_kv, int32_t head_dim, float* out) {
  float* sc = (float*)xcalloc((size_t)n_heads * ctx, 4);
  attention_core(q, kc, vc, (size_t)n_kv * head_dim, ctx, n_heads, n_kv, head_dim, sc, out);
  free(sc);
}

/* ------------------------------------------------------------------------------------------ */
/* sequences and the token step                                                               */
/* ------------------------------------------------------------------------------------------ */
oc_seq* oc_seq_create(oc_model* m, int32_t max_len) {
  oc_seq* s = (oc_seq*)xcalloc(1, sizeof(oc_seq));
  s->m = m;
  s->max_len = max_len > m->c.max_seq_len ? m->c.max_seq_len : max_len;
  size_t per = (size_t)m->c.n_layers * s->max_len * m->c.n_kv_heads * m->c.head_dim;
  s->kc = (float*)xcalloc(per, 4);
  s->vc = (float*)xcalloc(per, 4);
  return s;
}
void oc_seq_destroy(oc_seq* s) { if (s) { free(s->kc); free(s->vc); free(s); } }
int32_t oc_seq_len(const oc_seq* s) { return s->len; }
void oc_seq_truncate(oc_seq* s, int32_t len) { if (len >= 0 && len < s->len) s->len = len; }
void oc_seq_fake_fill(oc_seq* s, int32_t len) {
  if (len > s->max_len) len = s->max_len;
  size_t kvd = (size_t)s->m->c.n_kv_heads * s->m->c.head_dim;
  for (int l = 0; l < s->m->c.n_layers; ++l) {
    float* kb = s->kc + (size_t)l * s->max_len * kvd;
    float* vb = s->vc + (size_t)l * s->max_len * kvd;
    for (size_t i = 0; i < (size_t)len * kvd; ++i) {
      kb[i] = (float)((int)((i * 2654435761u) >> 24 & 0xff) - 128) * (1.0f / 128.0f);
      vb[i] = (float)((int)((i * 40503u) >> 8 & 0xff) - 128) * (1.0f / 128.0f);
    }
  }
  s->len = len;
}

static inline float rnd(const oc_model* m, float v) { return m->act_rounding ? oc_bf16_round(v) : v; }

static int step(oc_model* m, oc_seq* s, int32_t id, float* logits) {
  const oc_config* c = &m->c;
  int d = c->d_model, hd = c->head_dim, H = c->n_heads, KV = c->n_kv_heads;
  int qd = H * hd, kvd = KV * hd;
  if (id < 0 || id >= c->vocab_size) return -1;
  if (s->len >= s->max_len) return -2;
  int pos = s->len;
  const float* tab = m->rope + (size_t)pos * (hd / 2) * 2;
  for (int i = 0; i < d; ++i) m->h[i] = oc_f32_from_bf16(m->embed[(size_t)id * d + i]);
  memcpy(m->dbg_hidden, m->h, (size_t)d * 4);
  for (int l = 0; l < c->n_layers; ++l) {
    oc_layer* L = &m->layers[l];
    oc_op_rmsnorm(m->h, L->attn_norm, c->rms_eps, d, m->act_rounding, m->xn);
    oc_op_gemv(L->wq, m->xn, m->q, qd, d);
    oc_op_gemv(L->wk, m->xn, m->k, kvd, d);
    oc_op_gemv(L->wv, m->xn, m->v, kvd, d);
    rope_with_table(m->q, H, hd, tab);
    rope_with_table(m->k, KV, hd, tab);
    float* kb = s->kc + ((size_t)l * s->max_len + pos) * kvd;
    float* vb = s->vc + ((size_t)l * s->max_len + pos) * kvd;
    for (int i = 0; i < kvd; ++i) { kb[i] = rnd(m, m->k[i]); vb[i] = rnd(m, m->v[i]); }
    for (int i = 0; i < qd; ++i) m->q[i] = rnd(m, m->q[i]);
    attention_core(m->q, s->kc + (size_t)l * s->max_len * kvd, s->vc + (size_t)l * s->max_len * kvd,
                   (size_t)kvd, pos + 1, H, KV, hd, m->scores, m->att);
    for (int i = 0; i < qd; ++i) m->att[i] = rnd(m, m->att[i]);
    oc_op_gemv(L->wo, m->att, m->tmp, d, qd);
    for (int i = 0; i < d; ++i) m->h[i] += m->tmp[i];
    oc_op_rmsnorm(m->h, L->ffn_norm, c->rms_eps, d, m->act_rounding, m->xn);
    oc_op_gemv(L->wg, m->xn, m->g, c->d_ff, d);
    oc_op_gemv(L->wu, m->xn, m->u, c->d_ff, d);
    for (int i = 0; i < c->d_ff; ++i) {
      float gg = m->g[i];
      float si = gg / (1.0f + expf(-gg));
      m->act[i] = rnd(m, si * m->u[i]);
    }
    oc_op_gemv(L->wd, m->act, m->tmp, d, c->d_ff);
    for (int i = 0; i < d; ++i) m->h[i] += m->tmp[i];
    memcpy(m->dbg_hidden + (size_t)(l + 1) * d, m->h, (size_t)d * 4);
  }
  s->len = pos + 1;
  if (logits) {
    oc_op_rmsnorm(m->h, m->final_norm, c->rms_eps, d, m->act_rounding, m->xn);
    oc_op_gemv(m->lm_head, m->xn, logits, c->vocab_size, d);
  }
  return 0;
}

int oc_forward(oc_model* m, oc_seq* s, const int32_t* ids, int32_t n, float* logits, int32_t all_logits) {
  for (int i = 0; i < n; ++i) {
    float* lg = NULL;
    if (logits) {
      if (all_logits) lg = logits + (size_t)i * m->c.vocab_size;
      else if (i == n - 1) lg = logits;
    }
    int rc = step(m, s, ids[i], lg);
    if (rc) return rc;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* layer-major prefill: the SAME per-token arithmetic as step() (every y[t][r] is the same      */
/* dot_bf16 of the same operands, so results are bit-identical to feeding the tokens one by one */
/* through oc_forward), but each weight row is reused over a tile of tokens — what makes a      */
/* 4096-token prompt affordable on the CPU.  Used only to check the engine's long prefill.      */
/* ------------------------------------------------------------------------------------------ */
#define OC_TB 32
static void gemm_tiled(const uint16_t* w, const float* x, size_t xs, float* y, size_t ys, int t, int n_rows, int k) {
  for (int t0 = 0; t0 < t; t0 += OC_TB) {
    int t1 = t0 + OC_TB < t ? t0 + OC_TB : t;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n_rows; ++r)
      for (int ti = t0; ti < t1; ++ti) y[(size_t)ti * ys + r] = dot_bf16(w + (size_t)r * k, x + (size_t)ti * xs, k);
  }
}

int oc_prefill_block(oc_model* m, oc_seq* s, const int32_t* ids, int32_t n, float* logits) {
  const oc_config* c = &m->c;
  int d = c->d_model, hd = c->head_dim, H = c->n_heads, KV = c->n_kv_heads, F = c->d_ff;
  int qd = H * hd, kvd = KV * hd, pos0 = s->len;
  if (n <= 0) return -1;
  if (pos0 + n > s->max_len) return -2;
  for (int t = 0; t < n; ++t) if (ids[t] < 0 || ids[t] >= c->vocab_size) return -1;
  size_t wide = (size_t)(F > qd ? F : qd);
  if ((size_t)d > wide) wide = (size_t)d;
  float* h = (float*)xcalloc((size_t)n * d, 4);
  float* xn = (float*)xcalloc((size_t)n * wide, 4);      /* bf16-rounded GEMV inputs (norm out / attention out / SwiGLU act) */
  float* q = (float*)xcalloc((size_t)n * qd, 4);
  float* kk = (float*)xcalloc((size_t)n * kvd, 4);
  float* vv = (float*)xcalloc((size_t)n * kvd, 4);
  float* g = (float*)xcalloc((size_t)n * F, 4);
  float* u = (float*)xcalloc((size_t)n * F, 4);
  float* tmp = (float*)xcalloc((size_t)n * d, 4);
  for (int t = 0; t < n; ++t)
    for (int i = 0; i < d; ++i) h[(size_t)t * d + i] = oc_f32_from_bf16(m->embed[(size_t)ids[t] * d + i]);
  memcpy(m->dbg_hidden, h + (size_t)(n - 1) * d, (size_t)d * 4);
  for (int l = 0; l < c->n_layers; ++l) {
    oc_layer* L = &m->layers[l];
    float* kc = s->kc + (size_t)l * s->max_len * kvd;
    float* vc = s->vc + (size_t)l * s->max_len * kvd;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < n; ++t) oc_op_rmsnorm(h + (size_t)t * d, L->attn_norm, c->rms_eps, d, m->act_rounding, xn + (size_t)t * d);
    gemm_tiled(L->wq, xn, d, q, qd, n, qd, d);
    gemm_tiled(L->wk, xn, d, kk, kvd, n, kvd, d);
    gemm_tiled(L->wv, xn, d, vv, kvd, n, kvd, d);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < n; ++t) {
      const float* tab = m->rope + (size_t)(pos0 + t) * (hd / 2) * 2;
      rope_with_table(q + (size_t)t * qd, H, hd, tab);
      rope_with_table(kk + (size_t)t * kvd, KV, hd, tab);
      float* kb = kc + (size_t)(pos0 + t) * kvd;
      float* vb = vc + (size_t)(pos0 + t) * kvd;
      for (int i = 0; i < kvd; ++i) { kb[i] = rnd(m, kk[(size_t)t * kvd + i]); vb[i] = rnd(m, vv[(size_t)t * kvd + i]); }
      for (int i = 0; i < qd; ++i) q[(size_t)t * qd + i] = rnd(m, q[(size_t)t * qd + i]);
    }
    for (int t = 0; t < n; ++t) {   /* attention_core parallelises over heads */
      float* att = xn + (size_t)t * qd;
      attention_core(q + (size_t)t * qd, kc, vc, (size_t)kvd, pos0 + t + 1, H, KV, hd, m->scores, att);
      for (int i = 0; i < qd; ++i) att[i] = rnd(m, att[i]);
    }
    gemm_tiled(L->wo, xn, qd, tmp, d, n, d, qd);
    for (size_t i = 0; i < (size_t)n * d; ++i) h[i] += tmp[i];
#pragma omp parallel for schedule(static)
    for (int t = 0; t < n; ++t) oc_op_rmsnorm(h + (size_t)t * d, L->ffn_norm, c->rms_eps, d, m->act_rounding, xn + (size_t)t * d);
    gemm_tiled(L->wg, xn, d, g, F, n, F, d);
    gemm_tiled(L->wu, xn, d, u, F, n, F, d);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < n; ++t)
      for (int i = 0; i < F; ++i) {
        float gg = g[(size_t)t * F + i];
        float si = gg / (1.0f + expf(-gg));
        xn[(size_t)t * F + i] = rnd(m, si * u[(size_t)t * F + i]);
      }
    gemm_tiled(L->wd, xn, F, tmp, d, n, d, F);
    for (size_t i = 0; i < (size_t)n * d; ++i) h[i] += tmp[i];
    memcpy(m->dbg_hidden + (size_t)(l + 1) * d, h + (size_t)(n - 1) * d, (size_t)d * 4);
  }
  s->len = pos0 + n;
  memcpy(m->h, h + (size_t)(n - 1) * d, (size_t)d * 4);
  if (logits) {
    oc_op_rmsnorm(m->h, m->final_norm, c->rms_eps, d, m->act_rounding, m->xn);
    oc_op_gemv(m->lm_head, m->xn, logits, c->vocab_size, d);
  }
  free(h); free(xn); free(q); free(kk); free(vv); free(g); free(u); free(tmp);
  return 0;
}

/* cached K (which = 0) or V (which = 1) of `layer`, tokens t0 .. t0+n-1: out[n][n_kv*head_dim] */
int oc_seq_kv(const oc_seq* s, int32_t layer, int32_t which, int32_t t0, int32_t n, float* out) {
  const oc_config* c = &s->m->c;
  size_t kvd = (size_t)c->n_kv_heads * c->head_dim;
  if (layer < 0 || layer >= c->n_layers || t0 < 0 || n < 0 || t0 + n > s->len) return -1;
  const float* src = (which ? s->vc : s->kc) + ((size_t)layer * s->max_len + t0) * kvd;
  memcpy(out, src, (size_t)n * kvd * 4);
  return 0;
}

int oc_debug_hidden(oc_model* m, int32_t layer, float* out) {
  if (layer < 0 || layer > m->c.n_layers) return -1;
  memcpy(out, m->dbg_hidden + (size_t)layer * m->c.d_model, (size_t)m->c.d_model * 4);
  return 0;
}

/* last-layer intermediates of the most recent token: 0 = q (roped, rounded), 1 = attention output (rounded),
 * 2 = SwiGLU activation (rounded), 3 = normalised input of the last GEMV group (xn) */
int oc_debug_vec(oc_model* m, int32_t which, float* out) {
  const oc_config* c = &m->c;
  const float* src = which == 0 ? m->q : which == 1 ? m->att : which == 2 ? m->act : which == 3 ? m->xn : NULL;
  const size_t n = which == 0 || which == 1 ? (size_t)c->n_heads * c->head_dim : which == 2 ? (size_t)c->d_ff : (size_t)c->d_model;
  if (!src) return -1;
  memcpy(out, src, n * 4);
  return (int)n;
}

int32_t oc_argmax(const float* logits, int32_t n) {
  int best = 0;
  for (int i = 1; i < n; ++i) if (logits[i] > logits[best]) best = i;
  return best;
}

int oc_greedy(oc_model* m, oc_seq* s, int32_t first_id, int32_t n_steps, int32_t* ids_out, float* margins_out) {
  float* lg = (float*)xcalloc(m->c.vocab_size, 4);
  int32_t id = first_id;
  for (int i = 0; i < n_steps; ++i) {
    int rc = step(m, s, id, lg);
    if (rc) { free(lg); return rc; }
    id = oc_argmax(lg, m->c.vocab_size);
    ids_out[i] = id;
    if (margins_out) {
      float top = lg[id], second = -INFINITY;
      for (int j = 0; j < m->c.vocab_size; ++j) if (j != id && lg[j] > second) second = lg[j];
      margins_out[i] = top - second;
    }
  }
  free(lg);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* sampler (restates Ollama's default sampling chain: repeat penalty -> top-k -> temperature   */
/* -> softmax -> top-p -> draw; UPSTREAM defaults listed in SURVEY.md §8c)                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float v; int32_t id; } cand_t;
static int cand_cmp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->v > y->v) return -1;
  if (x->v < y->v) return 1;
  return (x->id > y->id) - (x->id < y->id);
}
static double uniform01(uint64_t seed, uint64_t step) {
  uint64_t x = mix64(seed * 0x9E3779B97F4A7C15ull + step + 0x632BE59BD9B4E019ull);
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

int32_t oc_sample(const float* logits, int32_t vocab, const oc_sampling* sp, const int32_t* history,
                  int32_t n_history, uint64_t step) {
  if (sp->temperature <= 0.f) return oc_argmax(logits, vocab);
  cand_t* c = (cand_t*)xcalloc(vocab, sizeof(cand_t));
  for (int i = 0; i < vocab; ++i) { c[i].v = logits[i]; c[i].id = i; }
  if (sp->repeat_penalty != 1.0f && sp->repeat_last_n != 0 && history) {
    int start = sp->repeat_last_n > 0 && n_history > sp->repeat_last_n ? n_history - sp->repeat_last_n : 0;
    char* seen = (char*)xcalloc(vocab, 1);
    for (int i = start; i < n_history; ++i) {
      int id = history[i];
      if (id < 0 || id >= vocab || seen[id]) continue;
      seen[id] = 1;
      c[id].v = c[id].v > 0.f ? c[id].v / sp->repeat_penalty : c[id].v * sp->repeat_penalty;
    }
    free(seen);
  }
  qsort(c, vocab, sizeof(cand_t), cand_cmp);
  int n = vocab;
  if (sp->top_k > 0 && sp->top_k < n) n = sp->top_k;
  float mx = c[0].v;
  double den = 0.0;
  double* p = (double*)xcalloc(n, sizeof(double));
  for (int i = 0; i < n; ++i) { p[i] = exp((double)(c[i].v - mx) / (double)sp->temperature); den += p[i]; }
  int keep = n;
  if (sp->top_p > 0.f && sp->top_p < 1.f) {
    double cum = 0.0;
    for (int i = 0; i < n; ++i) { cum += p[i] / den; if (cum >= (double)sp->top_p) { keep = i + 1; break; } }
  }
  double tot = 0.0;
  for (int i = 0; i < keep; ++i) tot += p[i];
  double u = uniform01(sp->seed, step) * tot, cum = 0.0;
  int32_t pick = c[keep - 1].id;
  for (int i = 0; i < keep; ++i) { cum += p[i]; if (u < cum) { pick = c[i].id; break; } }
  free(p); free(c);
  return pick;
}
