// prefill.cu — prompt (prefill) path of the engine: chunks of up to prefill_chunk_tokens_ tokens go
// through tcgen05 GEMMs (gemm_tcgen05.cu) and the causal paged attention kernel
// (prefill_kernels.cu); the last position's logits come from the decode LM-head GEMV so that
// prefill and decode share one logits/argmax tail.  Also: cl_op_gemm_bf16 / cl_op_attn_prefill.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "engine.h"

namespace cl {

bool Engine::prefill_path_ok() const {
  static int disabled = -1;
  if (disabled < 0) {
    const char* v = getenv("CL_PREFILL_TCGEN05");
    disabled = (v && *v == '0') ? 1 : 0;
  }
  return !disabled && gemm_tcgen05_supported(16, cfg.d_model, cfg.d_model);
}

// split-K rule shared with the batched decode step (engine.cu)
int pick_splits_public(int n_rows, int K);

// CL_PREFILL_PROFILE=1: a CUDA event after every launch, per-kernel-class device time printed to stderr (in-pipeline
// times incl. the gap before each kernel; the ncu launch list measures cold, serialised launches instead).
struct PrefillProfiler {
  bool on;
  cudaStream_t st;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> name;
  explicit PrefillProfiler(cudaStream_t s) : st(s) {
    static const bool prof = getenv("CL_PREFILL_PROFILE") && atoi(getenv("CL_PREFILL_PROFILE")) != 0;
    on = prof;
  }
  void mark(const char* n) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    name.push_back(n);
  }
  void report(int n_tokens, const char* path) {
    if (!on || ev.size() < 2) return;
    cudaStreamSynchronize(st);
    std::map<std::string, std::pair<double, int>> acc;
    double total = 0.0;
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      const std::string full(name[i]);
      std::string key = full.substr(0, full.find('('));
      if (key == "launch_gemm_bf16_epi") key += full.find("L.wqkv") != std::string::npos ? ":qkv+rope" : ":gate|up+silu";
      if (key == "launch_gemm_bf16")     // split by projection: the weight argument names it
        key += full.find("L.wqkv") != std::string::npos ? ":qkv" : full.find("L.wo") != std::string::npos ? ":o" : full.find("L.wgu") != std::string::npos ? ":gate|up" : ":down";
      acc[key].first += ms; acc[key].second += 1;
      total += ms;
    }
    fprintf(stderr, "[prefill profile] %s path, %d tokens, %zu launches, %.3f ms on the device\n", path, n_tokens, ev.size() - 1, total);
    for (auto& kv : acc) fprintf(stderr, "[prefill profile]   %-32s n=%4d total %8.3f ms  mean %8.2f us\n", kv.first.c_str(), kv.second.second, kv.second.first, 1e3 * kv.second.first / kv.second.second);
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear(); name.clear();
  }
};

// Short prompts (prefill_min_tokens_ <= n <= 256, e.g. a 128-token chat): one token tile per W row tile means only
// N/128 CTAs stream the weights of a projection (48 for q|k|v, 32 for o and down) — 11 ms for 128 tokens where the
// weight stream alone needs 2.4.  Same recipe as the batched decode step: split-K so that ~148 CTAs share every
// projection's weights, fp32 partials in a workspace, the per-row glue kernels (batch_kernels.cu) fold them in fixed
// order into the residual stream / RoPE / SiLU.  Rows are the prompt's tokens, attention is the causal prefill kernel.
int Engine::prefill_small(cl_seq_t s, const int32_t* ids, int n, float* logits_out) {
  auto& q = seqs_[s];
  int rc = ensure_capacity(s, q.len + n);
  if (rc) return rc;
  const int d = cfg.d_model, F = cfg.d_ff, T = n, pos0 = q.len;
  const int s_qkv = pick_splits_public(qkv_dim_, d), s_o = pick_splits_public(d, q_dim_), s_gu = pick_splits_public(2 * F, d), s_dn = pick_splits_public(d, F);
  if (!sws_) {
    sws_.reset(new SmallPrefillWs());
    const size_t Tm = 256;
    size_t part = 0;
    part = std::max(part, (size_t)s_qkv * Tm * qkv_dim_);
    part = std::max(part, (size_t)s_o * Tm * d);
    part = std::max(part, (size_t)s_gu * Tm * 2 * F);
    part = std::max(part, (size_t)s_dn * Tm * d);
    auto alloc = [&](auto*& p, size_t bytes) -> int {
      void* v = nullptr;
      if (cudaMalloc(&v, bytes) != cudaSuccess) { cudaGetLastError(); set_last_error("short-prompt prefill workspace: out of memory"); return CL_ERR_OOM; }
      allocs_.push_back(v);
      p = reinterpret_cast<std::remove_reference_t<decltype(p)>>(v);
      return CL_OK;
    };
    if ((rc = alloc(sws_->part, part * 4)) || (rc = alloc(sws_->xn, Tm * d * 2)) || (rc = alloc(sws_->q, Tm * q_dim_ * 2)) ||
        (rc = alloc(sws_->attn, Tm * q_dim_ * 2)) || (rc = alloc(sws_->h, Tm * d * 4)) || (rc = alloc(sws_->act, Tm * F * 2)) ||
        (rc = alloc(sws_->iota, Tm * 4))) { sws_.reset(); return rc; }
    std::vector<int> io(Tm);
    for (size_t i = 0; i < Tm; ++i) io[i] = (int)i;
    CL_CUDA_OK(cudaMemcpy(sws_->iota, io.data(), Tm * 4, cudaMemcpyHostToDevice));
  }
  SmallPrefillWs& w = *sws_;
  CL_CUDA_OK(cudaMemcpyAsync(d_prompt_, ids, (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
  const int* bt = d_bt_ + (size_t)s * max_pages_per_seq_;
  int launches = 0, r;
  PrefillProfiler pp(stream_);
#define CL_LAUNCH(call) do { r = (call); if (r < 0) { set_last_error(std::string(#call) + ": " + cudaGetErrorString(cudaGetLastError())); return CL_ERR_CUDA; } launches += r; pp.mark(#call); } while (0)
  // programmatic dependent launch between the GEMMs and the per-row glue kernels, as in the batched decode step (the two
  // prefill kernels in between are launched normally and order the stream)
  static const bool small_pdl = !getenv("CL_SMALL_PDL") || atoi(getenv("CL_SMALL_PDL")) != 0;   // 146 tokens: 6.41 -> 5.90 ms (r2q)
  const bool bp = small_pdl && use_pdl_;
  pp.mark("start");
  CL_LAUNCH(launch_embed_rows(embed_, d, d_prompt_, w.h, T, stream_));
  const float* pending = nullptr;   // split-K partials of the previous residual projection, folded in by the next norm
  int pending_s = 0;
  for (int l = 0; l < cfg.n_layers; ++l) {
    const auto& L = layers_[l];
    CL_LAUNCH(launch_batch_resid_norm(w.h, d, pending, pending_s, T, L.attn_norm, cfg.rms_eps, w.xn, w.iota, stream_, bp));
    CL_LAUNCH(launch_gemm_bf16(w.xn, L.wqkv, w.part, nullptr, T, qkv_dim_, d, stream_, s_qkv, bp));
    RopeScatterArgs ra{w.part, qkv_dim_, rope_, pos0, T, w.q, kpool_ + (size_t)l * kv_layer_elems_, vpool_ + (size_t)l * kv_layer_elems_,
                       bt, page_size_, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim};
    ra.n_split = s_qkv; ra.split_stride = (size_t)T * qkv_dim_;
    CL_LAUNCH(launch_rope_scatter(ra, stream_));
    AttnPrefillArgs aa{w.q, kpool_ + (size_t)l * kv_layer_elems_, vpool_ + (size_t)l * kv_layer_elems_, bt, page_size_,
                       pos0, T, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, w.attn};
    if (have_kv_maps_ && attn_prefill_tc_supported(cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, page_size_, pos0, T))
      CL_LAUNCH(launch_attn_prefill_tc(aa, kmap_, vmap_, (long long)l * n_pages_ * cfg.n_kv_heads * page_size_, stream_));
    else
      CL_LAUNCH(launch_attn_prefill(aa, stream_));
    CL_LAUNCH(launch_gemm_bf16(w.attn, L.wo, w.part, nullptr, T, d, q_dim_, stream_, s_o, bp));
    CL_LAUNCH(launch_batch_resid_norm(w.h, d, w.part, s_o, T, L.ffn_norm, cfg.rms_eps, w.xn, w.iota, stream_, bp));
    CL_LAUNCH(launch_gemm_bf16(w.xn, L.wgu, w.part, nullptr, T, 2 * F, d, stream_, s_gu, bp));
    CL_LAUNCH(launch_batch_silu(w.part, s_gu, T, F, w.act, stream_, bp));
    CL_LAUNCH(launch_gemm_bf16(w.act, L.wdown, w.part, nullptr, T, d, F, stream_, s_dn, bp));
    pending = w.part; pending_s = s_dn;
  }
  // fold the last down-projection into the residual rows (the normalised copy it also writes is not used), then the shared tail
  CL_LAUNCH(launch_batch_resid_norm(w.h, d, pending, pending_s, T, final_norm_, cfg.rms_eps, w.xn, w.iota, stream_, bp));
  CL_CUDA_OK(cudaMemcpyAsync(d_h_ + (size_t)s * d, w.h + (size_t)(T - 1) * d, (size_t)d * 4, cudaMemcpyDeviceToDevice, stream_));
  const int last_pos = q.len + n - 1;
  CL_CUDA_OK(cudaMemcpyAsync(d_pos_ + s, &last_pos, 4, cudaMemcpyHostToDevice, stream_));
  rc = set_single_slot(s);
  if (rc) return rc;
  GemvArgs lm;
  lm.slots = d_slots_; lm.batch = 1;
  lm.W = lm_head_; lm.N = cfg.vocab_size; lm.K = d; lm.h = d_h_; lm.gain = final_norm_; lm.eps = cfg.rms_eps;
  lm.y = d_logits_; lm.x_stride = d; lm.y_stride = cfg.vocab_size;
  CL_LAUNCH(launch_gemv(gemv_variant_, EPI_STORE, true, lm, stream_, false));
  StepTailArgs t;
  t.logits = d_logits_; t.vocab = cfg.vocab_size; t.tok = d_tok_; t.pos = d_pos_; t.ids_ring = d_ids_ring_;
  t.step_counter = d_step_counter_; t.ring_steps = ring_steps_; t.ring_stride = max_batch_;
  t.part_val = d_tail_val_; t.part_idx = d_tail_idx_; t.counters = d_tail_cnt_; t.slots = d_slots_; t.batch = 1;
  CL_LAUNCH(launch_step_tail(t, stream_));
#undef CL_LAUNCH
  pp.report(n, "short-prompt");
  launches_ += launches;
  q.len += n;
  q.history.insert(q.history.end(), ids, ids + n);
  if (logits_out) return read_logits(s, logits_out);
  CL_CUDA_OK(cudaStreamSynchronize(stream_));
  return CL_OK;
}

// (Re)allocate the tile-path workspace for chunks of up to `tokens` rows; old buffers stay in allocs_ until the engine dies
// (grows at most once or twice).
int Engine::ensure_prefill_ws(int tokens) {
  const int d = cfg.d_model, F = cfg.d_ff;
  int rc;
  if (!pws_) pws_.reset(new PrefillWs());
  if (pws_->cap_tokens >= tokens) return CL_OK;
  const size_t T = (size_t)std::max(tokens, std::min(prefill_chunk_tokens_, cfg.max_seq_len));
  auto alloc = [&](auto*& p, size_t bytes) -> int {
    void* v = nullptr;
    if (cudaMalloc(&v, bytes) != cudaSuccess) { cudaGetLastError(); set_last_error("prefill workspace: out of memory"); return CL_ERR_OOM; }
    allocs_.push_back(v);
    p = reinterpret_cast<std::remove_reference_t<decltype(p)>>(v);
    return CL_OK;
  };
  if ((rc = alloc(pws_->xn, T * d * 2))) return rc;
  if ((rc = alloc(pws_->qkv, T * qkv_dim_ * 4))) return rc;
  if ((rc = alloc(pws_->q, T * q_dim_ * 2))) return rc;
  if ((rc = alloc(pws_->attn, T * q_dim_ * 2))) return rc;
  if ((rc = alloc(pws_->h, T * d * 4))) return rc;
  if ((rc = alloc(pws_->gu, T * 2 * F * 4))) return rc;
  if ((rc = alloc(pws_->act, T * F * 2))) return rc;
  pws_->cap_tokens = (int)T;
  return CL_OK;
}

int Engine::prefill_chunked(cl_seq_t s, const int32_t* ids, int n, float* logits_out) {
  auto& q = seqs_[s];
  int rc = ensure_capacity(s, q.len + n);
  if (rc) return rc;
  const int d = cfg.d_model, F = cfg.d_ff;
  const int CH = std::min(n, prefill_chunk_tokens_);
  if ((rc = ensure_prefill_ws(CH))) return rc;
  PrefillWs& w = *pws_;
  const bool fused_silu = (prefill_fused_ & 1) != 0, fused_rope = (prefill_fused_ & 2) != 0;   // CL_PREFILL_FUSED bit mask
  CL_CUDA_OK(cudaMemcpyAsync(d_prompt_, ids, (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
  const int* bt = d_bt_ + (size_t)s * max_pages_per_seq_;
  int launches = 0, r;
  PrefillProfiler pp(stream_);
#define CL_LAUNCH(call) do { r = (call); if (r < 0) { set_last_error(std::string(#call) + ": " + cudaGetErrorString(cudaGetLastError())); return CL_ERR_CUDA; } launches += r; pp.mark(#call); } while (0)
  pp.mark("start");
  for (int c0 = 0; c0 < n; c0 += w.cap_tokens) {
    const int T = std::min(w.cap_tokens, n - c0);
    const int pos0 = q.len + c0;
    CL_LAUNCH(launch_embed_rows(embed_, d, d_prompt_ + c0, w.h, T, stream_));
    for (int l = 0; l < cfg.n_layers; ++l) {
      const auto& L = layers_[l];
      CL_LAUNCH(launch_rmsnorm_bf16(w.h, L.attn_norm, cfg.rms_eps, w.xn, T, d, stream_));
      if (fused_rope && cfg.head_dim == 128) {   // RoPE + bf16 + q / paged-cache scatter in the GEMM epilogue
        GemmEpi eq;
        eq.kind = 2; eq.rope = rope_; eq.pos0 = pos0; eq.q_out = w.q; eq.q_dim = q_dim_;
        eq.kpool = kpool_ + (size_t)l * kv_layer_elems_; eq.vpool = vpool_ + (size_t)l * kv_layer_elems_;
        eq.block_table = bt; eq.page_size = page_size_; eq.n_heads = cfg.n_heads; eq.n_kv = cfg.n_kv_heads;
        CL_LAUNCH(launch_gemm_bf16_epi(w.xn, L.wqkv, T, qkv_dim_, d, eq, stream_));
      } else {
        CL_LAUNCH(launch_gemm_bf16(w.xn, L.wqkv, w.qkv, nullptr, T, qkv_dim_, d, stream_));
        RopeScatterArgs ra{w.qkv, qkv_dim_, rope_, pos0, T, w.q, kpool_ + (size_t)l * kv_layer_elems_,
                           vpool_ + (size_t)l * kv_layer_elems_, bt, page_size_, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim};
        CL_LAUNCH(launch_rope_scatter(ra, stream_));
      }
      AttnPrefillArgs aa{w.q, kpool_ + (size_t)l * kv_layer_elems_, vpool_ + (size_t)l * kv_layer_elems_, bt, page_size_,
                         pos0, T, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, w.attn};
      if (have_kv_maps_ && attn_prefill_tc_supported(cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, page_size_, pos0, T))
        CL_LAUNCH(launch_attn_prefill_tc(aa, kmap_, vmap_, (long long)l * n_pages_ * cfg.n_kv_heads * page_size_, stream_));
      else
        CL_LAUNCH(launch_attn_prefill(aa, stream_));
      CL_LAUNCH(launch_gemm_bf16(w.attn, L.wo, w.h, w.h, T, d, q_dim_, stream_));
      CL_LAUNCH(launch_rmsnorm_bf16(w.h, L.ffn_norm, cfg.rms_eps, w.xn, T, d, stream_));
      if (fused_silu) {                     // SiLU(g) * u -> bf16 in the GEMM epilogue
        GemmEpi eg;
        eg.kind = 1; eg.act = w.act; eg.ld_act = F;
        CL_LAUNCH(launch_gemm_bf16_epi(w.xn, L.wgu, T, 2 * F, d, eg, stream_));
      } else {
        CL_LAUNCH(launch_gemm_bf16(w.xn, L.wgu, w.gu, nullptr, T, 2 * F, d, stream_));
        CL_LAUNCH(launch_silu_mul_bf16(w.gu, w.act, T, F, stream_));
      }
      CL_LAUNCH(launch_gemm_bf16(w.act, L.wdown, w.h, w.h, T, d, F, stream_));
    }
    if (c0 + T == n) {
      // last position -> slot residual row, then the shared decode tail (final norm + LM head + argmax)
      CL_CUDA_OK(cudaMemcpyAsync(d_h_ + (size_t)s * d, w.h + (size_t)(T - 1) * d, (size_t)d * 4, cudaMemcpyDeviceToDevice, stream_));
    }
  }
  const int last_pos = q.len + n - 1;
  CL_CUDA_OK(cudaMemcpyAsync(d_pos_ + s, &last_pos, 4, cudaMemcpyHostToDevice, stream_));
  rc = set_single_slot(s);
  if (rc) return rc;
  GemvArgs lm;
  lm.slots = d_slots_; lm.batch = 1;
  lm.W = lm_head_; lm.N = cfg.vocab_size; lm.K = d; lm.h = d_h_; lm.gain = final_norm_; lm.eps = cfg.rms_eps;
  lm.y = d_logits_; lm.x_stride = d; lm.y_stride = cfg.vocab_size;
  CL_LAUNCH(launch_gemv(gemv_variant_, EPI_STORE, true, lm, stream_, false));
  StepTailArgs t;
  t.logits = d_logits_; t.vocab = cfg.vocab_size; t.tok = d_tok_; t.pos = d_pos_; t.ids_ring = d_ids_ring_;
  t.step_counter = d_step_counter_; t.ring_steps = ring_steps_; t.ring_stride = max_batch_;
  t.part_val = d_tail_val_; t.part_idx = d_tail_idx_; t.counters = d_tail_cnt_; t.slots = d_slots_; t.batch = 1;
  CL_LAUNCH(launch_step_tail(t, stream_));
#undef CL_LAUNCH
  pp.report(n, "tile");
  launches_ += launches;
  q.len += n;
  q.history.insert(q.history.end(), ids, ids + n);
  if (logits_out) return read_logits(s, logits_out);
  CL_CUDA_OK(cudaStreamSynchronize(stream_));
  return CL_OK;
}


// Several prompts in ONE pass of the tile path (admission of a burst of short chats): the rows of all prompts are
// concatenated, so every projection streams its weights once for all of them and runs on full 128 x 256 tiles —
// 64 prompts of 130 tokens cost 64 x 6.0 ms one by one (the split-K short-prompt path streams all weights per prompt)
// against about 2.3 ms each in passes of 1024-4096 rows.  Everything row-wise (embedding, RMSNorm, the four GEMMs,
// SiLU) is unchanged; RoPE + cache scatter and the causal attention run per sequence on its row range (pointer
// offsets into the same workspace; each sequence has its own position offset and block table); the last row of every
// sequence goes through the batched LM head of the decode step (final norm + tcgen05 GEMM + argmax), which leaves
// d_tok_ / d_pos_ of every slot ready for the first decode step.  A row's results do not depend on the other rows of
// the pass: logits and cache contents are bit-identical to prefill_chunked on each prompt alone (tests/test_gpu_prefill.py).
int Engine::prefill_multi(int n_seqs, const cl_seq_t* ss, const int32_t* const* ids, const int* lens, float* logits_out) {
  if (n_seqs <= 0 || !ss || !ids || !lens) { set_last_error("prefill_multi: bad arguments"); return CL_ERR_INVALID_ARG; }
  if (!prefill_path_ok() || !bws_ || n_seqs > max_batch_) { set_last_error("prefill_multi: tensor-core paths unavailable or more sequences than max_batch"); return CL_ERR_INVALID_ARG; }
  const int d = cfg.d_model, F = cfg.d_ff, V = cfg.vocab_size;
  int total = 0, rc;
  for (int i = 0; i < n_seqs; ++i) {
    const int s = ss[i];
    if (s < 0 || s >= max_seqs_ || !seqs_[s].live) return CL_ERR_BAD_SEQ;
    for (int c = 0; c < i; ++c) if (ss[c] == s) { set_last_error("duplicate sequence in batch"); return CL_ERR_INVALID_ARG; }
    if (lens[i] <= 0 || !ids[i]) { set_last_error("empty prompt"); return CL_ERR_INVALID_ARG; }
    for (int t = 0; t < lens[i]; ++t)
      if (ids[i][t] < 0 || ids[i][t] >= V) { set_last_error("token id out of range"); return CL_ERR_INVALID_ARG; }
    if (seqs_[s].len + lens[i] > cfg.max_seq_len) { set_last_error("sequence exceeds max_seq_len"); return CL_ERR_TOO_LONG; }
    total += lens[i];
  }
  if (total > prefill_chunk_tokens_ || total > prompt_cap_) { set_last_error("prefill_multi: more rows than one pass holds"); return CL_ERR_TOO_LONG; }
  for (int i = 0; i < n_seqs; ++i)
    if ((rc = ensure_capacity(ss[i], seqs_[ss[i]].len + lens[i]))) return rc;
  if ((rc = ensure_prefill_ws(total))) return rc;
  PrefillWs& w = *pws_;
  BatchWs& bw = *bws_;
  const bool fused_silu = (prefill_fused_ & 1) != 0;
  std::vector<int32_t> all((size_t)total);
  std::vector<int> row0(n_seqs), pos0(n_seqs);
  for (int i = 0, r = 0; i < n_seqs; r += lens[i], ++i) {
    memcpy(all.data() + r, ids[i], (size_t)lens[i] * 4);
    row0[i] = r;
    pos0[i] = seqs_[ss[i]].len;
  }
  CL_CUDA_OK(cudaMemcpyAsync(d_prompt_, all.data(), (size_t)total * 4, cudaMemcpyHostToDevice, stream_));
  int launches = 0, r;
  PrefillProfiler pp(stream_);
#define CL_LAUNCH(call) do { r = (call); if (r < 0) { set_last_error(std::string(#call) + ": " + cudaGetErrorString(cudaGetLastError())); return CL_ERR_CUDA; } launches += r; pp.mark(#call); } while (0)
  pp.mark("start");
  const int T = total;
  CL_LAUNCH(launch_embed_rows(embed_, d, d_prompt_, w.h, T, stream_));
  for (int l = 0; l < cfg.n_layers; ++l) {
    const auto& L = layers_[l];
    __nv_bfloat16* kp = kpool_ + (size_t)l * kv_layer_elems_;
    __nv_bfloat16* vp = vpool_ + (size_t)l * kv_layer_elems_;
    CL_LAUNCH(launch_rmsnorm_bf16(w.h, L.attn_norm, cfg.rms_eps, w.xn, T, d, stream_));
    CL_LAUNCH(launch_gemm_bf16(w.xn, L.wqkv, w.qkv, nullptr, T, qkv_dim_, d, stream_));
    for (int i = 0; i < n_seqs; ++i) {                    // RoPE at the sequence's own positions, K/V into its own pages
      const int* bt = d_bt_ + (size_t)ss[i] * max_pages_per_seq_;
      RopeScatterArgs ra{w.qkv + (size_t)row0[i] * qkv_dim_, qkv_dim_, rope_, pos0[i], lens[i], w.q + (size_t)row0[i] * q_dim_, kp, vp, bt, page_size_,
                         cfg.n_heads, cfg.n_kv_heads, cfg.head_dim};
      CL_LAUNCH(launch_rope_scatter(ra, stream_));
    }
    for (int i = 0; i < n_seqs; ++i) {                    // causal attention inside each sequence
      const int* bt = d_bt_ + (size_t)ss[i] * max_pages_per_seq_;
      AttnPrefillArgs aa{w.q + (size_t)row0[i] * q_dim_, kp, vp, bt, page_size_, pos0[i], lens[i], cfg.n_heads, cfg.n_kv_heads, cfg.head_dim,
                         w.attn + (size_t)row0[i] * q_dim_};
      if (have_kv_maps_ && attn_prefill_tc_supported(cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, page_size_, pos0[i], lens[i]))
        CL_LAUNCH(launch_attn_prefill_tc(aa, kmap_, vmap_, (long long)l * n_pages_ * cfg.n_kv_heads * page_size_, stream_));
      else
        CL_LAUNCH(launch_attn_prefill(aa, stream_));
    }
    CL_LAUNCH(launch_gemm_bf16(w.attn, L.wo, w.h, w.h, T, d, q_dim_, stream_));
    CL_LAUNCH(launch_rmsnorm_bf16(w.h, L.ffn_norm, cfg.rms_eps, w.xn, T, d, stream_));
    if (fused_silu) {
      GemmEpi eg;
      eg.kind = 1; eg.act = w.act; eg.ld_act = F;
      CL_LAUNCH(launch_gemm_bf16_epi(w.xn, L.wgu, T, 2 * F, d, eg, stream_));
    } else {
      CL_LAUNCH(launch_gemm_bf16(w.xn, L.wgu, w.gu, nullptr, T, 2 * F, d, stream_));
      CL_LAUNCH(launch_silu_mul_bf16(w.gu, w.act, T, F, stream_));
    }
    CL_LAUNCH(launch_gemm_bf16(w.act, L.wdown, w.h, w.h, T, d, F, stream_));
  }
  // last row of every sequence -> its slot's residual row and position, then the decode step's batched head
  std::vector<int> slots(n_seqs);
  for (int i = 0; i < n_seqs; ++i) {
    const int s = ss[i], last_pos = pos0[i] + lens[i] - 1;
    slots[i] = s;
    CL_CUDA_OK(cudaMemcpyAsync(d_h_ + (size_t)s * d, w.h + (size_t)(row0[i] + lens[i] - 1) * d, (size_t)d * 4, cudaMemcpyDeviceToDevice, stream_));
    CL_CUDA_OK(cudaMemcpyAsync(d_pos_ + s, &last_pos, 4, cudaMemcpyHostToDevice, stream_));
  }
  CL_CUDA_OK(cudaMemcpyAsync(d_slots_, slots.data(), (size_t)n_seqs * 4, cudaMemcpyHostToDevice, stream_));
  slots_dirty_ = true;
  last_single_slot_ = n_seqs == 1 ? slots[0] : -1;
  CL_LAUNCH(launch_batch_resid_norm(d_h_, d, nullptr, 0, n_seqs, final_norm_, cfg.rms_eps, bw.xn, d_slots_, stream_, false));
  CL_LAUNCH(launch_gemm_bf16(bw.xn, lm_head_, bw.logits, nullptr, n_seqs, V, d, stream_, 1, false));
  CL_LAUNCH(launch_batch_scatter_rows(bw.logits, V, d_logits_, V, d_slots_, n_seqs, stream_));
  StepTailArgs t;
  t.logits = d_logits_; t.vocab = V; t.tok = d_tok_; t.pos = d_pos_; t.ids_ring = d_ids_ring_;
  t.step_counter = d_step_counter_; t.ring_steps = ring_steps_; t.ring_stride = max_batch_;
  t.part_val = d_tail_val_; t.part_idx = d_tail_idx_; t.counters = d_tail_cnt_; t.slots = d_slots_; t.batch = n_seqs;
  CL_LAUNCH(launch_step_tail(t, stream_));
#undef CL_LAUNCH
  pp.report(total, "multi-prompt tile");
  launches_ += launches;
  for (int i = 0; i < n_seqs; ++i) {
    auto& q = seqs_[ss[i]];
    q.len += lens[i];
    q.history.insert(q.history.end(), ids[i], ids[i] + lens[i]);
  }
  if (logits_out) {
    for (int i = 0; i < n_seqs; ++i)
      if ((rc = read_logits(ss[i], logits_out + (size_t)i * V))) return rc;
    return CL_OK;
  }
  CL_CUDA_OK(cudaStreamSynchronize(stream_));
  return CL_OK;
}

}  // namespace cl

// ================================================================================================
using namespace cl;

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <typename T> T* as() { return reinterpret_cast<T*>(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  cudaError_t upload(const void* h, size_t bytes) {
    cudaError_t e = alloc(bytes);
    return e != cudaSuccess ? e : cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
  }
};
int check_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    set_last_error("no CUDA device (libclengine has no CPU fallback)");
    return CL_ERR_NO_DEVICE;
  }
  CL_CUDA_OK(cudaSetDevice(device));
  return CL_OK;
}
}  // namespace

extern "C" {

int cl_op_gemm_bf16(int device, const uint16_t* x, const uint16_t* w, float* y, int32_t t, int32_t n, int32_t k, int32_t iters,
                    float* ms) {
  if (!x || !w || !y || t <= 0 || n <= 0 || k <= 0) return CL_ERR_INVALID_ARG;
  int rc = check_device(device);
  if (rc) return rc;
  if (!gemm_tcgen05_supported(t, n, k)) { set_last_error("shape not supported by the tcgen05 GEMM (K % 8)"); return CL_ERR_INVALID_ARG; }
  DevBuf dx, dw, dy;
  CL_CUDA_OK(dx.upload(x, (size_t)t * k * 2));
  CL_CUDA_OK(dw.upload(w, (size_t)n * k * 2));
  CL_CUDA_OK(dy.alloc((size_t)t * n * 4));
  CL_CUDA_OK(cudaMemset(dy.p, 0xff, (size_t)t * n * 4));  // NaN pattern: every element must be written
  if (launch_gemm_bf16(dx.as<__nv_bfloat16>(), dw.as<__nv_bfloat16>(), dy.as<float>(), nullptr, t, n, k, nullptr) < 0) {
    CL_CUDA_OK(cudaGetLastError());
    set_last_error("launch_gemm_bf16 failed");
    return CL_ERR_CUDA;
  }
  CL_CUDA_OK(cudaDeviceSynchronize());
  CL_CUDA_OK(cudaMemcpy(y, dy.p, (size_t)t * n * 4, cudaMemcpyDeviceToHost));
  if (iters > 0 && ms) {
    cudaEvent_t e0, e1;
    CL_CUDA_OK(cudaEventCreate(&e0));
    CL_CUDA_OK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch_gemm_bf16(dx.as<__nv_bfloat16>(), dw.as<__nv_bfloat16>(), dy.as<float>(), nullptr, t, n, k, nullptr);
    CL_CUDA_OK(cudaDeviceSynchronize());
    CL_CUDA_OK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch_gemm_bf16(dx.as<__nv_bfloat16>(), dw.as<__nv_bfloat16>(), dy.as<float>(), nullptr, t, n, k, nullptr);
    CL_CUDA_OK(cudaEventRecord(e1));
    CL_CUDA_OK(cudaDeviceSynchronize());
    float tm = 0.f;
    CL_CUDA_OK(cudaEventElapsedTime(&tm, e0, e1));
    *ms = tm / (float)iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  return CL_OK;
}

// variant: -1 auto (tcgen05 kernel when the shape allows, CL_PREFILL_ATTN_TC=0 forces the other), 0 mma.sync kernel, 1 tcgen05 kernel
static int attn_prefill_impl(int device, const uint16_t* q, const uint16_t* k, const uint16_t* v, int32_t t, int32_t n_heads, int32_t n_kv,
                             int32_t head_dim, int variant, int iters, float* out, float* ms) {
  if (!q || !k || !v || !out || t <= 0) return CL_ERR_INVALID_ARG;
  int rc = check_device(device);
  if (rc) return rc;
  if ((head_dim != 64 && head_dim != 128) || n_kv <= 0 || n_heads % n_kv) { set_last_error("unsupported attention shape"); return CL_ERR_INVALID_ARG; }
  const int P = 32, HD = head_dim;
  const int n_pages = (t + P - 1) / P;
  std::vector<int> bt(n_pages);
  for (int i = 0; i < n_pages; ++i) bt[i] = n_pages - 1 - i;  // reversed page order
  const size_t pool = (size_t)n_pages * n_kv * P * HD;
  std::vector<uint16_t> kp(pool, 0), vp(pool, 0);
  for (int tt = 0; tt < t; ++tt)
    for (int g = 0; g < n_kv; ++g) {
      const size_t dst = (((size_t)bt[tt / P] * n_kv + g) * P + tt % P) * HD;
      memcpy(&kp[dst], k + ((size_t)tt * n_kv + g) * HD, (size_t)HD * 2);
      memcpy(&vp[dst], v + ((size_t)tt * n_kv + g) * HD, (size_t)HD * 2);
    }
  const size_t qd = (size_t)n_heads * HD;
  DevBuf dq, dk, dv, dbt, dout, dout32;
  CL_CUDA_OK(dq.upload(q, (size_t)t * qd * 2));
  CL_CUDA_OK(dk.upload(kp.data(), pool * 2));
  CL_CUDA_OK(dv.upload(vp.data(), pool * 2));
  CL_CUDA_OK(dbt.upload(bt.data(), bt.size() * 4));
  CL_CUDA_OK(dout.alloc((size_t)t * qd * 2));
  CL_CUDA_OK(cudaMemset(dout.p, 0xff, (size_t)t * qd * 2));   // NaN pattern: every output must be written
  CL_CUDA_OK(dout32.alloc((size_t)t * qd * 4));
  AttnPrefillArgs a{dq.as<__nv_bfloat16>(), dk.as<__nv_bfloat16>(), dv.as<__nv_bfloat16>(), dbt.as<int>(), P, 0, t, n_heads, n_kv, HD,
                    dout.as<__nv_bfloat16>()};
  CUtensorMap km, vm;
  bool tc = variant != 0 && attn_prefill_tc_supported(n_heads, n_kv, HD, P, 0, t);
  if (variant == 1 && !tc && head_dim == 128) tc = (t + P - 1) / P <= 320;          // explicit request overrides the env switch
  if (tc) tc = make_tmap_2d_bf16(&km, dk.p, (uint64_t)n_pages * n_kv * P, HD, 64, 32) && make_tmap_2d_bf16(&vm, dv.p, (uint64_t)n_pages * n_kv * P, HD, 64, 32);
  if (variant == 1 && !tc) { set_last_error("tcgen05 prefill attention does not support this shape"); return CL_ERR_INVALID_ARG; }
  auto run = [&]() { return tc ? launch_attn_prefill_tc(a, km, vm, 0, nullptr) : launch_attn_prefill(a, nullptr); };
  if (run() < 0) { CL_CUDA_OK(cudaGetLastError()); set_last_error("prefill attention launch failed"); return CL_ERR_CUDA; }
  if (launch_bf16_to_f32(dout.as<__nv_bfloat16>(), dout32.as<float>(), (int64_t)t * qd, nullptr) < 0) return CL_ERR_CUDA;
  CL_CUDA_OK(cudaDeviceSynchronize());
  CL_CUDA_OK(cudaMemcpy(out, dout32.p, (size_t)t * qd * 4, cudaMemcpyDeviceToHost));
  if (iters > 0 && ms) {
    cudaEvent_t e0, e1;
    CL_CUDA_OK(cudaEventCreate(&e0));
    CL_CUDA_OK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) run();
    CL_CUDA_OK(cudaDeviceSynchronize());
    CL_CUDA_OK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) run();
    CL_CUDA_OK(cudaEventRecord(e1));
    CL_CUDA_OK(cudaDeviceSynchronize());
    float tm = 0.f;
    CL_CUDA_OK(cudaEventElapsedTime(&tm, e0, e1));
    *ms = tm / (float)iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  return CL_OK;
}

int cl_op_attn_prefill(int device, const uint16_t* q, const uint16_t* k, const uint16_t* v, int32_t t, int32_t n_heads, int32_t n_kv,
                       int32_t head_dim, float* out) {
  return attn_prefill_impl(device, q, k, v, t, n_heads, n_kv, head_dim, -1, 0, out, nullptr);
}
int cl_op_attn_prefill_variant(int device, int variant, const uint16_t* q, const uint16_t* k, const uint16_t* v, int32_t t, int32_t n_heads,
                               int32_t n_kv, int32_t head_dim, float* out, int32_t iters, float* ms) {
  return attn_prefill_impl(device, q, k, v, t, n_heads, n_kv, head_dim, variant, iters, out, ms);
}

}  // extern "C"
