// C ABI of the fused FRI commit phase (kernels in fri.cu).
#include "capi_internal.cuh"

extern "C" {

// ---------------------------------------------------------------- the transcript as a device object (layer-by-layer callers)
int32_t sc_dchan_create(sc_ctx* ctx, const uint32_t digest[8], uint32_t max_mixes, sc_dchan** out) {
  ENTER();
  if (!digest || !out || !max_mixes || max_mixes > 4096) return fail(SC_EINVAL, "dchan_create: bad argument");
  sc_dchan* dc = new sc_dchan;
  dc->max = max_mixes;
  int32_t r = new_col(ctx, 8 + 12 * (uint64_t)max_mixes, &dc->buf);
  if (r) { delete dc; return r; }
  void* st = nullptr;
  r = stage(ctx, digest, 32, &st);
  if (r) { sc_col_free(ctx, dc->buf); delete dc; return r; }
  CK(cudaMemcpyAsync(dc->buf->d, st, 32, cudaMemcpyDeviceToDevice, ctx->st));
  *out = dc;
  return SC_OK;
}
const uint32_t* sc_dchan_coeff_ptr(const sc_dchan* dc, uint32_t k) { return dc && k < dc->n ? dc->buf->d + 8 + 4 * (size_t)k : nullptr; }
// Blake2sMerkleChannel::mix_root(root_col[0..8)) then Blake2sChannel::draw_felt -> coefficient #n (n = mixes so far), on the stream
int32_t sc_dchan_mix_root_draw(sc_ctx* ctx, sc_dchan* dc, const sc_col* root_col) {
  ENTER();
  if (!dc || !root_col || root_col->len < 8) return fail(SC_EINVAL, "dchan_mix_root_draw: bad argument");
  if (dc->n >= dc->max) return fail(SC_EINVAL, "dchan_mix_root_draw: more mixes than the channel was created for");
  uint32_t* base = dc->buf->d;
  { ProfScope ps(ctx, "fri_channel");
    CKL(launch_fri_channel(base, root_col->d, base + 8 + 4 * (size_t)dc->n, base + 8 + 4 * (size_t)dc->max + 8 * (size_t)dc->n, ctx->st)); }
  dc->n++;
  return SC_OK;
}
// Waits for the stream and returns the roots mixed so far (8 words each); frees the channel.
int32_t sc_dchan_finish(sc_ctx* ctx, sc_dchan* dc, uint32_t* roots_out) {
  if (!dc) return SC_OK;
  ENTER();
  int32_t r = SC_OK;
  if (roots_out && dc->n) r = sc_col_read(ctx, dc->buf, 8 + 4 * (uint64_t)dc->max, 8 * (uint64_t)dc->n, roots_out);
  sc_col_free(ctx, dc->buf);
  delete dc;
  return r;
}

// The FRI layers of <= 2^FRI_TAIL_LOG values on a device transcript, in ONE launch (fri_tail_kernel): for lg = start_log ..
// last_log + 1: circle-fold quot_cols[4 * t ..] (the quotient column of log lg + 1, NULL entries when there is none) into the
// line evaluation with coefficient #0, commit the evaluation (evals_out, layers_out: lg + 1 layers per tree, leaves first
// at index lg ... root at 0 within each tree's block), mix the root, draw, fold_line.  last_out: the 2^last_log values left.
// The layer-by-layer driver of the sharded prover calls this once its line evaluation is replicated and small: ten launches
// per layer (fold, tree, top, channel, fold) become one for the whole tail.  dc advances by start_log - last_log mixes.
int32_t sc_dchan_fri_tail(sc_ctx* ctx, sc_dchan* dc, const sc_twiddles* tw, sc_col* const layer_in[4], uint32_t start_log, uint32_t last_log,
                          sc_col* const* quot_cols, sc_col** evals_out, sc_col** layers_out, sc_col* last_out[4]) {
  ENTER();
  if (!dc || !tw || !layer_in || !quot_cols || !evals_out || !layers_out || !last_out) return fail(SC_EINVAL, "dchan_fri_tail: null argument");
  if (start_log > FRI_TAIL_LOG || start_log <= last_log || start_log > tw->root_log) return fail(SC_EINVAL, "dchan_fri_tail: bad log sizes");
  const uint32_t n_tail = start_log - last_log;
  if (!dc->n || dc->n + n_tail > dc->max) return fail(SC_EINVAL, "dchan_fri_tail: channel too short (or the first layer was not mixed)");
  for (int k = 0; k < 4; k++)
    if (!layer_in[k] || layer_in[k]->len != (1ull << start_log)) return fail(SC_EINVAL, "dchan_fri_tail: bad input layer");
  const uint32_t* itw_end = tw->itw + ((size_t)1 << tw->root_log);
  std::vector<sc_col*> made;
  auto undo = [&](int32_t r) { for (sc_col* c : made) sc_col_free(ctx, c); return r; };
  std::vector<uint32_t*> evp, trp;
  std::vector<const uint32_t*> qp;
  size_t loff = 0;
  for (uint32_t lg = start_log, t = 0; lg > last_log; lg--, t++) {
    for (int k = 0; k < 4; k++) {
      sc_col* c = nullptr; int32_t r = new_col(ctx, 1ull << lg, &c); if (r) return undo(r);
      made.push_back(c); evals_out[4 * t + k] = c; evp.push_back(c->d);
    }
    for (int kk = (int)lg; kk >= 0; kk--) {
      sc_col* c = nullptr; int32_t r = new_col(ctx, 8ull << kk, &c); if (r) return undo(r);
      made.push_back(c); layers_out[loff + kk] = c;
    }
    for (int kk = (int)lg; kk >= 0; kk--) trp.push_back(layers_out[loff + kk]->d);
    loff += lg + 1;
    for (int k = 0; k < 4; k++) {
      const sc_col* q = quot_cols[4 * t + k];
      if (q && q->len != (2ull << lg)) return undo(fail(SC_EINVAL, "dchan_fri_tail: bad quotient column"));
      if ((q == nullptr) != (quot_cols[4 * t] == nullptr)) return undo(fail(SC_EINVAL, "dchan_fri_tail: quotient coordinates must come in fours"));
      qp.push_back(q ? q->d : nullptr);
    }
  }
  for (int k = 0; k < 4; k++) { sc_col* c = nullptr; int32_t r = new_col(ctx, 1ull << last_log, &c); if (r) return undo(r); made.push_back(c); last_out[k] = c; }
  void *d_evp, *d_trp, *d_qp;
  { int32_t r = stage(ctx, evp.data(), evp.size() * sizeof(void*), &d_evp); if (r) return undo(r); }
  { int32_t r = stage(ctx, trp.data(), trp.size() * sizeof(void*), &d_trp); if (r) return undo(r); }
  { int32_t r = stage(ctx, qp.data(), qp.size() * sizeof(void*), &d_qp); if (r) return undo(r); }
  uint32_t* base = dc->buf->d;
  FriTailArgs a;
  a.start_log = start_log; a.last_log = last_log; a.one = 1u;
  for (int k = 0; k < 4; k++) { a.layer_in[k] = layer_in[k]->d; a.last_out[k] = last_out[k]->d; }
  a.itw_end = itw_end; a.digest = base; a.circle_alpha = base + 8;
  a.eval_out = (uint32_t* const*)d_evp; a.tree_out = (uint32_t* const*)d_trp; a.quot = (const uint32_t* const*)d_qp;
  a.roots_out = base + 8 + 4 * (size_t)dc->max + 8 * (size_t)dc->n;
  { ProfScope ps(ctx, "fri_tail"); int e = launch_fri_tail(a, ctx->st); if (e) { undo(0); CKL(e); } }
  dc->n += n_tail;
  return SC_OK;
}

// FriProver::commit (stwo-prover 0.1.1 @ 31e8dbc core/fri.rs; reached from prover::prove, crates/brainfuck_prover/src/
// brainfuck_air/mod.rs:732) with the transcript kept on the device — see fri.cu.
int32_t sc_fri_commit(sc_ctx* ctx, const sc_twiddles* tw, sc_col* const* quot_cols, const uint32_t* quot_logs, uint32_t nq,
                      const uint32_t channel_digest[8], uint32_t last_log, sc_col** first_layers_out, sc_col** inner_evals_out,
                      sc_col** inner_layers_out, uint32_t* roots_out, uint32_t* last_values_out) {
  ENTER();
  if (!tw || !quot_cols || !quot_logs || !nq || !channel_digest || !first_layers_out || !inner_evals_out || !inner_layers_out || !roots_out ||
      !last_values_out)
    return fail(SC_EINVAL, "fri_commit: null argument");
  const uint32_t top = quot_logs[0];
  if (top < 3 || top > tw->root_log + 1 || last_log + 1 >= top || last_log > FRI_TAIL_LOG)
    return fail(SC_EINVAL, "fri_commit: bad log sizes / twiddle tree too small");
  for (uint32_t q = 0; q < nq; q++) {
    if (q && quot_logs[q] >= quot_logs[q - 1]) return fail(SC_EINVAL, "fri_commit: columns must be strictly descending in size");
    if (quot_logs[q] < 3) return fail(SC_EINVAL, "fri_commit: column too small");
    for (int k = 0; k < 4; k++)
      if (!quot_cols[4 * q + k] || quot_cols[4 * q + k]->len != (1ull << quot_logs[q])) return fail(SC_EINVAL, "fri_commit: bad quotient column");
  }
  const uint32_t n_inner = top - 1 - last_log;
  const uint32_t* itw_end = tw->itw + ((size_t)1 << tw->root_log);
  // device scratch: digest | one coefficient per layer (first layer + inner) | roots | last layer values
  const size_t n_alpha = 4 * (size_t)(n_inner + 1), n_roots = 8 * (size_t)(n_inner + 1), n_last = (size_t)4 << last_log;
  sc_col* buf = nullptr;
  { int32_t r = new_col(ctx, 8 + n_alpha + n_roots + n_last, &buf); if (r) return r; }
  uint32_t *d_digest = buf->d, *d_alpha = d_digest + 8, *d_roots = d_alpha + n_alpha, *d_last = d_roots + n_roots;
  std::vector<sc_col*> temps{buf};
  auto done = [&](int32_t r) { for (sc_col* c : temps) sc_col_free(ctx, c); return r; };
  { void* st = nullptr; int32_t r = stage(ctx, channel_digest, 32, &st); if (r) return done(r);
    cudaError_t e = cudaMemcpyAsync(d_digest, st, 32, cudaMemcpyDeviceToDevice, ctx->st);
    if (e != cudaSuccess) { done(0); CK(e); } }
  // ---- first layer: one tree over every quotient coordinate column
  { int32_t r = merkle_commit_impl(ctx, quot_cols, 4 * nq, 0, first_layers_out, nullptr, nullptr); if (r) return done(r); }
  { ProfScope ps(ctx, "fri_channel");
    int e = launch_fri_channel(d_digest, first_layers_out[0]->d, d_alpha, d_roots, ctx->st); if (e) { done(0); CKL(e); } }
  const uint32_t* d_circle_alpha = d_alpha;
  // ---- inner layers above the tail: fold, tree, channel, fold
  sc_col* layer[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t line_log = top - 1, qi = 0, li = 0;
  size_t loff = 0;
  const uint32_t tail_from = std::max(FRI_TAIL_LOG, last_log);
  while (line_log > tail_from) {
    bool fresh = false;
    if (!layer[0]) {
      for (int k = 0; k < 4; k++) { int32_t r = new_col(ctx, 1ull << line_log, &layer[k]); if (r) return done(r); }
      fresh = true;
    }
    while (qi < nq && quot_logs[qi] == line_log + 1) {
      const uint32_t* s[4]; uint32_t* d[4];
      for (int k = 0; k < 4; k++) { s[k] = quot_cols[4 * qi + k]->d; d[k] = layer[k]->d; }
      ProfScope ps(ctx, "fold_circle_into_line");
      int e = launch_fold_circle_dev(s, line_log + 1, d_circle_alpha, d, itw_end, fresh, ctx->st); if (e) { done(0); CKL(e); }
      fresh = false;
      qi++;
    }
    if (fresh) return done(fail(SC_EINVAL, "fri_commit: no column for the first line layer"));
    for (int k = 0; k < 4; k++) inner_evals_out[4 * li + k] = layer[k];
    { int32_t r = merkle_commit_impl(ctx, layer, 4, 0, inner_layers_out + loff, nullptr, nullptr); if (r) return done(r); }
    { ProfScope ps(ctx, "fri_channel");
      int e = launch_fri_channel(d_digest, inner_layers_out[loff]->d, d_alpha + 4 * (li + 1), d_roots + 8 * (li + 1), ctx->st);
      if (e) { done(0); CKL(e); } }
    sc_col* next[4];
    for (int k = 0; k < 4; k++) { int32_t r = new_col(ctx, 1ull << (line_log - 1), &next[k]); if (r) return done(r); }
    { const uint32_t* s[4]; uint32_t* d[4];
      for (int k = 0; k < 4; k++) { s[k] = layer[k]->d; d[k] = next[k]->d; }
      ProfScope ps(ctx, "fold_line");
      int e = launch_fold_line_dev(s, line_log, d_alpha + 4 * (li + 1), d, itw_end, ctx->st); if (e) { done(0); CKL(e); } }
    for (int k = 0; k < 4; k++) layer[k] = next[k];   // the previous evaluation now belongs to the caller (inner_evals_out)
    loff += line_log + 1;
    line_log--; li++;
  }
  // ---- the tail: one CTA
  if (line_log > last_log) {
    const uint32_t start = line_log, n_tail = start - last_log;
    std::vector<uint32_t*> evp, trp;
    std::vector<const uint32_t*> qp;
    for (uint32_t lg = start, t = 0; lg > last_log; lg--, t++) {
      for (int k = 0; k < 4; k++) {
        sc_col* c = nullptr; int32_t r = new_col(ctx, 1ull << lg, &c); if (r) return done(r);
        inner_evals_out[4 * (li + t) + k] = c; evp.push_back(c->d);
      }
      for (int kk = (int)lg; kk >= 0; kk--) {
        sc_col* c = nullptr; int32_t r = new_col(ctx, 8ull << kk, &c); if (r) return done(r);
        inner_layers_out[loff + kk] = c;
      }
      for (int kk = (int)lg; kk >= 0; kk--) trp.push_back(inner_layers_out[loff + kk]->d);
      loff += lg + 1;
      if (qi < nq && quot_logs[qi] == lg + 1) { for (int k = 0; k < 4; k++) qp.push_back(quot_cols[4 * qi + k]->d); qi++; }
      else for (int k = 0; k < 4; k++) qp.push_back(nullptr);
    }
    void *d_evp, *d_trp, *d_qp;
    { int32_t r = stage(ctx, evp.data(), evp.size() * sizeof(void*), &d_evp); if (r) return done(r); }
    { int32_t r = stage(ctx, trp.data(), trp.size() * sizeof(void*), &d_trp); if (r) return done(r); }
    { int32_t r = stage(ctx, qp.data(), qp.size() * sizeof(void*), &d_qp); if (r) return done(r); }
    FriTailArgs a;
    a.start_log = start; a.last_log = last_log; a.one = 1u;
    for (int k = 0; k < 4; k++) { a.layer_in[k] = layer[0] ? layer[k]->d : nullptr; a.last_out[k] = d_last + ((size_t)k << last_log); }
    a.itw_end = itw_end; a.digest = d_digest; a.circle_alpha = d_circle_alpha;
    a.eval_out = (uint32_t* const*)d_evp; a.tree_out = (uint32_t* const*)d_trp; a.quot = (const uint32_t* const*)d_qp;
    a.roots_out = d_roots + 8 * (li + 1);
    { ProfScope ps(ctx, "fri_tail"); int e = launch_fri_tail(a, ctx->st); if (e) { done(0); CKL(e); } }
    for (int k = 0; k < 4; k++) if (layer[k]) temps.push_back(layer[k]);   // the tail's input was a temporary
    li += n_tail;
  } else {
    for (int k = 0; k < 4; k++) {
      cudaError_t e = cudaMemcpyAsync(d_last + ((size_t)k << last_log), layer[k]->d, (size_t)4 << last_log, cudaMemcpyDeviceToDevice, ctx->st);
      if (e != cudaSuccess) { done(0); CK(e); }
      temps.push_back(layer[k]);
    }
  }
  if (qi != nq) return done(fail(SC_EINVAL, "fri_commit: not all columns consumed"));
  // ---- the one read-back of the phase: every root and the last layer
  {
    std::vector<uint32_t> host(n_roots + n_last);
    cudaError_t e = cudaMemcpyAsync(host.data(), d_roots, host.size() * 4, cudaMemcpyDeviceToHost, ctx->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
    if (e != cudaSuccess) { done(0); CK(e); }
    memcpy(roots_out, host.data(), n_roots * 4);
    memcpy(last_values_out, host.data() + n_roots, n_last * 4);
  }
  return done(SC_OK);
}

}  // extern "C"
