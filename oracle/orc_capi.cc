// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_field.h header).
// C entry points for ctypes (tests/, smoke(), bench.py cpu_baseline / --impl reference).
#include "orc_blake2s.h"
#include "orc_circle.h"
#include "orc_ops.h"
#include <map>
#include <mutex>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {
struct Tree { std::vector<uint32_t> tw, itw; };
std::map<uint32_t, Tree> g_trees;
std::mutex g_mu;
// Tree rooted at CanonicCoset(root_log+1).circle_domain().half_coset (log root_log), cf. brainfuck_air/mod.rs:480-484.
const Tree& tree_for(uint32_t root_log) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_trees.find(root_log);
  if (it != g_trees.end()) return it->second;
  Tree& t = g_trees[root_log];
  precompute_twiddles(coset_half_odds(root_log), t.tw, t.itw);
  return t;
}
}  // namespace

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);   // a launcher may have exported OMP_NUM_THREADS=1 (torchrun does)
#else
  (void)n;
#endif
}

void orc_precompute_twiddles(uint32_t root_log, uint32_t* tw, uint32_t* itw) {
  const Tree& t = tree_for(root_log);
  memcpy(tw, t.tw.data(), t.tw.size() * 4);
  memcpy(itw, t.itw.data(), t.itw.size() * 4);
}

// In-place interpolate of ncols columns (each 2^log) stored back to back.
void orc_interpolate(uint32_t* v, uint32_t log, uint32_t ncols, uint32_t root_log) {
  const Tree& t = tree_for(root_log);
#pragma omp parallel for schedule(dynamic)
  for (uint32_t c = 0; c < ncols; c++) interpolate(v + ((size_t)c << log), log, t.itw);
}

// coeffs (2^log each) -> evaluations on canonic_domain(log + log_blowup) (2^(log+blowup) each).
void orc_evaluate(const uint32_t* coeffs, uint32_t log, uint32_t log_blowup, uint32_t ncols, uint32_t root_log, uint32_t* out) {
  const Tree& t = tree_for(root_log);
  uint32_t elog = log + log_blowup;
#pragma omp parallel for schedule(dynamic)
  for (uint32_t c = 0; c < ncols; c++) {
    uint32_t* o = out + ((size_t)c << elog);
    memcpy(o, coeffs + ((size_t)c << log), (size_t)4 << log);
    memset(o + ((size_t)1 << log), 0, ((size_t)4 << elog) - ((size_t)4 << log));  // extend(): zero-pad at the end
    evaluate(o, elog, t.tw);
  }
}

void orc_eval_at_point(const uint32_t* coeffs, uint32_t log, const uint32_t pxy[8], uint32_t out[4]) {
  QPt p = {qfrom(pxy[0], pxy[1], pxy[2], pxy[3]), qfrom(pxy[4], pxy[5], pxy[6], pxy[7])};
  QM31 r = eval_at_point(coeffs, log, p);
  out[0] = r.a.a; out[1] = r.a.b; out[2] = r.b.a; out[3] = r.b.b;
}

// domain point i (natural circle-domain order) of canonic_domain(log)
void orc_domain_at(uint32_t log, uint32_t i, uint32_t out[2]) {
  Pt p = canonic_domain(log).at(i);
  out[0] = p.x; out[1] = p.y;
}

void orc_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1, uint32_t f0, uint32_t f1) {
  b2s_compress(h, m, t0, t1, f0, f1);
}
void orc_blake2s256(const uint8_t* data, size_t len, uint8_t out[32]) { blake2s256(data, len, out); }

void orc_commit_on_layer(uint32_t log, const uint32_t* prev, const uint32_t* const* cols, uint32_t ncols, uint32_t* out) {
  commit_on_layer(log, prev, cols, ncols, out);
}

// Full mixed-size tree.  out_layers: concatenation of layers max_log, max_log-1, …, 0 (8 words per node).
void orc_merkle_commit(const uint32_t* const* cols, const uint32_t* logs, uint32_t ncols, uint32_t* out_layers) {
  std::vector<std::vector<uint32_t>> layers;
  merkle_commit(cols, logs, ncols, layers);
  size_t off = 0;
  for (int lg = (int)layers.size() - 1; lg >= 0; lg--) {
    memcpy(out_layers + off, layers[lg].data(), layers[lg].size() * 4);
    off += layers[lg].size();
  }
}

void orc_bit_reverse(uint32_t* v, uint32_t log) { bit_reverse_column(v, log); }
void orc_batch_inverse_m31(const uint32_t* src, uint32_t* dst, size_t n) {
  for (size_t i = 0; i < n; i++) dst[i] = minv(src[i]);
}
// QM31 columns by coordinates: src[4][n]
void orc_batch_inverse_qm31(const uint32_t* const* src, uint32_t* const* dst, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    QM31 r = qinv(qfrom(src[0][i], src[1][i], src[2][i], src[3][i]));
    dst[0][i] = r.a.a; dst[1][i] = r.a.b; dst[2][i] = r.b.a; dst[3][i] = r.b.b;
  }
}

void orc_fold_line(const uint32_t* const* src, uint32_t log, const uint32_t alpha[4], uint32_t* const* dst) {
  fold_line(src, log, qfrom(alpha[0], alpha[1], alpha[2], alpha[3]), dst);
}
void orc_fold_circle_into_line(const uint32_t* const* src, uint32_t log, const uint32_t alpha[4], uint32_t* const* dst) {
  fold_circle_into_line(src, log, qfrom(alpha[0], alpha[1], alpha[2], alpha[3]), dst);
}
void orc_accumulate(uint32_t* const* dst, const uint32_t* const* src, size_t n) {
  for (int k = 0; k < 4; k++)
    for (size_t i = 0; i < n; i++) dst[k][i] = madd(dst[k][i], src[k][i]);
}
void orc_secure_powers(const uint32_t felt[4], uint32_t n, uint32_t* out /* n*4 */) {
  QM31 f = qfrom(felt[0], felt[1], felt[2], felt[3]), acc = qfromm(1);
  for (uint32_t i = 0; i < n; i++) {
    out[4 * i] = acc.a.a; out[4 * i + 1] = acc.a.b; out[4 * i + 2] = acc.b.a; out[4 * i + 3] = acc.b.b;
    acc = qmul(acc, f);
  }
}
void orc_gen_is_first(uint32_t log, uint32_t* out) {
  memset(out, 0, (size_t)4 << log);
  out[0] = 1;
}
void orc_prefix_sum_bitrev(uint32_t* v, uint32_t log) { prefix_sum_bitrev(v, log); }

// accumulate_quotients: cols[ncols][2^log] on canonic_domain(log); batches flattened:
//   batch b: point (8 words), n_b entries of (col index, value[4]).
void orc_accumulate_quotients(uint32_t log, const uint32_t* const* cols, uint32_t ncols, const uint32_t alpha[4],
                              const uint32_t* batch_points /*nb*8*/, const uint32_t* batch_sizes /*nb*/,
                              const uint32_t* entry_cols, const uint32_t* entry_vals /*ne*4*/, uint32_t nb,
                              uint32_t* const* out) {
  accumulate_quotients(log, cols, ncols, qfrom(alpha[0], alpha[1], alpha[2], alpha[3]), batch_points, batch_sizes,
                       entry_cols, entry_vals, nb, out);
}

uint64_t orc_grind(const uint32_t digest[8], uint32_t pow_bits) { return grind(digest, pow_bits); }

}  // extern "C"
