// Proof objects and shared protocol helpers (queries, sample batching, circle points over QM31, config).
// Restates stwo-prover 0.1.1 @ 31e8dbc core/{pcs/mod.rs,pcs/quotients.rs,queries.rs,fri.rs,vcs/prover.rs} data shapes
//  and the reference's BrainfuckProof (crates/brainfuck_prover/src/brainfuck_air/mod.rs:71-76).
#pragma once
#include <algorithm>
#include <map>
#include <set>
#include <charconv>
#include <cstring>
#include <sstream>
#include <string>
#include "backend.hpp"

namespace sbf {

struct ProverConfig {          // PcsConfig::default() + LOG_MAX_ROWS (brainfuck_air/mod.rs:427-433,479)
  uint32_t log_max_rows = 24;  // 20 under cfg(test) in the reference
  uint32_t pow_bits = 5;
  uint32_t log_blowup = 1;
  uint32_t n_queries = 3;
  uint32_t log_last_layer_degree_bound = 0;
  uint32_t shard_min_log = 0;  // sharded driver: columns below 2^shard_min_log rows are replicated on every rank (0: shard whatever can be)
  bool overlap_host = true;    // build the tables on the host while the device commits the preprocessed tree
};

struct MerkleDecommitment {
  std::vector<Hash> hash_witness;
  std::vector<uint32_t> column_witness;
};
struct FriLayerProof {
  std::vector<QM31> fri_witness;
  MerkleDecommitment decommitment;
  Hash commitment;
};
struct FriProof {
  FriLayerProof first_layer;
  std::vector<FriLayerProof> inner_layers;
  std::vector<QM31> last_layer_poly;
};
struct CommitmentSchemeProof {
  std::vector<Hash> commitments;
  std::vector<std::vector<std::vector<QM31>>> sampled_values;      // tree -> column -> sample
  std::vector<MerkleDecommitment> decommitments;                   // per tree
  std::vector<std::vector<std::vector<uint32_t>>> queried_values;  // tree -> column -> query
  uint64_t proof_of_work = 0;
  FriProof fri_proof;
};
struct BrainfuckProof {
  uint32_t log_size[N_COMPONENTS];   // BrainfuckClaim
  QM31 claimed_sum[N_COMPONENTS];    // BrainfuckInteractionClaim
  CommitmentSchemeProof proof;       // StarkProof
};

// ---- circle points over QM31
inline QPoint qp_add(QPoint p, QPoint q) {
  return {q_sub(q_mul(p.x, q.x), q_mul(p.y, q.y)), q_add(q_mul(p.x, q.y), q_mul(p.y, q.x))};
}
inline Pt point_at_index(uint32_t idx) {
  Pt r{1u, 0u}, b{GEN_X, GEN_Y};
  idx &= 0x7fffffffu;
  while (idx) { if (idx & 1u) r = p_add(r, b); b = p_dbl(b); idx >>= 1; }
  return r;
}
inline QPoint to_qpoint(Pt p) { return {q_fromm(p.x), q_fromm(p.y)}; }
// CanonicCoset(log).circle_domain().at(i)
inline Pt canonic_domain_at(uint32_t log, uint32_t i) {
  uint32_t half = 1u << (log - 1);
  uint32_t init = 1u << (30 - log), step = 1u << (32 - log);
  if (i < half) return point_at_index(init + (uint32_t)(((uint64_t)step * i) & 0x7fffffffu));
  Pt p = point_at_index(init + (uint32_t)(((uint64_t)step * (i - half)) & 0x7fffffffu));
  return p_conj(p);
}
// CirclePoint::get_random_point
inline QPoint random_point(Channel& ch) {
  QM31 t = ch.draw_felt();
  QM31 t2 = q_sqr(t);
  QM31 inv = q_inv(q_add(t2, q_fromm(1)));
  return {q_mul(q_sub(q_fromm(1), t2), inv), q_mul(q_add(t, t), inv)};
}
inline bool qm31_less(const QM31& a, const QM31& b) {
  uint32_t x[4] = {a.a.a, a.a.b, a.b.a, a.b.b}, y[4] = {b.a.a, b.a.b, b.b.a, b.b.b};
  for (int i = 0; i < 4; i++) if (x[i] != y[i]) return x[i] < y[i];
  return false;
}
struct QPointLess {
  bool operator()(const QPoint& p, const QPoint& q) const {
    if (!q_eq(p.x, q.x)) return qm31_less(p.x, q.x);
    return qm31_less(p.y, q.y);
  }
};
struct PointSample { QPoint point; QM31 value; };

// ColumnSampleBatch::new_vec for the columns of one size group (BTreeMap keyed by point, stable inside a point)
inline SampleBatchesFlat batch_samples(const std::vector<const std::vector<PointSample>*>& cols) {
  std::map<QPoint, std::vector<std::pair<uint32_t, QM31>>, QPointLess> grouped;
  for (uint32_t c = 0; c < cols.size(); c++)
    for (const auto& s : *cols[c]) grouped[s.point].push_back({c, s.value});
  SampleBatchesFlat f;
  for (auto& kv : grouped) {
    const QPoint& p = kv.first;
    uint32_t w[8] = {p.x.a.a, p.x.a.b, p.x.b.a, p.x.b.b, p.y.a.a, p.y.a.b, p.y.b.a, p.y.b.b};
    f.points.insert(f.points.end(), w, w + 8);
    f.sizes.push_back((uint32_t)kv.second.size());
    for (auto& e : kv.second) {
      f.entry_cols.push_back(e.first);
      uint32_t v[4] = {e.second.a.a, e.second.a.b, e.second.b.a, e.second.b.b};
      f.entry_vals.insert(f.entry_vals.end(), v, v + 4);
    }
  }
  return f;
}

// accumulate_row_quotients (core/pcs/quotients.rs) — verifier side, one row.
inline QM31 row_quotient(const SampleBatchesFlat& f, const std::vector<uint32_t>& row_vals, QM31 alpha, Pt dp) {
  QM31 acc = q_zero();
  size_t e = 0;
  for (size_t b = 0; b < f.sizes.size(); b++) {
    const uint32_t* q = &f.points[8 * b];
    QM31 sx = q_make(q[0], q[1], q[2], q[3]), sy = q_make(q[4], q[5], q[6], q[7]);
    CM31 den = c_sub(c_mul(c_sub(sx.a, CM31{dp.x, 0}), sy.b), c_mul(c_sub(sy.a, CM31{dp.y, 0}), sx.b));
    QM31 num = q_zero(), al = q_fromm(1);
    QM31 c = q_sub(q_conj(sy), sy);
    for (uint32_t j = 0; j < f.sizes[b]; j++, e++) {
      al = q_mul(al, alpha);
      QM31 v = q_make(f.entry_vals[4 * e], f.entry_vals[4 * e + 1], f.entry_vals[4 * e + 2], f.entry_vals[4 * e + 3]);
      QM31 a = q_sub(q_conj(v), v);
      QM31 bb = q_sub(q_mul(v, c), q_mul(a, sy));
      QM31 value = q_mulm(q_mul(al, c), row_vals[f.entry_cols[e]]);
      QM31 lin = q_add(q_mulm(q_mul(al, a), dp.y), q_mul(al, bb));
      num = q_add(num, q_sub(value, lin));
    }
    acc = q_add(q_mul(acc, q_pow(alpha, f.sizes[b])), q_mulc(num, c_inv(den)));
  }
  return acc;
}

// ---- Queries (core/queries.rs)
struct Queries {
  std::vector<size_t> positions;
  uint32_t log_domain_size;
  static Queries generate(Channel& ch, uint32_t log_domain_size, uint32_t n_queries) {
    std::set<size_t> s;
    uint32_t cnt = 0;
    size_t max_query = ((size_t)1 << log_domain_size) - 1;
    for (;;) {
      Hash w = ch.draw_random_bytes();
      for (uint32_t x : w) {
        s.insert((size_t)x & max_query);
        if (++cnt == n_queries) return {std::vector<size_t>(s.begin(), s.end()), log_domain_size};
      }
    }
  }
  Queries fold(uint32_t n) const {
    std::vector<size_t> p;
    for (size_t q : positions) if (p.empty() || p.back() != (q >> n)) p.push_back(q >> n);
    return {p, log_domain_size - n};
  }
};

// compute_decommitment_positions (the fold cosets that contain a query), shared by prover and verifier
inline std::vector<size_t> decommitment_positions(const std::vector<size_t>& queries, uint32_t fold_step) {
  std::vector<size_t> out;
  size_t i = 0;
  while (i < queries.size()) {
    size_t start = (queries[i] >> fold_step) << fold_step;
    for (size_t p = start; p < start + ((size_t)1 << fold_step); p++) out.push_back(p);
    while (i < queries.size() && (queries[i] >> fold_step) == (start >> fold_step)) i++;
  }
  return out;
}

// ---- JSON in the shape serde_json gives the reference's `BrainfuckProof` (bin/brainfuck_prover.rs:127-131 writes it, :145-152
// reads it back for `verify`).  In-tree part: `Claim { log_size, _marker: PhantomData }` derives Serialize without a skip
// attribute (components/mod.rs:85-93), so every claim carries `"_marker":null` and the reference's Deserialize REQUIRES it;
// `InteractionClaim { claimed_sum }` (:70-76).  Upstream part [U]: M31 -> number, CM31 -> [a,b], QM31 -> [[a,b],[c,d]] (tuple
// structs); `Blake2sHash(pub [u8; 32])` derives Serialize on a newtype over a byte array -> an array of 32 numbers.
inline std::string hex(const Hash& h) {   // the same digest as `[b0,b1,...,b31]`
  std::string s = "[";
  const uint8_t* b = (const uint8_t*)h.data();
  for (int i = 0; i < 32; i++) { if (i) s.push_back(','); s += std::to_string((unsigned)b[i]); }
  s.push_back(']');
  return s;
}
// Append-only text writer: std::to_chars into one pre-sized string.  (The first version streamed ~50 000 numbers through an
// ostringstream: 2.1 ms per fib19 proof on the GPU box's host — inside the end-to-end time, 4 % of it on one GPU and 12 % on
// eight; this one takes 0.2 ms for the same bytes.)
struct JsonOut {
  std::string s;   // s[0 .. n) is the text so far; the rest is spare room that every writer checks before it writes
  size_t n = 0;
  JsonOut() { s.resize(1 << 18); }
  char* room(size_t k) { if (n + k > s.size()) s.resize(std::max(s.size() * 2, n + k)); return &s[n]; }
  JsonOut& operator<<(const char* t) { const size_t k = strlen(t); memcpy(room(k), t, k); n += k; return *this; }
  JsonOut& operator<<(uint64_t v) { char* w = room(24); n += (size_t)(std::to_chars(w, w + 24, v).ptr - w); return *this; }
  JsonOut& operator<<(uint32_t v) { char* w = room(12); n += (size_t)(std::to_chars(w, w + 12, v).ptr - w); return *this; }
  JsonOut& operator<<(const Hash& h) {   // `[b0,b1,...,b31]`: most of a proof's text
    const uint8_t* b = (const uint8_t*)h.data();
    char* const w0 = room(2 + 32 * 4);
    char* w = w0;
    *w++ = '[';
    for (int i = 0; i < 32; i++) {
      if (i) *w++ = ',';
      const unsigned v = b[i];
      if (v >= 100) { *w++ = (char)('0' + v / 100); *w++ = (char)('0' + (v / 10) % 10); *w++ = (char)('0' + v % 10); }
      else if (v >= 10) { *w++ = (char)('0' + v / 10); *w++ = (char)('0' + v % 10); }
      else *w++ = (char)('0' + v);
    }
    *w++ = ']';
    n += (size_t)(w - w0);
    return *this;
  }
  std::string str() { s.resize(n); return std::move(s); }
};
inline void jq(JsonOut& o, const QM31& q) { o << "[[" << q.a.a << "," << q.a.b << "],[" << q.b.a << "," << q.b.b << "]]"; }
inline void jdec(JsonOut& o, const MerkleDecommitment& d) {
  o << "{\"hash_witness\":[";
  for (size_t i = 0; i < d.hash_witness.size(); i++) o << (i ? "," : "") << d.hash_witness[i];
  o << "],\"column_witness\":[";
  for (size_t i = 0; i < d.column_witness.size(); i++) o << (i ? "," : "") << d.column_witness[i];
  o << "]}";
}
inline void jlayer(JsonOut& o, const FriLayerProof& l) {
  o << "{\"fri_witness\":[";
  for (size_t i = 0; i < l.fri_witness.size(); i++) { if (i) o << ","; jq(o, l.fri_witness[i]); }
  o << "],\"decommitment\":";
  jdec(o, l.decommitment);
  o << ",\"commitment\":" << l.commitment << "}";
}
inline std::string proof_to_json(const BrainfuckProof& p) {
  JsonOut o;
  o << "{\"claim\":{";
  for (int c = 0; c < N_COMPONENTS; c++) o << (c ? "," : "") << "\"" << COMPONENT_NAMES[c] << "\":{\"log_size\":" << p.log_size[c] << ",\"_marker\":null}";
  o << "},\"interaction_claim\":{";
  for (int c = 0; c < N_COMPONENTS; c++) { o << (c ? "," : "") << "\"" << COMPONENT_NAMES[c] << "\":{\"claimed_sum\":"; jq(o, p.claimed_sum[c]); o << "}"; }
  const CommitmentSchemeProof& s = p.proof;
  o << "},\"proof\":{\"commitments\":[";
  for (size_t i = 0; i < s.commitments.size(); i++) o << (i ? "," : "") << s.commitments[i];
  o << "],\"sampled_values\":[";
  for (size_t t = 0; t < s.sampled_values.size(); t++) {
    o << (t ? "," : "") << "[";
    for (size_t c = 0; c < s.sampled_values[t].size(); c++) {
      o << (c ? "," : "") << "[";
      for (size_t k = 0; k < s.sampled_values[t][c].size(); k++) { if (k) o << ","; jq(o, s.sampled_values[t][c][k]); }
      o << "]";
    }
    o << "]";
  }
  o << "],\"decommitments\":[";
  for (size_t t = 0; t < s.decommitments.size(); t++) { if (t) o << ","; jdec(o, s.decommitments[t]); }
  o << "],\"queried_values\":[";
  for (size_t t = 0; t < s.queried_values.size(); t++) {
    o << (t ? "," : "") << "[";
    for (size_t c = 0; c < s.queried_values[t].size(); c++) {
      o << (c ? "," : "") << "[";
      for (size_t k = 0; k < s.queried_values[t][c].size(); k++) o << (k ? "," : "") << s.queried_values[t][c][k];
      o << "]";
    }
    o << "]";
  }
  o << "],\"proof_of_work\":" << (uint64_t)s.proof_of_work << ",\"fri_proof\":{\"first_layer\":";
  jlayer(o, s.fri_proof.first_layer);
  o << ",\"inner_layers\":[";
  for (size_t i = 0; i < s.fri_proof.inner_layers.size(); i++) { if (i) o << ","; jlayer(o, s.fri_proof.inner_layers[i]); }
  // LinePoly { coeffs: Vec<SecureField>, log_size: u32 } [U core/poly/line.rs]: a struct with named fields -> an object
  o << "],\"last_layer_poly\":{\"coeffs\":[";
  for (size_t i = 0; i < s.fri_proof.last_layer_poly.size(); i++) { if (i) o << ","; jq(o, s.fri_proof.last_layer_poly[i]); }
  uint32_t ll_log = 0;
  while (((size_t)1 << ll_log) < s.fri_proof.last_layer_poly.size()) ll_log++;
  o << "],\"log_size\":" << ll_log << "}}}}";
  return o.str();
}

// ---- The way back: serde_json text -> BrainfuckProof (what `brainfuck_prover verify` does with `serde_json::from_str`,
// bin/brainfuck_prover.rs:145-152).  A small recursive-descent reader for exactly the shapes proof_to_json writes; every
// structural surprise is an error, never a guess.
struct JsonReader {
  const char* p; const char* end;
  explicit JsonReader(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
  [[noreturn]] void bad(const char* what) const { throw std::runtime_error(std::string("proof JSON: ") + what); }
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool peek(char c) { ws(); return p < end && *p == c; }
  void expect(char c) { ws(); if (p >= end || *p != c) bad("unexpected character"); p++; }
  bool maybe(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
  uint64_t number() {
    ws();
    if (p >= end || *p < '0' || *p > '9') bad("number expected");
    uint64_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') { if (v > (UINT64_MAX - 9) / 10) bad("number too large"); v = v * 10 + (uint64_t)(*p - '0'); p++; }
    return v;
  }
  uint32_t m31() { uint64_t v = number(); if (v >= sb::P) bad("field element out of range"); return (uint32_t)v; }
  void key(const char* name) {
    expect('"');
    size_t n = strlen(name);
    if ((size_t)(end - p) < n + 1 || memcmp(p, name, n) != 0 || p[n] != '"') bad("unexpected field name");
    p += n + 1;
    expect(':');
  }
  void null() { ws(); if (end - p < 4 || memcmp(p, "null", 4) != 0) bad("null expected"); p += 4; }
  template <class F> void array(F item) {   // item() reads one element
    expect('[');
    if (maybe(']')) return;
    do { item(); } while (maybe(','));
    expect(']');
  }
  QM31 qm31() {
    expect('['); expect('['); uint32_t a = m31(); expect(','); uint32_t b = m31(); expect(']'); expect(',');
    expect('['); uint32_t c = m31(); expect(','); uint32_t d = m31(); expect(']'); expect(']');
    return q_make(a, b, c, d);
  }
  Hash hash() {
    uint8_t b[32]; size_t n = 0;
    array([&] { uint64_t v = number(); if (v > 255 || n >= 32) bad("digest byte"); b[n++] = (uint8_t)v; });
    if (n != 32) bad("digest length");
    Hash h; memcpy(h.data(), b, 32); return h;
  }
  MerkleDecommitment decommitment() {
    MerkleDecommitment d;
    expect('{'); key("hash_witness"); array([&] { d.hash_witness.push_back(hash()); });
    expect(','); key("column_witness"); array([&] { d.column_witness.push_back(m31()); });
    expect('}');
    return d;
  }
  FriLayerProof layer() {
    FriLayerProof l;
    expect('{'); key("fri_witness"); array([&] { l.fri_witness.push_back(qm31()); });
    expect(','); key("decommitment"); l.decommitment = decommitment();
    expect(','); key("commitment"); l.commitment = hash();
    expect('}');
    return l;
  }
};
inline BrainfuckProof proof_from_json(const std::string& text) {
  JsonReader r(text);
  BrainfuckProof p;
  r.expect('{'); r.key("claim"); r.expect('{');
  for (int c = 0; c < N_COMPONENTS; c++) {
    if (c) r.expect(',');
    r.key(COMPONENT_NAMES[c]); r.expect('{'); r.key("log_size");
    uint64_t v = r.number(); if (v > 31) r.bad("log_size");
    p.log_size[c] = (uint32_t)v;
    r.expect(','); r.key("_marker"); r.null(); r.expect('}');
  }
  r.expect('}'); r.expect(','); r.key("interaction_claim"); r.expect('{');
  for (int c = 0; c < N_COMPONENTS; c++) {
    if (c) r.expect(',');
    r.key(COMPONENT_NAMES[c]); r.expect('{'); r.key("claimed_sum"); p.claimed_sum[c] = r.qm31(); r.expect('}');
  }
  r.expect('}'); r.expect(','); r.key("proof"); r.expect('{');
  CommitmentSchemeProof& s = p.proof;
  r.key("commitments"); r.array([&] { s.commitments.push_back(r.hash()); });
  r.expect(','); r.key("sampled_values");
  r.array([&] { s.sampled_values.emplace_back(); r.array([&] { s.sampled_values.back().emplace_back(); r.array([&] { s.sampled_values.back().back().push_back(r.qm31()); }); }); });
  r.expect(','); r.key("decommitments"); r.array([&] { s.decommitments.push_back(r.decommitment()); });
  r.expect(','); r.key("queried_values");
  r.array([&] { s.queried_values.emplace_back(); r.array([&] { s.queried_values.back().emplace_back(); r.array([&] { s.queried_values.back().back().push_back(r.m31()); }); }); });
  r.expect(','); r.key("proof_of_work"); s.proof_of_work = r.number();
  r.expect(','); r.key("fri_proof"); r.expect('{');
  r.key("first_layer"); s.fri_proof.first_layer = r.layer();
  r.expect(','); r.key("inner_layers"); r.array([&] { s.fri_proof.inner_layers.push_back(r.layer()); });
  r.expect(','); r.key("last_layer_poly"); r.expect('{'); r.key("coeffs"); r.array([&] { s.fri_proof.last_layer_poly.push_back(r.qm31()); });
  r.expect(','); r.key("log_size");
  uint64_t ll = r.number();
  if (ll > 20 || s.fri_proof.last_layer_poly.size() != ((size_t)1 << ll)) r.bad("last_layer_poly size");
  r.expect('}'); r.expect('}'); r.expect('}'); r.expect('}');
  r.ws();
  if (r.p != r.end) r.bad("trailing characters");
  return p;
}

}  // namespace sbf
