// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_field.h header).
//
// Blake2s compression (RFC 7693 F function), Blake2s-256, Blake2sMerkleHasher::hash_node and
// MerkleProver::commit as used by stwo-prover 0.1.1 @ 31e8dbc (upstream core/vcs/{blake2s_ref,
// blake2_hash,blake2_merkle,prover}.rs — absent here).
// Pinned against: RFC 7693 Appendix B "abc" known-answer and python hashlib.blake2s (tests/).
// hash_node starts from an ALL-ZERO state and uses zero counters/flags — that convention is recalled, unpinned.
// Reference call sites: crates/brainfuck_prover/src/brainfuck_air/mod.rs:500,583,723 (tree_builder.commit).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace orc {

static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint32_t rotr32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }

// h' = F(h, m, t0, t1, f0, f1).  One definition over the word type W: uint32_t (the scalar form every test pins against
// RFC 7693 / hashlib) or a vector of 16 u32 lanes hashing 16 independent rows at once, which is how upstream's SimdBackend
// runs it (`compress16`, core/backend/simd/blake2s.rs).
template <class W>
static inline __attribute__((always_inline)) void b2s_compress_t(W h[8], const W m[16], uint32_t t0, uint32_t t1, uint32_t f0, uint32_t f1) {
  W v[16];
  for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = (h[i] ^ h[i]) + B2S_IV[i]; }
  v[12] ^= t0; v[13] ^= t1; v[14] ^= f0; v[15] ^= f1;
#define ORC_ROTR(x, r) (((x) >> (r)) | ((x) << (32 - (r))))
#define ORC_G(a, b, c, d, x, y)                                  \
  v[a] = v[a] + v[b] + (x); v[d] = ORC_ROTR(v[d] ^ v[a], 16);    \
  v[c] = v[c] + v[d];       v[b] = ORC_ROTR(v[b] ^ v[c], 12);    \
  v[a] = v[a] + v[b] + (y); v[d] = ORC_ROTR(v[d] ^ v[a], 8);     \
  v[c] = v[c] + v[d];       v[b] = ORC_ROTR(v[b] ^ v[c], 7);
  for (int r = 0; r < 10; r++) {
    const uint8_t* s = B2S_SIGMA[r];
    ORC_G(0, 4, 8, 12, m[s[0]], m[s[1]])   ORC_G(1, 5, 9, 13, m[s[2]], m[s[3]])
    ORC_G(2, 6, 10, 14, m[s[4]], m[s[5]])  ORC_G(3, 7, 11, 15, m[s[6]], m[s[7]])
    ORC_G(0, 5, 10, 15, m[s[8]], m[s[9]])  ORC_G(1, 6, 11, 12, m[s[10]], m[s[11]])
    ORC_G(2, 7, 8, 13, m[s[12]], m[s[13]]) ORC_G(3, 4, 9, 14, m[s[14]], m[s[15]])
  }
#undef ORC_G
#undef ORC_ROTR
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
}
static inline void b2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1, uint32_t f0, uint32_t f1) {
  b2s_compress_t<uint32_t>(h, m, t0, t1, f0, f1);
}

// Standard unkeyed Blake2s-256 (the `blake2` crate's Blake2s256) — used by the channel.
static inline void blake2s256(const uint8_t* data, size_t len, uint8_t out[32]) {
  uint32_t h[8];
  for (int i = 0; i < 8; i++) h[i] = B2S_IV[i];
  h[0] ^= 0x01010020u;
  uint64_t t = 0;
  uint32_t m[16];
  while (len > 64) {
    memcpy(m, data, 64);
    t += 64;
    b2s_compress(h, m, (uint32_t)t, (uint32_t)(t >> 32), 0, 0);
    data += 64; len -= 64;
  }
  uint8_t last[64] = {0};
  memcpy(last, data, len);
  memcpy(m, last, 64);
  t += len;
  b2s_compress(h, m, (uint32_t)t, (uint32_t)(t >> 32), 0xFFFFFFFFu, 0);
  memcpy(out, h, 32);
}

// Blake2sMerkleHasher::hash_node
static inline void hash_node(const uint32_t* children /*16 words or null*/, const uint32_t* vals, size_t nvals, uint32_t out[8]) {
  uint32_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (children) b2s_compress(st, children, 0, 0, 0, 0);
  for (size_t o = 0; o < nvals; o += 16) {
    uint32_t m[16] = {0};
    size_t k = nvals - o < 16 ? nvals - o : 16;
    for (size_t j = 0; j < k; j++) m[j] = vals[o + j];
    b2s_compress(st, m, 0, 0, 0, 0);
  }
  memcpy(out, st, 32);
}

// MerkleOps::commit_on_layer: row i hashes children (2i, 2i+1) of prev (if any) then element i of each column.
static inline void commit_on_layer_rows(size_t lo, size_t hi, const uint32_t* prev, const uint32_t* const* cols, size_t ncols, uint32_t* out) {
  for (size_t i = lo; i < hi; i++) {
    uint32_t small[64];
    std::vector<uint32_t> big;
    uint32_t* vals = small;
    if (ncols > 64) { big.resize(ncols); vals = big.data(); }
    for (size_t c = 0; c < ncols; c++) vals[c] = cols[c][i];
    hash_node(prev ? prev + 16 * i : nullptr, vals, ncols, out + 8 * i);
  }
}
// The same function on 16 consecutive rows at a time, one row per vector lane (upstream: SimdBackend::commit_on_layer over
// compress16).  Compiled for AVX-512 and used when the host has it; ORC_SCALAR=1 forces the scalar definition above
// (tests/test_oracle.py compares the two).
typedef uint32_t b2s_v16 __attribute__((vector_size(64), aligned(4)));
__attribute__((target("avx512f"))) static void commit_on_layer_rows16(size_t lo, size_t hi, const uint32_t* prev, const uint32_t* const* cols,
                                                                      size_t ncols, uint32_t* out) {
  for (size_t i = lo; i < hi; i += 16) {
    b2s_v16 st[8], m[16];
    for (int k = 0; k < 8; k++) st[k] = b2s_v16{};
    if (prev) {
      for (int j = 0; j < 16; j++) for (int r = 0; r < 16; r++) m[j][r] = prev[16 * (i + r) + j];
      b2s_compress_t<b2s_v16>(st, m, 0, 0, 0, 0);
    }
    for (size_t o = 0; o < ncols; o += 16) {
      size_t k = ncols - o < 16 ? ncols - o : 16;
      for (size_t j = 0; j < 16; j++) {
        if (j < k) memcpy(&m[j], cols[o + j] + i, 64);
        else m[j] = b2s_v16{};
      }
      b2s_compress_t<b2s_v16>(st, m, 0, 0, 0, 0);
    }
    for (int k = 0; k < 8; k++) for (int r = 0; r < 16; r++) out[8 * (i + r) + k] = st[k][r];
  }
}
static inline bool b2s_use_simd() {
  static const bool on = __builtin_cpu_supports("avx512f") && !getenv("ORC_SCALAR");
  return on;
}
static inline void commit_on_layer(uint32_t log_size, const uint32_t* prev, const uint32_t* const* cols, size_t ncols, uint32_t* out) {
  const size_t rows = (size_t)1 << log_size, chunk = 256;
  if (rows < chunk) { commit_on_layer_rows(0, rows, prev, cols, ncols, out); return; }
  const bool simd = b2s_use_simd();
#pragma omp parallel for schedule(static)
  for (size_t c = 0; c < rows / chunk; c++) {
    if (simd) commit_on_layer_rows16(c * chunk, (c + 1) * chunk, prev, cols, ncols, out);
    else commit_on_layer_rows(c * chunk, (c + 1) * chunk, prev, cols, ncols, out);
  }
}

// MerkleProver::commit: columns stable-sorted by length descending; layer `log` = commit_on_layer(log, layer log+1, cols of 2^log).
// layers[k] holds the layer of log size k (root = layers[0]).
static inline void merkle_commit(const uint32_t* const* cols, const uint32_t* logs, size_t ncols, std::vector<std::vector<uint32_t>>& layers) {
  uint32_t max_log = 0;
  for (size_t c = 0; c < ncols; c++) if (logs[c] > max_log) max_log = logs[c];
  layers.assign(max_log + 1, {});
  for (int lg = (int)max_log; lg >= 0; lg--) {
    std::vector<const uint32_t*> lc;
    for (size_t c = 0; c < ncols; c++) if (logs[c] == (uint32_t)lg) lc.push_back(cols[c]);  // stable order
    layers[lg].resize((size_t)8 << lg);
    commit_on_layer(lg, lg == (int)max_log ? nullptr : layers[lg + 1].data(), lc.data(), lc.size(), layers[lg].data());
  }
}

}  // namespace orc
