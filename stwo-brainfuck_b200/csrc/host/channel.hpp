// Blake2sChannel + Blake2sMerkleChannel — the Fiat–Shamir transcript (host).  Restates stwo-prover 0.1.1 @ 31e8dbc
// core/channel/blake2s.rs and core/vcs/blake2_merkle.rs; used by the reference at
// crates/brainfuck_prover/src/brainfuck_air/mod.rs:485,564-581,591,704-721 and components/mod.rs:82,133.
#pragma once
#include <array>
#include <cstring>
#include <vector>
#include "../m31.cuh"

namespace sbf {
using sb::QM31;

namespace b2s {
static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
inline uint32_t ror(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
inline void compress(uint32_t h[8], const uint32_t m[16], uint64_t t, uint32_t f0) {
  uint32_t v[16];
  for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = IV[i]; }
  v[12] ^= (uint32_t)t; v[13] ^= (uint32_t)(t >> 32); v[14] ^= f0;
#define SBF_G(a, b, c, d, x, y) \
  v[a] += v[b] + (x); v[d] = ror(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = ror(v[b] ^ v[c], 12); \
  v[a] += v[b] + (y); v[d] = ror(v[d] ^ v[a], 8);  v[c] += v[d]; v[b] = ror(v[b] ^ v[c], 7);
  for (int r = 0; r < 10; r++) {
    const uint8_t* s = SIGMA[r];
    SBF_G(0, 4, 8, 12, m[s[0]], m[s[1]]) SBF_G(1, 5, 9, 13, m[s[2]], m[s[3]])
    SBF_G(2, 6, 10, 14, m[s[4]], m[s[5]]) SBF_G(3, 7, 11, 15, m[s[6]], m[s[7]])
    SBF_G(0, 5, 10, 15, m[s[8]], m[s[9]]) SBF_G(1, 6, 11, 12, m[s[10]], m[s[11]])
    SBF_G(2, 7, 8, 13, m[s[12]], m[s[13]]) SBF_G(3, 4, 9, 14, m[s[14]], m[s[15]])
  }
#undef SBF_G
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
}
// unkeyed Blake2s-256
inline std::array<uint32_t, 8> hash(const uint8_t* data, size_t len) {
  uint32_t h[8];
  for (int i = 0; i < 8; i++) h[i] = IV[i];
  h[0] ^= 0x01010020u;
  uint64_t t = 0;
  uint32_t m[16];
  while (len > 64) { memcpy(m, data, 64); t += 64; compress(h, m, t, 0); data += 64; len -= 64; }
  uint8_t last[64] = {0};
  memcpy(last, data, len);
  memcpy(m, last, 64);
  t += len;
  compress(h, m, t, 0xFFFFFFFFu);
  std::array<uint32_t, 8> out;
  memcpy(out.data(), h, 32);
  return out;
}
// Blake2sMerkleHasher::hash_node (zero initial state, zero counters/flags)
inline std::array<uint32_t, 8> hash_node(const uint32_t* children16, const uint32_t* vals, size_t n) {
  uint32_t st[8] = {0};
  if (children16) compress(st, children16, 0, 0);
  for (size_t o = 0; o < n; o += 16) {
    uint32_t m[16] = {0};
    for (size_t j = 0; j < 16 && o + j < n; j++) m[j] = vals[o + j];
    compress(st, m, 0, 0);
  }
  std::array<uint32_t, 8> out;
  memcpy(out.data(), st, 32);
  return out;
}
}  // namespace b2s

typedef std::array<uint32_t, 8> Hash;

struct Channel {
  Hash digest{};  // 32 zero bytes
  uint32_t n_sent = 0;

  void update(const Hash& d) { digest = d; n_sent = 0; }
  void mix_root(const Hash& root) {
    uint32_t buf[16];
    memcpy(buf, digest.data(), 32);
    memcpy(buf + 8, root.data(), 32);
    update(b2s::hash((const uint8_t*)buf, 64));
  }
  void mix_felts(const std::vector<QM31>& felts) {
    std::vector<uint32_t> buf(8 + 4 * felts.size());
    memcpy(buf.data(), digest.data(), 32);
    for (size_t i = 0; i < felts.size(); i++) {
      buf[8 + 4 * i] = felts[i].a.a; buf[9 + 4 * i] = felts[i].a.b; buf[10 + 4 * i] = felts[i].b.a; buf[11 + 4 * i] = felts[i].b.b;
    }
    update(b2s::hash((const uint8_t*)buf.data(), buf.size() * 4));
  }
  void mix_u64(uint64_t v) {
    uint32_t h[8], m[16] = {0};
    memcpy(h, digest.data(), 32);
    m[0] = (uint32_t)v; m[1] = (uint32_t)(v >> 32);
    b2s::compress(h, m, 0, 0);
    Hash d;
    memcpy(d.data(), h, 32);
    update(d);
  }
  Hash draw_random_bytes() {
    uint32_t buf[16] = {0};
    memcpy(buf, digest.data(), 32);
    buf[8] = n_sent;  // counter padded to 32 bytes, little endian
    n_sent++;
    return b2s::hash((const uint8_t*)buf, 64);
  }
  std::array<uint32_t, 8> draw_base_felts() {
    for (;;) {
      Hash w = draw_random_bytes();
      bool ok = true;
      for (uint32_t x : w) if (x >= 2u * sb::P) ok = false;
      if (!ok) continue;
      std::array<uint32_t, 8> out;
      for (int i = 0; i < 8; i++) out[i] = w[i] >= sb::P ? w[i] - sb::P : w[i];
      return out;
    }
  }
  QM31 draw_felt() {
    auto f = draw_base_felts();
    return sb::q_make(f[0], f[1], f[2], f[3]);
  }
  std::vector<QM31> draw_felts(size_t n) {
    std::vector<QM31> out;
    while (out.size() < n) {
      auto f = draw_base_felts();
      out.push_back(sb::q_make(f[0], f[1], f[2], f[3]));
      if (out.size() < n) out.push_back(sb::q_make(f[4], f[5], f[6], f[7]));
    }
    return out;
  }
  uint32_t trailing_zeros() const {
    for (int w = 0; w < 4; w++) if (digest[w]) return 32 * w + (uint32_t)__builtin_ctz(digest[w]);
    return 128;
  }
};

}  // namespace sbf
