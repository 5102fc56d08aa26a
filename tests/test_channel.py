"""The Fiat–Shamir channel (csrc/host/channel.hpp — host logic shared by the product and the oracle, so the proof-equality
tests cannot see a mistake in it) against an independent model built on hashlib.blake2s and an RFC 7693 compression function
written out here.  Restated from Stwo @ 31e8dbc, core/channel/blake2s.rs and core/vcs/blake2_merkle.rs:
  mix_root    digest <- H(digest ‖ root)
  mix_felts   digest <- H(digest ‖ felts as little-endian u32 words, 4 per QM31)
  mix_u64     digest <- compress(h = digest, m = [lo, hi, 0...], t = 0, f = 0)        (raw compression, no parameter block)
  draw_random_bytes   H(digest ‖ counter as 32 little-endian bytes), counter += 1
  draw_base_felts     8 words of draw_random_bytes, retried until every word < 2P, each reduced mod P
  every mix resets the counter.  The reference reaches these at brainfuck_air/mod.rs:485,564-581,591,704-721."""
import ctypes
import hashlib
import struct

import numpy as np

P = (1 << 31) - 1
IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
SIGMA = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
         [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
         [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
         [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
         [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0]]
M32 = 0xFFFFFFFF


def compress(h, m, t=0, f=0):
    """RFC 7693 §3.2 F for BLAKE2s."""
    v = list(h) + list(IV)
    v[12] ^= t & M32
    v[13] ^= t >> 32
    v[14] ^= f
    ror = lambda x, r: ((x >> r) | (x << (32 - r))) & M32

    def g(a, b, c, d, x, y):
        v[a] = (v[a] + v[b] + x) & M32; v[d] = ror(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & M32; v[b] = ror(v[b] ^ v[c], 12)
        v[a] = (v[a] + v[b] + y) & M32; v[d] = ror(v[d] ^ v[a], 8)
        v[c] = (v[c] + v[d]) & M32; v[b] = ror(v[b] ^ v[c], 7)
    for r in range(10):
        s = SIGMA[r]
        g(0, 4, 8, 12, m[s[0]], m[s[1]]); g(1, 5, 9, 13, m[s[2]], m[s[3]])
        g(2, 6, 10, 14, m[s[4]], m[s[5]]); g(3, 7, 11, 15, m[s[6]], m[s[7]])
        g(0, 5, 10, 15, m[s[8]], m[s[9]]); g(1, 6, 11, 12, m[s[10]], m[s[11]])
        g(2, 7, 8, 13, m[s[12]], m[s[13]]); g(3, 4, 9, 14, m[s[14]], m[s[15]])
    return [h[i] ^ v[i] ^ v[8 + i] for i in range(8)]


def test_python_compress_is_blake2s():
    """the F above + the parameter block reproduces hashlib on one block (so mix_u64's raw use of it is anchored)"""
    msg = bytes(range(47))
    h = list(IV)
    h[0] ^= 0x01010020
    out = compress(h, list(struct.unpack("<16I", msg.ljust(64, b"\0"))), t=len(msg), f=M32)
    assert struct.pack("<8I", *out) == hashlib.blake2s(msg).digest()


class ModelChannel:
    def __init__(self):
        self.digest, self.n_sent = bytes(32), 0

    def _update(self, d):
        self.digest, self.n_sent = d, 0

    def mix_root(self, root):
        self._update(hashlib.blake2s(self.digest + root).digest())

    def mix_felts(self, felts):
        self._update(hashlib.blake2s(self.digest + b"".join(struct.pack("<4I", *f) for f in felts)).digest())

    def mix_u64(self, v):
        out = compress(list(struct.unpack("<8I", self.digest)), [v & M32, v >> 32] + [0] * 14)
        self._update(struct.pack("<8I", *out))

    def draw_random_bytes(self):
        d = hashlib.blake2s(self.digest + struct.pack("<Q", self.n_sent) + bytes(24)).digest()
        self.n_sent += 1
        return d

    def draw_base_felts(self):
        while True:
            w = struct.unpack("<8I", self.draw_random_bytes())
            if all(x < 2 * P for x in w):
                return [x - P if x >= P else x for x in w]

    def draw_felt(self):
        return self.draw_base_felts()[:4]

    def draw_felts(self, n):
        out = []
        while len(out) < n:
            f = self.draw_base_felts()
            out.append(f[:4])
            if len(out) < n:
                out.append(f[4:])
        return out

    def trailing_zeros(self):
        v = int.from_bytes(self.digest[:16], "little")
        return 128 if v == 0 else (v & -v).bit_length() - 1


def run_script(orc, script: bytes) -> bytes:
    lib = orc.lib
    lib.orc_channel_script.restype = ctypes.c_size_t
    out = ctypes.create_string_buffer(1 << 16)
    n = lib.orc_channel_script(script, ctypes.c_size_t(len(script)), out, ctypes.c_size_t(len(out)))
    assert 0 < n <= len(out)
    return out.raw[:n]


def test_channel_matches_the_hashlib_model(orc):
    rng = np.random.default_rng(0xC4A77E1)
    m = ModelChannel()
    script, want = b"", b""
    felt = lambda: [int(x) for x in rng.integers(0, P, size=4)]
    # the shape of a proof transcript: root, log sizes as u64s, root, draws, claimed sums, root, draws, nonce, many draws
    for step in range(200):
        op = "RUFDSBZ"[int(rng.integers(0, 7))] if step >= 12 else "RUURDDFRDSUB"[step]
        if op == "R":
            root = rng.bytes(32); script += b"R" + root; m.mix_root(root)
        elif op == "U":
            v = int(rng.integers(0, 1 << 62)) if step % 3 else int(rng.integers(0, 30)); script += b"U" + struct.pack("<Q", v); m.mix_u64(v)
        elif op == "F":
            fs = [felt() for _ in range(int(rng.integers(0, 14)))]
            script += b"F" + struct.pack("<I", len(fs)) + b"".join(struct.pack("<4I", *f) for f in fs); m.mix_felts(fs)
        elif op == "D":
            script += b"D"; want += struct.pack("<4I", *m.draw_felt())
        elif op == "S":
            n = int(rng.integers(1, 9)); script += b"S" + struct.pack("<I", n)
            want += b"".join(struct.pack("<4I", *f) for f in m.draw_felts(n))
        elif op == "B":
            script += b"B"; want += m.draw_random_bytes()
        else:
            script += b"Z"; want += struct.pack("<I", m.trailing_zeros())
    assert run_script(orc, script) == want + m.digest


def test_draws_reject_words_of_two_p_or_more(orc):
    """draw_base_felts retries until all eight words are < 2P: over many draws the model must have retried at least once
    (a word >= 2P has probability 2^-31 per word, so force it: search a digest whose first draw is rejected is infeasible —
    instead check the reduction branch: words in [P, 2P) come back minus P, which happens in about half of all words)."""
    m = ModelChannel()
    script = b""
    seen_high = False
    for i in range(64):
        script += b"U" + struct.pack("<Q", i) + b"D"
    got = run_script(orc, script)
    off = 0
    for i in range(64):
        m.mix_u64(i)
        raw = struct.unpack("<8I", hashlib.blake2s(m.digest + bytes(32)).digest())
        seen_high |= any(P <= x < 2 * P for x in raw[:4])
        assert list(struct.unpack("<4I", got[off:off + 16])) == m.draw_felt()
        off += 16
    assert seen_high
