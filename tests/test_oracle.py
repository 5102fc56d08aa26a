"""CPU tests that pin the oracle: published KATs (RFC 7693, hashlib), group identities, and FFT == point evaluation.
The reference holds no golden vector at the Backend boundary, so these are the strongest pins available."""
import hashlib

import numpy as np
import pytest

from oracle_lib import P

ROOT_LOG = 12


def brev(i, n):
    return int(format(i, "0%db" % n)[::-1], 2) if n else 0


def test_blake2s_rfc7693_abc(orc):
    # RFC 7693 Appendix B
    want = "508c5e8c327c14e2e1a72ba34eeb452f37458b209ed63a294d999b4c86675982"
    assert orc.blake2s256(b"abc").hex() == want


@pytest.mark.parametrize("n", [0, 1, 31, 32, 63, 64, 65, 127, 128, 129, 1000])
def test_blake2s_vs_hashlib(orc, n):
    d = bytes(np.random.default_rng(n).integers(0, 256, size=n, dtype=np.uint8))
    assert orc.blake2s256(d) == hashlib.blake2s(d).digest()


def test_compress_matches_hashlib_single_block(orc):
    # Blake2s-256 of a <=64-byte message is one F call on the parameter-block state with t0 = len, f0 = ~0.
    iv = np.array([0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19], dtype=np.uint32)
    h = iv.copy()
    h[0] ^= 0x01010020
    msg = bytes(range(40))
    m = np.frombuffer(msg + bytes(24), dtype="<u4")
    out = orc.compress(h, m, t0=40, f0=0xFFFFFFFF)
    assert out.tobytes() == hashlib.blake2s(msg).digest()


def test_hash_node_conventions(orc):
    # neither children nor columns -> 32 zero bytes; one column value -> F(0, [v,0..]) with zero counters
    assert (orc.commit_on_layer(0, None, []) == 0).all()
    v = np.array([12345], dtype=np.uint32)
    m = np.zeros(16, dtype=np.uint32)
    m[0] = 12345
    assert (orc.commit_on_layer(0, None, [v]) == orc.compress(np.zeros(8, dtype=np.uint32), m)).all()


def test_circle_generator_and_domains(orc):
    gx, gy = 2, 1268011823
    assert (gx * gx + gy * gy) % P == 1
    # canonic domains: points on the circle, second half = conjugates, different sizes disjoint
    seen = set()
    for log in [1, 2, 3, 4, 5]:
        pts = [orc.domain_at(log, i) for i in range(1 << log)]
        for (x, y) in pts:
            assert (x * x + y * y) % P == 1
        h = 1 << (log - 1)
        for i in range(h):
            assert pts[i + h] == (pts[i][0], (P - pts[i][1]) % P)
        assert not (seen & set(pts))
        seen |= set(pts)


@pytest.mark.parametrize("log", [1, 2, 3, 4, 5, 7, 9])
def test_fft_is_point_evaluation(orc, log):
    rng = np.random.default_rng(100 + log)
    v = rng.integers(0, P, size=1 << log, dtype=np.uint32)
    c = orc.interpolate(v, ROOT_LOG)
    assert (orc.evaluate(c, 0, ROOT_LOG) == v).all()
    lde = orc.evaluate(c, 1, ROOT_LOG)
    for k in range(2 << log):
        x, y = orc.domain_at(log + 1, brev(k, log + 1))
        got = orc.eval_at_point(c, [x, 0, 0, 0, y, 0, 0, 0])
        assert got[0] == lde[k] and not got[1:].any()
    # trace-domain values are reproduced by point evaluation too
    for k in range(1 << log):
        x, y = orc.domain_at(log, brev(k, log))
        assert orc.eval_at_point(c, [x, 0, 0, 0, y, 0, 0, 0])[0] == v[k]


def test_interpolate_low_degree_basis(orc):
    # f = 3 + 5y + 7x on CanonicCoset(4): coefficients land on indices 0 (1), 1 (y), 2 (x)
    log = 4
    vals = np.zeros(1 << log, dtype=np.uint32)
    for k in range(1 << log):
        x, y = orc.domain_at(log, brev(k, log))
        vals[k] = (3 + 5 * y + 7 * x) % P
    c = orc.interpolate(vals, ROOT_LOG)
    want = np.zeros(1 << log, dtype=np.uint32)
    want[0], want[1], want[2] = 3, 5, 7
    assert (c == want).all()


def test_twiddle_tree_layout(orc):
    tw, itw = orc.twiddles(6)
    assert tw[-1] == 1 and ((tw.astype(np.uint64) * itw) % P == 1).all()
    # level 0 = x of the first half of half_odds(6) in bit-reversed order
    from_oracle = [orc.domain_at(7, brev(i, 5))[0] for i in range(32)]  # CanonicCoset(7).circle_domain().half_coset = half_odds(6)
    assert list(tw[:32]) == from_oracle


def test_fold_identities(orc):
    # folding the LDE of a polynomial with FRI folds down to a constant layer (degree bound 1 at blowup 2)
    rng = np.random.default_rng(5)
    log = 6
    coords = []
    for _ in range(4):
        c = np.zeros(1 << log, dtype=np.uint32)
        c[: 1 << (log - 1)] = rng.integers(0, P, size=1 << (log - 1), dtype=np.uint32)  # degree < 2^(log-1)
        coords.append(orc.evaluate(c, 0, ROOT_LOG))
    alpha = [1, 2, 3, 4]
    line = orc.fold_circle_into_line([np.zeros(1 << (log - 1), dtype=np.uint32)] * 4, coords, alpha)
    while line[0].size > 2:
        line = orc.fold_line(line, alpha)
    assert all(l[0] == l[1] for l in line)


def test_prefix_sum_last_row_is_storage_index_1(orc):
    log = 5
    v = np.arange(1, 33, dtype=np.uint32)
    s = orc.prefix_sum_bitrev(v)
    assert s[1] == v.sum() % P  # claimed_sum = col.at(1)
    assert s[0] == v[0]         # coset row 0 is storage index 0


def test_grind_smallest_nonce(orc):
    d = np.arange(8, dtype=np.uint32)
    n = orc.grind(d, 5)
    for k in range(n + 1):
        m = np.zeros(16, dtype=np.uint32)
        m[0] = k
        h = orc.compress(d, m)
        tz = (int(h[0]) & -int(h[0])).bit_length() - 1 if h[0] else 32
        assert (tz >= 5) == (k == n)
