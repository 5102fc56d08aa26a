"""GPU parity for the lane-repeated column ops (include/stwo_cuda.h `*_repeated`): each must equal the plain op applied to
the expanded column — computed here by the CPU oracle — bit for bit."""
import numpy as np
import pytest

from oracle_lib import P

pytestmark = pytest.mark.gpu
ROOT_LOG = 20
REP = 4


@pytest.fixture(scope="module")
def tw(be):
    return be.precompute_twiddles(ROOT_LOG)


def rnd(seed, n):
    return np.random.default_rng(seed).integers(0, P, size=n, dtype=np.uint32)


@pytest.mark.parametrize("m", [0, 1, 2, 3, 4, 7, 10, 13, 14, 16])
def test_interpolate_and_evaluate_repeated(be, orc, tw, m):
    vals = [rnd(31 * m + k, 1 << m) for k in range(3)]
    src = [be.column(v) for v in vals]
    cols = be.interpolate_repeated(src, REP, tw, in_place=False)
    for c, v in zip(src, vals):
        assert (c.to_cpu() == v).all(), "out-of-place interpolate must leave its input alone"
    be.interpolate_repeated(src, REP, tw)
    for c, d in zip(src, cols):
        assert (c.to_cpu() == d.to_cpu()).all(), "in-place and out-of-place interpolate agree"
    full = [orc.interpolate(np.repeat(v, 1 << REP), ROOT_LOG) for v in vals]
    for c, f in zip(cols, full):
        assert not f.reshape(-1, 1 << REP)[:, 1:].any(), "oracle: coefficients off the 16-grid must vanish"
        assert (c.to_cpu() == f[:: 1 << REP]).all(), f"compact coefficients, m={m}"
    for blow in (0, 1):
        ev = be.evaluate_repeated(cols, REP, blow, tw)
        for e, f in zip(ev, full):
            want = orc.evaluate(f, blow, ROOT_LOG)
            got = e.to_cpu()
            assert len(got) == len(want) and (got == want).all(), f"evaluate_repeated m={m} blowup={blow}"
            assert (got.reshape(-1, 1 << REP) == got.reshape(-1, 1 << REP)[:, :1]).all()   # the LDE repeats as well


def test_interpolate_repeated_mixed_sizes_other_rep_and_errors(be, orc, tw, pkg):
    ms = [0, 5, 2, 12, 5, 9]
    vals = [rnd(900 + i, 1 << m) for i, m in enumerate(ms)]
    for rep in (2, 4, 5):
        cols = [be.column(v) for v in vals]
        be.interpolate_repeated(cols, rep, tw)
        for c, v in zip(cols, vals):
            assert (c.to_cpu() == orc.interpolate(np.repeat(v, 1 << rep), ROOT_LOG)[:: 1 << rep]).all()
        ev = be.evaluate_repeated(cols, rep, 1, tw)
        for e, c in zip(ev, cols):
            full = np.zeros(len(c) << rep, dtype=np.uint32)
            full[:: 1 << rep] = c.to_cpu()
            assert (e.to_cpu() == orc.evaluate(full, 1, ROOT_LOG)).all()
    be.interpolate_repeated([], REP, tw)
    with pytest.raises(pkg.BackendError):
        be.interpolate_repeated([be.column(np.zeros(12, dtype=np.uint32))], REP, tw)          # not a power of two
    with pytest.raises(pkg.BackendError):
        be.interpolate_repeated([be.zeros(1 << (ROOT_LOG - 2))], REP, tw)                     # full domain exceeds the twiddle tree
    with pytest.raises(pkg.BackendError):
        be.evaluate_repeated([be.zeros(16)], REP, 2, tw)                                       # blow-up > 1 unsupported here


@pytest.mark.parametrize("m", [0, 1, 3, 9, 13, 15])
def test_eval_at_point_repeated(be, orc, m):
    c = rnd(m + 150, 1 << m)
    plain = rnd(m + 151, 1 << 6)
    pts = rnd(m + 152, 16).reshape(2, 8)
    got = be.eval_at_point_repeated([be.column(c), be.column(plain)], [REP, 0], pts)
    full = np.zeros((1 << m) << REP, dtype=np.uint32)
    full[:: 1 << REP] = c
    assert (got[0] == orc.eval_at_point(full, pts[0])).all()
    assert (got[1] == orc.eval_at_point(plain, pts[1])).all()


def test_merkle_commit_repeated(be, orc):
    logs = [13, 11, 13, 6, 11, 11, 5, 13, 8, 4]
    host = [np.repeat(rnd(2500 + i, (1 << lg) >> REP), 1 << REP) for i, lg in enumerate(logs)]
    cols = [be.column(c) for c in host]
    layers, root = be.merkle_commit_repeated(cols, REP)
    plain_layers, plain_root = be.merkle_commit(cols)
    ref = orc.merkle_commit(host)
    assert len(layers) == len(ref) == 14
    for k, (g, pl, r) in enumerate(zip(layers, plain_layers, ref)):
        assert (g.to_cpu() == r).all() and (pl.to_cpu() == r).all(), f"layer {k}"
    assert (root == ref[0]).all() and (plain_root == ref[0]).all()
