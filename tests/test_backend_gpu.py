"""GPU parity: every Backend op through the C ABI vs the CPU oracle on the same seeded inputs (bit-exact)."""
import numpy as np
import pytest

from oracle_lib import P

pytestmark = pytest.mark.gpu
ROOT_LOG = 20


@pytest.fixture(scope="module")
def tw(be):
    return be.precompute_twiddles(ROOT_LOG)


def rnd(seed, n):
    return np.random.default_rng(seed).integers(0, P, size=n, dtype=np.uint32)


def test_twiddles_match_oracle(be, orc, tw):
    g_tw, g_itw = tw.to_cpu()
    o_tw, o_itw = orc.twiddles(ROOT_LOG)
    assert (g_tw == o_tw).all() and (g_itw == o_itw).all()


@pytest.mark.parametrize("log", [3, 4, 5, 6, 8, 10, 12, 13, 14, 15, 17, 19, 21])
def test_interpolate_evaluate(be, orc, tw, log):
    kinds = {"random": rnd(0x5EED0000 + log, 1 << log),
             "broadcast16": np.repeat(rnd(log, max(1, (1 << log) // 16)), 16)[: 1 << log],
             "is_first": np.eye(1, 1 << log, 0, dtype=np.uint32)[0]}
    cols = [be.column(v) for v in kinds.values()]
    be.interpolate_columns(cols, tw)
    refs = [orc.interpolate(v, ROOT_LOG) for v in kinds.values()]
    for c, r, k in zip(cols, refs, kinds):
        assert (c.to_cpu() == r).all(), f"interpolate {k} log {log}"
    if log + 1 <= ROOT_LOG + 1:
        ldes = be.evaluate_polynomials(cols, 1, tw)
        for l, r, k in zip(ldes, refs, kinds):
            assert (l.to_cpu() == orc.evaluate(r, 1, ROOT_LOG)).all(), f"evaluate {k} log {log}"
    same = be.evaluate_polynomials(cols[:1], 0, tw)
    assert (same[0].to_cpu() == list(kinds.values())[0]).all()


def test_interpolate_evaluate_log23(be, orc):
    """log 23 = 13 low layers + ONE strided pass of ten layers (8-word rows); log 22 still takes two 16-word passes."""
    root = 22
    tw23 = be.precompute_twiddles(root)
    v22, v23 = rnd(2201, 1 << 22), rnd(2301, 1 << 23)
    c22, c23 = be.column(v22), be.column(v23)
    be.interpolate_columns([c22, c23], tw23)
    r22, r23 = orc.interpolate(v22, root), orc.interpolate(v23, root)
    assert (c22.to_cpu() == r22).all() and (c23.to_cpu() == r23).all()
    lde = be.evaluate_polynomials([c22], 1, tw23)[0]
    assert (lde.to_cpu() == orc.evaluate(r22, 1, root)).all()
    back = be.evaluate_polynomials([c23], 0, tw23)[0]
    assert (back.to_cpu() == v23).all()


def test_interpolate_mixed_sizes_and_errors(be, orc, tw, pkg):
    logs = [4, 9, 4, 13, 16, 9]
    host = [rnd(i, 1 << lg) for i, lg in enumerate(logs)]
    cols = [be.column(h) for h in host]
    be.interpolate_columns(cols, tw)
    for c, h in zip(cols, host):
        assert (c.to_cpu() == orc.interpolate(h, ROOT_LOG)).all()
    with pytest.raises(pkg.BackendError):
        be.interpolate_columns([be.column(np.zeros(12, dtype=np.uint32))], tw)      # not a power of two
    with pytest.raises(pkg.BackendError):
        be.interpolate_columns([be.zeros(1 << (ROOT_LOG + 2))], tw)                  # twiddle tree too small
    be.interpolate_columns([], tw)                                                   # empty is fine


@pytest.mark.parametrize("log", [0, 3, 5, 11, 12, 16, 20])
def test_eval_at_point(be, orc, log):
    c = rnd(log + 50, 1 << log)
    pt = rnd(log + 51, 8)
    got = be.eval_at_point([be.column(c)], pt)
    assert (got[0] == orc.eval_at_point(c, pt)).all()


def test_eval_at_point_batch(be, orc):
    logs = [4, 13, 7, 18, 12]
    cs = [rnd(i + 70, 1 << lg) for i, lg in enumerate(logs)]
    pts = rnd(99, 8 * len(logs)).reshape(-1, 8)
    got = be.eval_at_point([be.column(c) for c in cs], pts)
    for g, c, p in zip(got, cs, pts):
        assert (g == orc.eval_at_point(c, p)).all()


@pytest.mark.parametrize("log,ncols,prev", [(0, 1, False), (0, 0, True), (3, 5, False), (5, 16, True), (8, 17, True),
                                            (10, 0, True), (12, 33, False), (14, 60, True), (16, 4, True)])
def test_commit_on_layer(be, orc, log, ncols, prev):
    cols = [rnd(1000 + i, 1 << log) for i in range(ncols)]
    pv = np.random.default_rng(log).integers(0, 2**32, size=16 << log, dtype=np.uint32) if prev else None
    got = be.commit_on_layer(log, be.column(pv) if prev else None, [be.column(c) for c in cols])
    assert (got.to_cpu() == orc.commit_on_layer(log, pv, cols)).all()


def test_merkle_commit_mixed(be, orc):
    logs = [12, 10, 12, 5, 10, 10, 4, 12, 7]
    cols = [rnd(2000 + i, 1 << lg) for i, lg in enumerate(logs)]
    layers, root = be.merkle_commit([be.column(c) for c in cols])
    ref = orc.merkle_commit(cols)
    assert len(layers) == len(ref) == 13
    for k, (g, r) in enumerate(zip(layers, ref)):
        assert (g.to_cpu() == r).all(), f"layer {k}"
    assert (root == ref[0]).all()


@pytest.mark.parametrize("log", [0, 1, 4, 9, 10, 11, 14, 17, 20])
def test_bit_reverse(be, orc, log):
    v = rnd(log, 1 << log)
    c = be.column(v)
    be.bit_reverse_column(c)
    assert (c.to_cpu() == orc.bit_reverse(v)).all()


def test_batch_inverse(be, orc):
    v = rnd(1, 5000) | 1
    d = be.zeros(5000)
    be.batch_inverse(be.column(v), d)
    assert (d.to_cpu() == orc.batch_inverse_m31(v)).all()
    coords = [rnd(10 + k, 3000) for k in range(4)]
    dst = [be.zeros(3000) for _ in range(4)]
    be.batch_inverse_secure([be.column(c) for c in coords], dst)
    for g, r in zip(dst, orc.batch_inverse_qm31(coords)):
        assert (g.to_cpu() == r).all()


@pytest.mark.parametrize("log", [1, 2, 5, 10, 16, 20])
def test_fold_line(be, orc, tw, log):
    coords = [rnd(300 + 4 * log + k, 1 << log) for k in range(4)]
    alpha = [1, 2, 3, 4] if log % 2 else list(rnd(log, 4))
    got = be.fold_line([be.column(c) for c in coords], log, alpha, tw)
    for g, r in zip(got, orc.fold_line(coords, alpha)):
        assert (g.to_cpu() == r).all()


@pytest.mark.parametrize("log", [3, 4, 7, 12, 18, 21])
def test_fold_circle_into_line(be, orc, tw, log):
    coords = [rnd(400 + 4 * log + k, 1 << log) for k in range(4)]
    dst = [rnd(500 + 4 * log + k, 1 << (log - 1)) for k in range(4)]
    alpha = list(rnd(log + 1, 4))
    gd = [be.column(d) for d in dst]
    be.fold_circle_into_line(gd, [be.column(c) for c in coords], log, alpha, tw)
    for g, r in zip(gd, orc.fold_circle_into_line(dst, coords, alpha)):
        assert (g.to_cpu() == r).all()


def test_fri_fold_to_constant(be, tw):
    # size-independent property at a large size: folding the LDE of a low-degree secure poly ends on a constant layer
    log = 20
    coords = []
    for k in range(4):
        c = np.zeros(1 << (log - 1), dtype=np.uint32)
        c[:] = rnd(600 + k, 1 << (log - 1))
        col = be.column(c)
        coords.append(be.evaluate_polynomials([col], 1, tw)[0])
    alpha = [5, 6, 7, 8]
    line = [be.zeros(1 << (log - 1)) for _ in range(4)]
    be.fold_circle_into_line(line, coords, log, alpha, tw)
    lg = log - 1
    while lg > 1:
        line = be.fold_line(line, lg, alpha, tw)
        lg -= 1
    for l in line:
        v = l.to_cpu()
        assert v[0] == v[1]


@pytest.mark.parametrize("log", [1, 4, 10, 11, 12, 15, 20])
def test_prefix_sum(be, orc, log):
    v = rnd(700 + log, 1 << log)
    c = be.column(v)
    be.inclusive_prefix_sum(c)
    assert (c.to_cpu() == orc.prefix_sum_bitrev(v)).all()


def test_accumulate_and_powers(be, orc):
    a = [rnd(800 + k, 1 << 12) for k in range(4)]
    b = [rnd(810 + k, 1 << 12) for k in range(4)]
    ga = [be.column(x) for x in a]
    be.accumulate(ga, [be.column(x) for x in b])
    for g, x, y in zip(ga, a, b):
        assert (g.to_cpu() == (x.astype(np.uint64) + y) % P).all()
    f = rnd(3, 4)
    assert (be.generate_secure_powers(f, 103) == orc.secure_powers(f, 103)).all()


def test_gen_is_first_and_broadcast(be):
    c = be.gen_is_first(10).to_cpu()
    assert c[0] == 1 and not c[1:].any()
    v = rnd(9, 100)
    assert (be.broadcast16(be.column(v)).to_cpu() == np.repeat(v, 16)).all()


@pytest.mark.parametrize("log,ncols", [(4, 3), (9, 17), (14, 40)])
def test_accumulate_quotients(be, orc, log, ncols):
    cols = [rnd(900 + i, 1 << log) for i in range(ncols)]
    alpha = rnd(1, 4)
    # two batches: every column at point A, the last 4 columns also at point B
    pts = rnd(2, 16)
    bsizes = [ncols, min(4, ncols)]
    ecols = list(range(ncols)) + list(range(ncols - bsizes[1], ncols))
    evals = rnd(3, 4 * len(ecols))
    got = be.accumulate_quotients(log, [be.column(c) for c in cols], alpha, pts, bsizes, ecols, evals)
    ref = orc.accumulate_quotients(log, cols, alpha, pts, bsizes, ecols, evals)
    for g, r in zip(got, ref):
        assert (g.to_cpu() == r).all()
    # no batches -> zero column
    z = be.accumulate_quotients(log, [be.column(c) for c in cols], alpha, [], [], [], [])
    assert not z[0].to_cpu().any()


@pytest.mark.parametrize("log,ncols,nbatch", [(10, 5, 1), (11, 9, 3), (12, 4, 5), (16, 24, 2), (17, 6, 4)])
def test_accumulate_quotients_many_batches(be, orc, pkg, log, ncols, nbatch):
    """Domains of >= 2^10 rows take the table-driven kernel; batch counts around its chunk size (2), row ranges included."""
    import ctypes
    cols = [rnd(1900 + i, 1 << log) for i in range(ncols)]
    alpha = rnd(11, 4)
    pts = rnd(12, 8 * nbatch)
    bsizes = [max(1, ncols - b) for b in range(nbatch)]
    ecols = [c for b in range(nbatch) for c in range(ncols - bsizes[b], ncols)]
    evals = rnd(13, 4 * len(ecols))
    dcols = [be.column(c) for c in cols]
    got = be.accumulate_quotients(log, dcols, alpha, pts, bsizes, ecols, evals)
    ref = orc.accumulate_quotients(log, cols, alpha, pts, bsizes, ecols, evals)
    for g, r in zip(got, ref):
        assert (g.to_cpu() == r).all()
    # a row range that starts and ends off the 512-row table granularity (sharded prover's call)
    off, n = (1 << log) // 4 + 4 * 37, (1 << log) // 2 + 4 * 5
    views = [be.column(c[off:off + n]) for c in cols]
    out = (ctypes.c_void_p * 4)()
    rc, bp, bs = (np.ascontiguousarray(a, dtype=np.uint32) for a in (alpha, pts, bsizes))
    ec, ev = (np.ascontiguousarray(a, dtype=np.uint32) for a in (ecols, evals))
    u32p = ctypes.POINTER(ctypes.c_uint32)
    be._ck(be._lib.sc_accumulate_quotients_range(be._ctx, ctypes.c_uint32(log), ctypes.c_uint64(off), ctypes.c_uint64(n), be._arr(views),
                                                 ctypes.c_uint32(len(views)), rc.ctypes.data_as(u32p), bp.ctypes.data_as(u32p),
                                                 bs.ctypes.data_as(u32p), ec.ctypes.data_as(u32p), ev.ctypes.data_as(u32p),
                                                 ctypes.c_uint32(bs.size), out))
    for k in range(4):
        part = pkg.Column(be, ctypes.c_void_p(out[k])).to_cpu()
        assert (part == ref[k][off:off + n]).all(), f"range coordinate {k}"


def test_grind(be, orc):
    for seed in range(4):
        d = np.random.default_rng(seed).integers(0, 2**32, size=8, dtype=np.uint32)
        assert be.grind(d, 5) == orc.grind(d, 5)
        assert be.grind(d, 12) == orc.grind(d, 12)


def test_column_api(be):
    c = be.zeros(64)
    assert len(c) == 64 and c.at(5) == 0
    c.set(5, 77)
    assert c.at(5) == 77 and c.clone().at(5) == 77


@pytest.mark.parametrize("log", [3, 4, 5, 9, 13, 14, 17, 20, 21])
def test_is_first_coeffs_closed_form(be, orc, tw, log):
    got = be.is_first_coeffs(log, tw).to_cpu()
    c = be.gen_is_first(log)
    be.interpolate_columns([c], tw)
    assert (got == c.to_cpu()).all()
    assert (got == orc.interpolate(np.eye(1, 1 << log, 0, dtype=np.uint32)[0], ROOT_LOG)).all()


def test_error_paths_return_status_not_crash(be, tw, pkg):
    """Precondition violations come back as BackendError (Stwo asserts / panics there); the context stays usable."""
    import ctypes
    a, b = be.column(rnd(1, 1 << 8)), be.column(rnd(2, 1 << 9))
    with pytest.raises(pkg.BackendError):
        be.commit_on_layer(8, None, [a, b])                                   # column length != 2^log_size
    with pytest.raises(pkg.BackendError):
        be.commit_on_layer(8, a, [a])                                         # previous layer of the wrong size
    with pytest.raises(pkg.BackendError):
        be.evaluate_polynomials([be.zeros(1 << (ROOT_LOG + 1))], 1, tw)       # domain exceeds the twiddle tree
    with pytest.raises(pkg.BackendError):
        be.fold_line([a, a, a, b], 8, rnd(3, 4), tw)                          # ragged coordinate columns
    with pytest.raises(pkg.BackendError):
        be.accumulate_quotients(8, [a], rnd(4, 4), rnd(5, 8), [1], [3], rnd(6, 4))   # column index out of range
    with pytest.raises(pkg.BackendError):
        be.accumulate([a, a, a, a], [b, b, b, b])                             # length mismatch
    with pytest.raises(pkg.BackendError):
        be.merkle_commit_repeated([a], 9)                                     # log_repeat out of range
    with pytest.raises(pkg.BackendError):
        be.eval_at_point_repeated([be.zeros(1 << 20)], [12], rnd(7, 8))       # logical size beyond 2^28
    assert be._lib.sc_col_free(be._ctx, None) == 0                            # freeing NULL is a no-op
    assert be._lib.sc_interpolate(be._ctx, None, ctypes.c_uint32(0), tw._h) == 0   # empty batch
    # the context still works
    c = be.column(rnd(8, 1 << 8))
    be.interpolate_columns([c], tw)
    assert len(c.to_cpu()) == 256


@pytest.mark.parametrize("log", [3, 4, 5, 7, 8, 9, 12, 13, 16, 19])
def test_is_first_lde_closed_form(be, orc, tw, log):
    """sc_is_first_lde writes the 2x extension of gen_is_first without a transform (the polynomial is a rank-one product):
    same values as evaluate(interpolate(e_0)) on the device and through the oracle's plain FFT, whole columns and row ranges
    (the ranges a rank of the sharded prover asks for)."""
    e0 = np.eye(1, 1 << log, 0, dtype=np.uint32)[0]
    ref = orc.evaluate(orc.interpolate(e0, ROOT_LOG), 1, ROOT_LOG)
    poly = be.is_first_coeffs(log, tw)
    dev = be.evaluate_polynomials([poly], 1, tw)[0].to_cpu()
    assert (dev == ref).all()
    got = be.is_first_lde(log, 1, tw).to_cpu()
    assert (got == ref).all()
    n = 2 << log
    for parts in (2, 4, 8):
        seg = n // parts
        if seg < 4:
            continue
        for r in (0, parts - 1, parts // 2):
            part = be.is_first_lde(log, 1, tw, r * seg, seg).to_cpu()
            assert (part == ref[r * seg:(r + 1) * seg]).all()
    same = be.is_first_lde(log, 0, tw).to_cpu()   # no blow-up: the indicator itself
    assert (same == e0).all()
