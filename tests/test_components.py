"""LogUp interaction traces of the 13 components against an independent pure-Python model (tests/logup_model.py) of the
reference's seven `interaction_trace_evaluation` functions and Stwo's LogupTraceGenerator.  The reference's own unit tests
make the same comparison against Stwo's generator (e.g. components/memory/table.rs:811-878, processor/table.rs:1070-…).
CPU: the oracle.  GPU: sc_logup_generate through the C ABI, on lane-compact and on full columns."""
import ctypes

import numpy as np
import pytest

import logup_model as M
from test_host_tables import load, table, u32p

PROGRAMS = [(b"+>,<[>+.<-]", b"\x01"), (b"++[>+<-]>[-]<", b""), (load("a-bc.bf"), b"a")]
DUMMY = [1, 0, 0, 0] * 8 * 3          # LookupElements::dummy(): z = 1 and every alpha power = 1, for the three relations


def random_elements(seed):
    rng = np.random.default_rng(seed)
    return [int(x) for x in rng.integers(1, M.P, size=96)]


def oracle_logup(orc, comp, rows, el):
    lib = orc.lib
    lib.orc_logup_table.restype = ctypes.c_size_t
    a = np.ascontiguousarray(rows, dtype=np.uint32)
    n_rows, n_cols = a.shape
    n_out = 4 * len(M.FRACTIONS[comp])
    out = np.zeros((n_out, 16 * n_rows), dtype=np.uint32)
    claimed = np.zeros(4, dtype=np.uint32)
    e = np.ascontiguousarray(el, dtype=np.uint32)
    got = lib.orc_logup_table(comp, a.ctypes.data_as(u32p), ctypes.c_size_t(n_rows), ctypes.c_size_t(n_cols), e.ctypes.data_as(u32p),
                              out.ctypes.data_as(u32p), claimed.ctypes.data_as(u32p))
    assert got == n_out, lib.orc_last_error()
    return out.tolist(), tuple(int(x) for x in claimed)


@pytest.mark.parametrize("prog", range(len(PROGRAMS)))
def test_oracle_logup_matches_the_python_model(orc, prog):
    code, stdin = PROGRAMS[prog]
    el = random_elements(0x10C0 + prog)
    total = (0, 0, 0, 0)
    for comp in range(13):
        rows = table(orc, code, stdin, comp)
        want_cols, want_sum = M.logup_columns(comp, rows, el)
        got_cols, got_sum = oracle_logup(orc, comp, rows, el)
        assert got_sum == want_sum, comp
        assert got_cols == want_cols, comp
        total = M.q_add(total, want_sum)
    # lookup_sum_valid (brainfuck_air/mod.rs:187-227): over a real execution every provided tuple is consumed
    assert total == (0, 0, 0, 0)


def test_memory_dummy_entries_do_not_change_the_claimed_sum(orc):
    """components/memory/table.rs:886-929: the same real entries with and without clk-gap / padding dummies."""
    with_dummies = [[0, 43, 91, 0, 1, 43, 91, 1], [1, 43, 91, 1, 2, 91, 9, 0], [2, 91, 9, 0, 2, 91, 9, 1], [2, 91, 9, 1, 3, 91, 9, 1]]
    real_only = [[0, 43, 91, 0, 2, 91, 9, 0], [2, 91, 9, 0, 3, 91, 9, 1]]
    a, b = oracle_logup(orc, 0, with_dummies, DUMMY)[1], oracle_logup(orc, 0, real_only, DUMMY)[1]
    assert a == b == M.logup_columns(0, real_only, DUMMY)[1]
    # by hand: -(1/(0+43+91-1) + 1/(2+91+9-1)) in M31, once per SIMD lane (a table row fills all 16 lanes of its vec_row)
    inv = lambda v: pow(v, M.P - 2, M.P)
    assert a == ((-16 * (inv(133) + inv(101))) % M.P, 0, 0, 0)


def test_memory_fixture_of_the_reference(orc):
    """components/memory/table.rs:811-878: real, dummy, real, dummy with LookupElements::dummy(): numerators -1, 0, -1, 0."""
    rows = [[0, 0, 0, 0, 1, 1, 0, 1], [1, 1, 0, 1, 2, 1, 0, 0], [2, 1, 0, 0, 3, 1, 0, 1], [3, 1, 0, 1, 4, 1, 0, 1]]
    el = [5, 6, 7, 8] + [1, 0, 0, 0] * 7 + DUMMY[32:]     # dummy() itself makes row 0's denominator 0 + 0 + 0 - 1 fine, row 1's zero
    cols, s = oracle_logup(orc, 0, rows, el)
    want_cols, want_s = M.logup_columns(0, rows, el)
    assert cols == want_cols and s == want_s
    order = M.coset_order_storage_indices(6)
    firsts = [tuple(cols[k][order[0]] for k in range(4)), tuple(cols[k][order[-1]] for k in range(4))]
    assert firsts[1] == s and cols[0][1] == s[0]          # claimed sum = last coset point = storage index 1


@pytest.mark.gpu
@pytest.mark.parametrize("prog", range(len(PROGRAMS)))
def test_cuda_logup_matches_the_python_model(be, orc, prog):
    code, stdin = PROGRAMS[prog]
    el = random_elements(0x10C0 + prog)
    for comp in range(13):
        rows = table(orc, code, stdin, comp)
        want_cols, want_sum = M.logup_columns(comp, rows, el)
        a = np.ascontiguousarray(rows, dtype=np.uint32)
        for log_repeat in (4, 0):
            host = [a[:, c] if log_repeat else np.repeat(a[:, c], 16) for c in range(a.shape[1])]
            cols = [be.column(h) for h in host]
            out, claimed = be.logup_generate(comp, cols, el, log_repeat)
            assert tuple(int(x) for x in claimed) == want_sum, (comp, log_repeat)
            assert [o.to_cpu().tolist() for o in out] == want_cols, (comp, log_repeat)
            for c in cols + out:
                c.free()


# ------------------------------------------------------------------------------------------------ constraint quotients
N_CONSTRAINTS = [12, 11, 5, 10, 9, 9, 7, 7, 8, 8, 8, 7, 2]     # per component, SURVEY.md §2


def oracle_constraints(orc, comp, rows, el, coeffs):
    lib = orc.lib
    lib.orc_constraints_table.restype = ctypes.c_size_t
    a = np.ascontiguousarray(rows, dtype=np.uint32)
    n_rows, n_cols = a.shape
    out = np.zeros((4, 32 * n_rows), dtype=np.uint32)
    claimed = np.zeros(4, dtype=np.uint32)
    e, cf = np.ascontiguousarray(el, dtype=np.uint32), np.ascontiguousarray(coeffs, dtype=np.uint32)
    got = lib.orc_constraints_table(comp, a.ctypes.data_as(u32p), ctypes.c_size_t(n_rows), ctypes.c_size_t(n_cols), e.ctypes.data_as(u32p),
                                    cf.ctypes.data_as(u32p), out.ctypes.data_as(u32p), claimed.ctypes.data_as(u32p))
    assert got == 4, lib.orc_last_error()
    return out, claimed


def random_coeffs(comp, seed):
    return np.random.default_rng(seed).integers(1, M.P, size=(N_CONSTRAINTS[comp], 4)).astype(np.uint32)


@pytest.mark.parametrize("prog", range(len(PROGRAMS)))
def test_constraint_quotients_are_low_degree(orc, prog):
    """Σ_k coeff_k · C_k / Z over CanonicCoset(log_size + 1), N = 2^log_size: a degree-2 constraint over trace polynomials of
    total degree <= N/2 has degree <= N, its quotient by the coset vanishing polynomial (degree N/2) degree <= N/2 — a space of
    dimension N + 1, which is the first N + 1 functions of the FFT basis on the 2N-point domain.  So when every constraint
    holds, coefficients N+1 .. 2N-1 of the interpolated quotient vanish (the one at N is why Stwo commits the composition
    polynomial at log_size + 1).  Any wrong constraint, mask offset, vanishing denominator or claimed sum leaves a
    full-degree function instead.  Checked for Memory, Instruction and Program, whose constraints are all of degree <= 2;
    the Processor family has degree-3 constraints (e.g. mv * (mv * mvi - 1)) under the same log_size + 1 bound
    (`max_constraint_log_degree_bound`, e.g. processor/component.rs), so its quotient legitimately fills the domain."""
    code, stdin = PROGRAMS[prog]
    el = random_elements(0xC0DE + prog)
    for comp in (0, 1, 2):
        rows = table(orc, code, stdin, comp)
        acc, _ = oracle_constraints(orc, comp, rows, el, random_coeffs(comp, comp))
        log = (len(rows) * 32).bit_length() - 1
        for k in range(4):
            coeffs = orc.interpolate(acc[k], log)
            assert not coeffs[len(coeffs) // 2 + 1:].any(), (comp, k)
            assert coeffs[:len(coeffs) // 2].any(), (comp, k)
    # a corrupted table does not pass: a dummy flag that is not a bit (memory/component.rs: d * (d - 1) = 0)
    rows = table(orc, code, stdin, 0)
    rows[0][3] = 2
    acc, _ = oracle_constraints(orc, 0, rows, el, random_coeffs(0, 0))
    log = (len(rows) * 32).bit_length() - 1
    assert any(orc.interpolate(acc[k], log)[16 * len(rows) + 1:].any() for k in range(4))


@pytest.mark.gpu
@pytest.mark.parametrize("prog", range(len(PROGRAMS)))
def test_cuda_eval_constraints_matches_oracle(be, orc, prog):
    """sc_eval_constraints through the C ABI, per component, on LDEs made by the C ABI's own transforms, against the oracle."""
    code, stdin = PROGRAMS[prog]
    el = random_elements(0xC0DE + prog)
    tw = be.precompute_twiddles(14)
    for comp in range(13):
        rows = table(orc, code, stdin, comp)
        coeffs = random_coeffs(comp, comp)
        want, want_sum = oracle_constraints(orc, comp, rows, el, coeffs)
        a = np.ascontiguousarray(rows, dtype=np.uint32)
        log_size = (len(rows) * 16).bit_length() - 1
        full = [be.column(np.repeat(a[:, c], 16)) for c in range(a.shape[1])]
        inter, claimed = be.logup_generate(comp, full, el, 0)
        assert claimed.tolist() == want_sum.tolist()
        isf = be.gen_is_first(log_size)
        be.interpolate_columns(full + inter + [isf], tw)
        main_lde, inter_lde = be.evaluate_polynomials(full, 1, tw), be.evaluate_polynomials(inter, 1, tw)
        isf_lde = be.evaluate_polynomials([isf], 1, tw)[0]
        acc = [be.zeros(2 << log_size) for _ in range(4)]
        be.eval_constraints(comp, log_size, main_lde, inter_lde, isf_lde, el, claimed, coeffs, acc)
        got = np.stack([c.to_cpu() for c in acc])
        assert np.array_equal(got, want), comp
        # accumulation: a second call adds the same values again
        be.eval_constraints(comp, log_size, main_lde, inter_lde, isf_lde, el, claimed, coeffs, acc)
        twice = (2 * want.astype(np.uint64) % M.P).astype(np.uint32)
        assert np.array_equal(np.stack([c.to_cpu() for c in acc]), twice), comp
        for c in full + inter + [isf, isf_lde] + main_lde + inter_lde + acc:
            c.free()


# ------------------------------------------------------------------------------------------------ the AIR itself
def oracle_constraint_values(orc, comp, rows, el, nat):
    lib = orc.lib
    lib.orc_constraint_values.restype = ctypes.c_size_t
    a = np.ascontiguousarray(rows, dtype=np.uint32)
    e = np.ascontiguousarray(el, dtype=np.uint32)
    out = np.zeros((N_CONSTRAINTS[comp], 4), dtype=np.uint32)
    got = lib.orc_constraint_values(comp, a.ctypes.data_as(u32p), ctypes.c_size_t(a.shape[0]), ctypes.c_size_t(a.shape[1]),
                                    e.ctypes.data_as(u32p), ctypes.c_size_t(nat), out.ctypes.data_as(u32p))
    assert got == N_CONSTRAINTS[comp], lib.orc_last_error()
    return [tuple(int(x) for x in row) for row in out]


N_MAIN = [8, 8, 4, 9, 13, 13, 11, 11, 11, 11, 11, 11, 7]


@pytest.mark.parametrize("comp", range(13))
def test_air_matches_the_python_transcription_of_the_reference(orc, comp):
    """Every constraint of every component, in the reference's order, on RANDOM tables (which violate the AIR, so that each
    value depends on the exact formula), at the first row (is_first = 1), the last row and rows in between.  This is the
    only check of csrc/host/air.hpp that does not go through air.hpp: the kernels and the oracle both instantiate it."""
    import air_model as A
    rng = np.random.default_rng(0xA1B + comp)
    el = random_elements(0xA1B0 + comp)
    for n_rows in (1, 4):
        rows = [[int(x) for x in rng.integers(0, M.P, size=N_MAIN[comp])] for _ in range(n_rows)]
        n = 16 * n_rows
        for nat in sorted({0, 1, n // 2 - 1, n // 2, n - 1, int(rng.integers(0, n))}):
            want = A.constraint_values(comp, rows, el, nat)
            assert len(want) == N_CONSTRAINTS[comp]
            assert oracle_constraint_values(orc, comp, rows, el, nat) == want, (n_rows, nat)


def test_honest_tables_satisfy_the_python_air(orc):
    """and on the tables of a real execution the transcription evaluates to zero everywhere (assert_constraints, in Python)"""
    import air_model as A
    code, stdin = PROGRAMS[0]
    el = random_elements(7)
    for comp in range(13):
        rows = table(orc, code, stdin, comp)
        if len(rows) > 8:
            continue                      # Python speed: the small tables are enough here, the oracle covers the rest
        for nat in range(16 * len(rows)):
            assert all(v == (0, 0, 0, 0) for v in A.constraint_values(comp, rows, el, nat)), (comp, nat)


def test_python_models_are_consistent_with_each_other():
    """VM -> tables -> LogUp -> AIR entirely in Python (tests/host_model.py, logup_model.py, air_model.py), no C++ involved: on
    real executions every constraint of every component vanishes on every row and the claimed sums cancel
    (`lookup_sum_valid`, brainfuck_air/mod.rs:187-227).  The three transcriptions were written from different files of the
    reference; that they fit together is the check that they were transcribed correctly."""
    import air_model as A
    import host_model as H
    el = random_elements(0x5E1F)
    for code, stdin in ((b"+>,<[>+.<-]", b"\x01"), (b"++[>+<-]>[-]<", b""), (b",[.-]+[[-]>+<]>.", b"\x03"), (b"+", b"")):
        prog = H.compile_bf(code)
        regs, _ = H.execute(prog, stdin)
        total = (0, 0, 0, 0)
        for comp in range(13):
            bad, s = A.first_violation(comp, H.build_table(comp, regs, prog), el)
            assert bad is None, (code, comp, bad)
            total = M.q_add(total, s)
        assert total == (0, 0, 0, 0), code
    # and a forged execution does not pass: the VM claims a cell went 1 -> 3 on a single `+`
    prog = H.compile_bf(b"++")
    regs, _ = H.execute(prog, b"")
    regs[2]["mv"], regs[2]["mvi"] = 3, pow(3, M.P - 2, M.P)
    assert A.first_violation(10, H.build_table(10, regs, prog), el)[0] is not None          # the `+` component objects


def test_air_transcription_against_the_reference_source_text():
    """Where the reference tree exists: every `eval.add_constraint(<expr>)` and `eval.add_to_relation(..)` of the 13 `evaluate()`
    bodies, translated mechanically from the Rust source (tests/ref_air_parser.py) and evaluated on random rows, equals the hand
    transcription of tests/air_model.py — which the oracle and, through proof equality, the CUDA kernels equal in turn."""
    import air_model as A
    import ref_air_parser as R
    if not R.available():
        pytest.skip("reference tree not available")
    rng = np.random.default_rng(0x50C)
    for comp in range(13):
        for trial in range(8):
            row = [int(x) for x in rng.integers(0, M.P, size=N_MAIN[comp])]
            if trial % 2:
                row = [v % 3 for v in row]                  # small values: products that vanish, flags that are bits
            for is_first in (0, 1):
                got_main, got_rel = R.evaluate(comp, row, is_first)
                want_main, want_rel = A.main_and_relations(comp, row, is_first)
                assert [(v, 0, 0, 0) for v in got_main] == want_main, (comp, row)
                assert got_rel == [(n % M.P, r, v) for n, r, v in want_rel], (comp, row)
        assert len(got_main) + len(got_rel) == N_CONSTRAINTS[comp]


def test_logup_fractions_against_the_reference_source_text():
    """Where the reference tree exists: numerator, dummy-flag column, relation and value columns of every
    `col_gen.write_frac(..)` in the seven `interaction_trace_evaluation` functions, read from the Rust source
    (tests/ref_air_parser.py), equal the table tests/logup_model.py was written from."""
    import ref_air_parser as R
    if not R.available():
        pytest.skip("reference tree not available")
    for comp in range(13):
        want = [((1, 0) if sign > 0 else (-1, 0) if d is not None else (-1, -1)) + (d, rel, cols) for sign, d, rel, cols in M.FRACTIONS[comp]]
        assert [tuple(f) for f in R.parse_fractions(comp)] == want, comp
