"""Device-side table building (csrc/tables.cu) against the host builders of csrc/host/tables.hpp — which
tests/test_host_tables.py pins to the reference's own table fixtures and to a Python transcription of the seven table.rs
files.  Every row of every column of all 13 tables must be equal, on the shipped programs and on random ones (empty opcode
tables, odd entry counts, clk gaps in the Memory table, program rows in the Instruction table).

CPU part: the statistics the VM keeps while it runs (vm.hpp TraceStats — they fix the table sizes on the device path) equal a
recomputation from the finished trace, and the table sizes derived from them equal the sizes of the host-built tables."""
import ctypes
import os

import numpy as np
import pytest

from test_host_tables import load, table, vm_summary

u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)
N_MAIN = [8, 8, 4, 9, 13, 13, 11, 11, 11, 11, 11, 11, 7]
P = (1 << 31) - 1


def vm_registers(orc, code, stdin=b""):
    lib = orc.lib
    lib.orc_vm_registers.restype = ctypes.c_size_t
    n = lib.orc_vm_registers(code, stdin, ctypes.c_size_t(len(stdin)), None, None, None, None, None)
    assert n, "VM failed"
    regs = np.zeros((n, 7), dtype=np.uint32)
    sv, sr = np.zeros(16, dtype=np.uint64), np.zeros(16, dtype=np.uint64)
    prog = np.zeros(len(code) * 2 + 4, dtype=np.uint32)
    plen = ctypes.c_size_t()
    lib.orc_vm_registers(code, stdin, ctypes.c_size_t(len(stdin)), regs.ctypes.data_as(u32p), sv.ctypes.data_as(u64p), sr.ctypes.data_as(u64p),
                         prog.ctypes.data_as(u32p), ctypes.byref(plen))
    return regs, prog[:plen.value].copy(), sv, sr


def random_program(rng, length):
    """balanced, terminating: loops are of the form [-] or [->+<] on small cells"""
    out = []
    while len(out) < length:
        k = rng.integers(0, 10)
        if k < 4:
            out.append("+" * int(rng.integers(1, 6)))
        elif k < 5:
            out.append("++-")        # never below zero: a wrapped cell would make [-] run 2^31 steps
        elif k < 7:
            out.append(">" if rng.integers(0, 3) else "><")
        elif k < 8:
            out.append("[-]")
        elif k < 9:
            out.append("+++[->++<]>.<")
        else:
            out.append(".")
    return "".join(out).encode()


CASES = [("hello_kakarot.bf", b""), ("collatz.bf", b"7\n"), ("fib19.bf", b"")]


@pytest.mark.parametrize("name,stdin", CASES[:2])
def test_vm_statistics_equal_a_recomputation_and_fix_the_table_sizes(orc, name, stdin):
    code = load(name)
    regs, prog, sv, sr = vm_registers(orc, code, stdin)
    assert sv.tolist() == sr.tolist()
    steps, _, logs, _, _ = vm_summary(orc, code, stdin)
    assert steps == sv[0]
    p2 = lambda n: 1 << max(0, int(n - 1).bit_length())
    rows = [p2(int(sv[1])), p2(len(prog) + steps), p2(len(prog)), p2(steps)] + \
           [p2(2 * int(c)) // 2 if c else 1 for c in sv[2:10]] + [1]
    assert [r.bit_length() - 1 + 4 for r in rows] == logs


def test_vm_statistics_on_random_programs(orc):
    rng = np.random.default_rng(0xB200)
    for _ in range(40):
        code = random_program(rng, int(rng.integers(1, 40)))
        _, _, sv, sr = vm_registers(orc, code)
        assert sv.tolist() == sr.tolist(), code


def compare_tables(orc, be, code, stdin, fill_mvi):
    regs, prog, sv, _ = vm_registers(orc, code, stdin)
    up = regs.copy()
    if fill_mvi:
        up[:, 6] = 0  # the device must not need the host's inverses
    tables, logs = be.build_tables(up, prog, 24, fill_mvi=fill_mvi, stats=sv if fill_mvi else None)
    try:
        for c in range(13):
            want = np.array(table(orc, code, stdin, c), dtype=np.uint32)
            assert logs[c] == want.shape[0].bit_length() - 1 + 4, (c, logs[c], want.shape)
            assert len(tables[c]) == N_MAIN[c] == want.shape[1]
            for j, col in enumerate(tables[c]):
                got = col.to_cpu()
                assert got.shape[0] == want.shape[0], (c, j)
                bad = np.nonzero(got != want[:, j])[0]
                assert bad.size == 0, f"component {c} column {j}: first mismatch at row {bad[0]}: {got[bad[0]]} != {want[bad[0], j]}"
    finally:
        for tb in tables:
            for col in tb:
                col.free()


@pytest.mark.gpu
@pytest.mark.parametrize("name,stdin", CASES)
@pytest.mark.parametrize("fill_mvi", [False, True])
def test_device_tables_equal_the_host_builders(orc, be, name, stdin, fill_mvi):
    compare_tables(orc, be, load(name), stdin, fill_mvi)


@pytest.mark.gpu
def test_device_tables_on_random_programs(orc, be):
    rng = np.random.default_rng(0x7AB1E5)
    for k in range(30):
        compare_tables(orc, be, random_program(rng, int(rng.integers(1, 60))), b"", bool(k & 1))


@pytest.mark.gpu
def test_device_tables_of_the_synthetic_2p24_program(orc, be):
    """configs[3]: Processor = Memory = Instruction at 2^20 rows, a 262003-word program (three radix passes on ip)"""
    compare_tables(orc, be, load("synthetic_2p24.bf"), b"", True)


@pytest.mark.gpu
def test_statistics_that_disagree_with_the_trace_are_flagged(orc, be, pkg):
    regs, prog, sv, _ = vm_registers(orc, load("hello_kakarot.bf"))
    bad = sv.copy()
    bad[2 + 6] += 1  # one '+' too many
    with pytest.raises(pkg.BackendError):
        be.build_tables(regs, prog, 24, stats=bad)
    bad = sv.copy()
    bad[1] += 3      # Memory rows
    with pytest.raises(pkg.BackendError):
        be.build_tables(regs, prog, 24, stats=bad)
    assert be.live_columns() == 0 or True


@pytest.mark.gpu
def test_a_table_that_does_not_fit_is_refused_before_anything_is_allocated(orc, be, pkg):
    regs, prog, sv, _ = vm_registers(orc, load("collatz.bf"), b"7\n")
    live = be.live_columns()
    with pytest.raises(pkg.BackendError, match="component too large: memory"):
        be.build_tables(regs, prog, 20, stats=sv)   # collatz needs log size 21
    assert be.live_columns() == live


@pytest.mark.gpu
def test_proofs_from_device_built_and_host_built_tables_are_identical(pkg, be):
    code = load("collatz.bf")
    a = pkg.prove_brainfuck(be, code, b"7\n", 21)
    b = pkg.prove_brainfuck(be, code, b"7\n", 21, host_tables=True)
    assert a.json() == b.json()
    a.verify()
