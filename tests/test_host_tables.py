"""Host side of the path — VM, compiler and the 13 table builders (csrc/host/{vm,tables}.hpp, shared with the oracle's prover)
— against the values the reference's own unit tests assert.  The expected rows below are transcribed from those tests:

* crates/brainfuck_vm/tests/integration.rs                    program outputs
* components/processor/table.rs:672-875                        processor table of "+>,<[>+.<-]" with input 1
* components/instruction/table.rs:609-742, 745-797             instruction tables of "+>,<[>+.<-]" and "[-]"
* components/processor/instructions/table.rs:653-728           `<` table of "+>,<[>+.<-]"
* components/processor/instructions/jump/table.rs:665-746      `]` table of "++>,<[>+.<-]"
* components/memory/table.rs:713-744                           memory table from three hand-written registers
* components/program/table.rs:357-381, end_of_execution/table.rs:339-371   program and end-of-execution tables
* components/memory/component.rs:211-609                       ten corrupted Memory tables: row and value of the first violation
* SURVEY.md Table S / the reference's component tests           log sizes of the shipped programs

`assert_constraints` (constraint_framework::assert_constraints, used by all 13 component tests of the reference) runs on the
oracle's evaluators over the same tables."""
import ctypes
import os

import numpy as np
import pytest

from oracle_lib import P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROGRAMS = os.path.join(ROOT, "tests", "golden", "programs")
u32p = ctypes.POINTER(ctypes.c_uint32)
MEMORY, INSTRUCTION, PROGRAM, PROCESSOR, JNZ, JZ, INPUT, LEFT, MINUS, OUTPUT, PLUS, RIGHT, EOE = range(13)
PLUS_C, RIGHT_C, READ_C, LEFT_C, JZ_C, JNZ_C, PUT_C, MINUS_C = 43, 62, 44, 60, 91, 93, 46, 45
INV2 = (P + 1) // 2


def load(name):
    return open(os.path.join(PROGRAMS, name), "rb").read()


def vm_summary(orc, code, stdin=b""):
    lib = orc.lib
    lib.orc_vm_summary.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    p = lib.orc_vm_summary(code, stdin, ctypes.c_size_t(len(stdin)))
    if not p:
        raise RuntimeError(lib.orc_last_error().decode())
    s = ctypes.string_at(p).decode()
    lib.orc_free(ctypes.c_void_p(p))
    steps, out, logs, prog, ram = s.split(";")
    ints = lambda t: [int(x) for x in t.split(",") if x]
    return int(steps), bytes(ints(out)), ints(logs), ints(prog), ints(ram)


def table(orc, code, stdin, comp):
    lib = orc.lib
    lib.orc_table_dump.restype = ctypes.c_size_t
    nc = ctypes.c_uint32()
    rows = lib.orc_table_dump(code, stdin, ctypes.c_size_t(len(stdin)), comp, None, ctypes.byref(nc))
    out = np.zeros(rows * nc.value, dtype=np.uint32)
    lib.orc_table_dump(code, stdin, ctypes.c_size_t(len(stdin)), comp, out.ctypes.data_as(u32p), ctypes.byref(nc))
    return out.reshape(rows, nc.value).tolist()


def table_from_registers(orc, regs, program, comp):
    lib = orc.lib
    lib.orc_table_from_registers.restype = ctypes.c_size_t
    r = np.ascontiguousarray(regs, dtype=np.uint32).reshape(-1, 7)
    pr = np.ascontiguousarray(program if len(program) else [0], dtype=np.uint32)
    nc = ctypes.c_uint32()
    rows = lib.orc_table_from_registers(r.ctypes.data_as(u32p), ctypes.c_size_t(len(r)), pr.ctypes.data_as(u32p),
                                        ctypes.c_size_t(len(program)), comp, None, ctypes.byref(nc))
    assert rows, lib.orc_last_error()
    out = np.zeros(rows * nc.value, dtype=np.uint32)
    lib.orc_table_from_registers(r.ctypes.data_as(u32p), ctypes.c_size_t(len(r)), pr.ctypes.data_as(u32p),
                                 ctypes.c_size_t(len(program)), comp, out.ctypes.data_as(u32p), ctypes.byref(nc))
    return out.reshape(rows, nc.value).tolist()


# ------------------------------------------------------------------------------------------------ VM (integration.rs)
@pytest.mark.parametrize("name,stdin,want", [
    ("a-bc.bf", b"a", b"bc"), ("collatz.bf", bytes([0x37, 10]), bytes([0x31, 0x36, 10])), ("hello1.bf", b"", b"Hello World!\n"),
    ("hello2.bf", b"", b"Hello World!\n"), ("hello3.bf", b"", b"Hello, World!\n"), ("hello4.bf", b"", b"Hello World!\n"),
    ("hello_kakarot.bf", b"", b"Hello Kakarot World!\n"), ("fib19.bf", b"", bytes([85]))])
def test_vm_outputs_of_shipped_programs(orc, name, stdin, want):
    steps, out, logs, prog, ram = vm_summary(orc, load(name), stdin)
    assert out == want
    if name == "fib19.bf":
        assert steps == 199246 and ram == [0, 2584, 4181, 0, 0]              # trace rows incl. the final one; README.md:122-126
        assert logs == [24, 22, 11, 22, 19, 11, 4, 20, 19, 4, 20, 20, 4]


def test_component_log_sizes(orc):
    assert vm_summary(orc, load("hello_kakarot.bf"))[2] == [17, 14, 12, 14, 8, 4, 4, 10, 10, 9, 13, 11, 4]
    assert vm_summary(orc, load("collatz.bf"), b"7\n")[2] == [21, 17, 13, 17, 14, 13, 5, 15, 14, 6, 14, 15, 4]


def test_sierpinski_exceeds_log_max_rows(orc):
    """BASELINE.json configs[2]: sierpinski.bf runs (257 750 trace rows) but its Memory table has 26 258 214 rows after the clk
    gaps are filled -> 2^25 rows -> log size 29 > LOG_MAX_ROWS 24 (brainfuck_air/mod.rs:427-428).  The reference
    cannot prove it; the prover reports the component instead of building the 1 GiB table."""
    steps, out, logs, prog, ram = vm_summary(orc, load("sierpinski.bf"))
    assert steps == 257750 and logs == [29, 22, 12, 22, 20, 17, 4, 19, 20, 15, 20, 19, 4]
    assert out.startswith(b" " * 31 + b"*\n") and out.count(b"\n") == 32
    lib = orc.lib
    lib.orc_prove_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    # LOG_MAX_ROWS 12 here only to keep the oracle's preprocessed phase short; the GPU test uses 24
    assert not lib.orc_prove_json(load("sierpinski.bf"), b"", ctypes.c_size_t(0), ctypes.c_uint32(12), 0)
    assert b"component too large: memory" in lib.orc_last_error()


def test_compiler_jump_targets(orc):
    # compiler.rs:13-37: `[` is followed by the index after the matching `]`'s slot, `]` by the index after the `[`'s slot
    assert vm_summary(orc, b"+>,<[>+.<-]", b"\x01")[3] == [43, 62, 44, 60, 91, 12, 62, 43, 46, 60, 45, 93, 6]
    assert vm_summary(orc, b"[-]")[3] == [91, 4, 45, 93, 2]
    assert vm_summary(orc, b" + \n+\t")[3] == [43, 43]                        # whitespace is dropped


def test_vm_errors(orc):
    with pytest.raises(RuntimeError):
        vm_summary(orc, b"+]")                # unbalanced
    with pytest.raises(RuntimeError):
        vm_summary(orc, b",", b"")            # input exhausted
    with pytest.raises(RuntimeError):
        vm_summary(orc, b"<+")                # memory pointer below zero (P - 1 is out of the RAM)


# ------------------------------------------------------------------------------------------------ tables
PROC = [  # clk ip ci ni mp mv mvi  (processor/table.rs:680-807)
    (0, 0, PLUS_C, RIGHT_C, 0, 0, 0), (1, 1, RIGHT_C, READ_C, 0, 1, 1), (2, 2, READ_C, LEFT_C, 1, 0, 0), (3, 3, LEFT_C, JZ_C, 1, 1, 1),
    (4, 4, JZ_C, 12, 0, 1, 1), (5, 6, RIGHT_C, PLUS_C, 0, 1, 1), (6, 7, PLUS_C, PUT_C, 1, 1, 1), (7, 8, PUT_C, LEFT_C, 1, 2, INV2),
    (8, 9, LEFT_C, MINUS_C, 1, 2, INV2), (9, 10, MINUS_C, JNZ_C, 0, 1, 1), (10, 11, JNZ_C, 6, 0, 0, 0), (11, 13, 0, 0, 0, 0, 0)]


def test_processor_table_example_program(orc):
    entries = [e + (0,) for e in PROC] + [(12 + i, 13, 0, 0, 0, 0, 0, 1) for i in range(5)]       # dummies 12..16
    want = [list(entries[i]) + [entries[i + 1][0]] for i in range(16)]
    assert table(orc, b"+>,<[>+.<-]", b"\x01", PROCESSOR) == want


def test_instruction_table_example_program(orc):
    ins = {0: (PLUS_C, RIGHT_C), 1: (RIGHT_C, READ_C), 2: (READ_C, LEFT_C), 3: (LEFT_C, JZ_C), 4: (JZ_C, 12), 5: (12, RIGHT_C),
           6: (RIGHT_C, PLUS_C), 7: (PLUS_C, PUT_C), 8: (PUT_C, LEFT_C), 9: (LEFT_C, MINUS_C), 10: (MINUS_C, JNZ_C), 11: (JNZ_C, 6),
           12: (6, 0), 13: (0, 0)}
    order = [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 13]      # instruction/table.rs:708-735
    entries = [(ip,) + ins[ip] + (0,) for ip in order] + [(13, 0, 0, 1)] * 8                       # 7 pads + the pairing dummy
    want = [list(entries[i]) + list(entries[i + 1]) for i in range(32)]
    assert table(orc, b"+>,<[>+.<-]", b"\x01", INSTRUCTION) == want


def test_instruction_table_unused_instruction(orc):
    e = [(0, JZ_C, 4, 0), (0, JZ_C, 4, 0), (1, 4, MINUS_C, 0), (2, MINUS_C, JNZ_C, 0), (3, JNZ_C, 2, 0), (4, 2, 0, 0), (5, 0, 0, 0),
         (5, 0, 0, 1), (5, 0, 0, 1)]
    assert table(orc, b"[-]", b"", INSTRUCTION) == [list(e[i]) + list(e[i + 1]) for i in range(8)]


def test_left_table_example_program(orc):
    # rows pair (the `<` step, the following step): clk ip ci ni mp mv mvi d | next_ip next_mp next_mv
    want = [[3, 3, LEFT_C, JZ_C, 1, 1, 1, 0, 4, 0, 1], [8, 9, LEFT_C, MINUS_C, 1, 2, INV2, 0, 10, 0, 1]]
    assert table(orc, b"+>,<[>+.<-]", b"\x01", LEFT) == want


def test_jump_if_not_zero_table_example_program(orc):
    # clk ip ci ni mp mv mvi | next_clk next_ip next_mp next_mv | d is_mv_zero   (jump/table.rs:688-727)
    want = [[11, 12, JNZ_C, 7, 0, 1, 1, 12, 7, 0, 1, 0, 0], [17, 12, JNZ_C, 7, 0, 0, 0, 18, 14, 0, 0, 0, 1]]
    assert table(orc, b"++>,<[>+.<-]", b"\x01", JNZ) == want


def test_program_table_example(orc):
    # program/table.rs:357-381: "+>-" -> rows (ip, ci, ni, d) padded with dummy(last ip)
    assert table(orc, b"+>-", b"", PROGRAM) == [[0, PLUS_C, RIGHT_C, 0], [1, RIGHT_C, MINUS_C, 0], [2, MINUS_C, 0, 0], [2, 0, 0, 1]]


def test_end_of_execution_table_example_program(orc):
    # end_of_execution/table.rs:339-371: the single row with ci == 0 of "+>,<[>+.<-]" (clk 11, ip 13, everything else 0)
    assert table(orc, b"+>,<[>+.<-]", b"\x01", EOE) == [[11, 13, 0, 0, 0, 0, 0]]


def test_empty_instruction_tables_are_one_dummy_row(orc):
    assert table(orc, b"+", b"", INPUT) == [[0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0]]
    assert table(orc, b"+", b"", JZ) == [[0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 1]]
    assert table(orc, b"+", b"", EOE) == [[1, 1, 0, 0, 0, 1, 1]]


def test_memory_table_from_registers(orc):
    regs = [(0, 0, 0, 0, 0, 0, 0), (1, 0, 0, 0, 1, 0, 0), (5, 0, 0, 0, 1, 1, 1)]                  # memory/table.rs:715-724, in clk order
    e = [(0, 0, 0, 0), (1, 1, 0, 0), (2, 1, 0, 1), (3, 1, 0, 1), (4, 1, 0, 1), (5, 1, 1, 0), (6, 1, 1, 1), (7, 1, 1, 1), (8, 1, 1, 1)]
    assert table_from_registers(orc, regs, [], MEMORY) == [list(e[i]) + list(e[i + 1]) for i in range(8)]


def test_memory_table_sorts_by_address_then_clock(orc):
    # "+>+<-": cell 0 is touched at clk 0,1 and 4,5; cell 1 at clk 2,3 -> (mp, clk) order with the clk gap 2..3 of cell 0 filled
    rows = table(orc, b"+>+<-", b"", MEMORY)
    got = [tuple(r[:4]) for r in rows]
    assert got[:6] == [(0, 0, 0, 0), (1, 0, 1, 0), (2, 0, 1, 1), (3, 0, 1, 1), (4, 0, 1, 0), (5, 0, 0, 0)]
    assert got[6:] == [(2, 1, 0, 0), (3, 1, 1, 0)]
    assert all(rows[i][4:] == rows[i + 1][:4] for i in range(len(rows) - 1))


# ------------------------------------------------------------------------------------------------ constraints
@pytest.mark.parametrize("code,stdin", [(b"+>,<[>+.<-]", b"\x01"), (b"+++>++<[->+<]>.", b""), (b"++[>+<-]>[-]<", b""), (b"+", b""),
                                        (None, b"")])
@pytest.mark.parametrize("dummy", [1, 0])
def test_assert_constraints_all_components(orc, code, stdin, dummy):
    if code is None:
        code = load("hello_kakarot.bf")
    lib = orc.lib
    lib.orc_assert_constraints.restype = ctypes.c_void_p
    p = lib.orc_assert_constraints(code, stdin, ctypes.c_size_t(len(stdin)), dummy)
    if p:
        msg = ctypes.string_at(p).decode()
        lib.orc_free(ctypes.c_void_p(p))
        pytest.fail(msg)


# ------------------------------------------------------------------------------------------------ negative vectors
# components/memory/component.rs:211-609: ten corrupted Memory tables with the row and value of the first constraint that
# `assert_constraints` reports (`#[should_panic = "... row: R\n  left: (v + 0i) + (0 + 0i)u ..."]`).  Tables are written out
# as the reference builds them: entries sorted by (mp, clk), each row paired with the next entry, one more dummy at the end.
# columns: clk mp mv d | next_clk next_mp next_mv next_d;  elements: 0 = drawn from a fresh channel, 2 = LookupElements::dummy()
def _mem(entries, last_dummy):
    e = list(entries) + [last_dummy]
    return [list(e[i]) + list(e[i + 1]) for i in range(len(entries))]


NEGATIVE = {
    "boundary_clk": (_mem([(1, 0, 0, 0)], (2, 0, 0, 1)), 0, 0, 1),
    "boundary_mp": (_mem([(0, 1, 0, 0)], (1, 1, 0, 1)), 0, 0, 1),
    "boundary_mv": (_mem([(0, 0, 1, 0)], (1, 0, 1, 1)), 0, 0, 1),
    "boundary_d": (_mem([(0, 0, 0, 1)], (1, 0, 0, 1)), 0, 0, 1),
    "transition_mp_increase": (_mem([(0, 0, 0, 0), (0, 2, 0, 0)], (1, 2, 0, 1)), 2, 0, 2),
    "transition_clk_increase": (_mem([(0, 0, 0, 0), (0, 0, 0, 0)], (1, 0, 0, 1)), 0, 0, 1),
    "transition_mp_increase_next_mv": (_mem([(0, 0, 0, 0), (0, 1, 1, 0)], (1, 1, 1, 1)), 0, 0, 1),
}
_t = _mem([(0, 0, 0, 0), (0, 1, 0, 0)], (1, 1, 0, 1))
_a = [list(r) for r in _t]; _a[0][7] = 2
NEGATIVE["transition_next_dummy"] = (_a, 0, 0, 2)
_b = [list(r) for r in _t]; _b[1][3] = 1; _b[1][5] = 2
NEGATIVE["transition_d_mp"] = (_b, 0, 1, 1)
_c = [list(r) for r in _t]; _c[1][3] = 1; _c[1][6] = 1
NEGATIVE["transition_d_mv"] = (_c, 0, 1, 1)


@pytest.mark.parametrize("name", sorted(NEGATIVE))
def test_memory_component_negative_vectors(orc, name):
    rows, elements, want_row, want_value = NEGATIVE[name]
    lib = orc.lib
    lib.orc_assert_table.restype = ctypes.c_void_p
    flat = np.ascontiguousarray(rows, dtype=np.uint32)
    p = lib.orc_assert_table(MEMORY, flat.ctypes.data_as(u32p), ctypes.c_size_t(len(rows)), ctypes.c_size_t(8), elements)
    assert p, f"{name}: the corrupted table passed"
    msg = ctypes.string_at(p).decode()
    lib.orc_free(ctypes.c_void_p(p))
    assert f" row {want_row} left ({want_value} + 0i) + (0 + 0i)u" in msg, msg


def test_memory_component_valid_table_passes(orc):
    lib = orc.lib
    lib.orc_assert_table.restype = ctypes.c_void_p
    flat = np.ascontiguousarray(_t, dtype=np.uint32)
    assert not lib.orc_assert_table(MEMORY, flat.ctypes.data_as(u32p), ctypes.c_size_t(2), ctypes.c_size_t(8), 0)


# ------------------------------------------------------------------------------------------------ random programs
def random_program(rng, depth=0):
    """Half of the programs are unstructured (balanced brackets only); the other half is built from counted loops
    `[` > body < `-]` whose body restores the pointer, so that they terminate and run for hundreds of steps."""
    if depth == 0 and rng.integers(0, 2) == 0:
        out, d = [], 0
        for _ in range(int(rng.integers(1, 40))):
            c = "+-<>.,[]"[int(rng.integers(0, 8))]
            if c == "]" and d == 0:
                c = "+"
            d += (c == "[") - (c == "]")
            out.append(c)
        return ("".join(out) + "]" * d).encode()
    out = []
    for _ in range(int(rng.integers(1, 5))):
        kind = int(rng.integers(0, 4))
        if kind == 0:
            out.append("+" * int(rng.integers(1, 6)))
        elif kind == 1:
            out.append(",." if rng.integers(0, 2) else ".")
        elif kind == 2 and depth < 2:
            k = int(rng.integers(1, 3))
            out.append("+" * int(rng.integers(0, 4)) + "[" + ">" * k + random_program(rng, depth + 1).decode() + "<" * k + "-]")
        else:
            out.append(">" + "+" * int(rng.integers(0, 3)) + "<" if rng.integers(0, 2) else "-" * int(rng.integers(1, 3)) + "+" * 3)
    return "".join(out).encode()


def test_vm_and_tables_match_the_python_transcription_on_random_programs(orc):
    """tests/host_model.py restates the reference's VM and its 13 table builders; the C++ host code must produce the same
    output, step count and table rows on random programs (including empty sub-tables, single-step programs, loops that never
    run and loops entered with a non-zero cell)."""
    import host_model as H
    rng = np.random.default_rng(0xB7A1)
    done = longest = 0
    kinds = set()
    for _ in range(400):
        code = random_program(rng)
        stdin = bytes(int(x) for x in rng.integers(0, 6, size=64))
        prog = H.compile_bf(code)
        try:
            regs, out = H.execute(prog, stdin, max_steps=3000)
        except (H.VmError, IndexError):
            continue
        if any(r["mp"] > 1000 for r in regs):
            continue
        steps, got_out, logs, got_prog, _ = vm_summary(orc, code, stdin)
        assert (steps, got_out, got_prog) == (len(regs), out, prog), code
        for comp in range(13):
            want = H.build_table(comp, regs, prog)
            assert table(orc, code, stdin, comp) == want, (code, comp)
            assert logs[comp] == (len(want) - 1).bit_length() + 4
        kinds |= {chr(r["ci"]) for r in regs[:-1]}
        longest = max(longest, len(regs))
        done += 1
        if done >= 60:
            break
    assert done >= 40 and kinds == set("+-<>.,[]") and longest >= 500
