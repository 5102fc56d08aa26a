"""Pure-Python transcription of the reference's VM and of its 13 table builders — an independent statement of the host logic
that produces the inputs of the proving path, checked against csrc/host/{vm,tables}.hpp over random programs.

Follows crates/brainfuck_vm/src/compiler.rs:13-37 and machine.rs:141-238 (compile, execute) and, under
crates/brainfuck_prover/src/components: memory/table.rs:113-160,240-300,306-318; instruction/table.rs:112-150,240-290;
program/table.rs:60-68,111-137; processor/table.rs:109-150,226-262; processor/instructions/table.rs:125-160,286-325;
.../jump/table.rs:113-215,262-295; .../end_of_execution/table.rs:100-109.  Test infrastructure only."""
P = (1 << 31) - 1


def compile_bf(code: bytes):
    ins, stack = [], []
    for ch in code:
        if chr(ch).isspace():
            continue
        ins.append(ch)
        if ch == ord("["):
            ins.append(0)
            stack.append(len(ins) - 1)
        elif ch == ord("]"):
            start = stack.pop()
            ins[start] = len(ins)
            ins.append(start + 1)
    assert not stack
    return ins


class VmError(Exception):
    pass


def execute(program, stdin: bytes, ram_size=30000, max_steps=4000):
    """Returns the trace: one dict per clock cycle plus the final row (ci = ni = 0)."""
    ram = [0] * ram_size
    r = dict(clk=0, ip=0, ci=0, ni=0, mp=0, mv=0, mvi=0)
    trace, pos, out = [], 0, bytearray()
    n = len(program)
    while r["ip"] < n:
        r["ci"] = program[r["ip"]]
        r["ni"] = 0 if r["ip"] == n - 1 else program[r["ip"] + 1]
        trace.append(dict(r))
        if len(trace) > max_steps:
            raise VmError("too long")
        c, jumped = chr(r["ci"]), False
        if c == ">":
            r["mp"] = (r["mp"] + 1) % P
        elif c == "<":
            r["mp"] = (r["mp"] - 1) % P
        elif c == "+":
            ram[r["mp"]] = (ram[r["mp"]] + 1) % P
        elif c == "-":
            ram[r["mp"]] = (ram[r["mp"]] - 1) % P
        elif c == ",":
            if pos >= len(stdin):
                raise VmError("input exhausted")
            ram[r["mp"]] = stdin[pos]
            pos += 1
        elif c == ".":
            out.append(ram[r["mp"]] & 0xFF)
        elif c == "[":
            arg = program[r["ip"] + 1]
            if ram[r["mp"]] == 0:
                r["ip"], jumped = arg, True
            else:
                r["ip"] += 1
        elif c == "]":
            arg = program[r["ip"] + 1]
            if ram[r["mp"]] != 0:
                r["ip"], jumped = arg - 1, True
            else:
                r["ip"] += 1
        else:
            raise VmError("invalid instruction")
        if r["mp"] >= ram_size:
            raise VmError("memory pointer out of range")
        if not jumped:
            r["mv"] = ram[r["mp"]]
            r["mvi"] = pow(r["mv"], P - 2, P) if r["mv"] else 0
        r["clk"] += 1
        r["ip"] += 1
    r["ci"] = r["ni"] = 0
    trace.append(dict(r))
    return trace, bytes(out)


def next_pow2(n):
    return 1 if n == 0 else 1 << (n - 1).bit_length()


def memory_table(regs):
    e = sorted(([r["clk"], r["mp"], r["mv"], 0] for r in regs), key=lambda x: (x[1], x[0]))
    filled, prev = [], e[0]
    for x in e:
        if x[1] == prev[1] and x[0] > prev[0] + 1:
            filled += [[c, prev[1], prev[2], 1] for c in range(prev[0] + 1, x[0])]
        filled.append(x)
        prev = x
    last = filled[-1]
    filled += [[(last[0] + i) % P, last[1], last[2], 1] for i in range(1, next_pow2(len(filled)) - len(filled) + 1)]
    last = filled[-1]
    filled.append([(last[0] + 1) % P, last[1], last[2], 1])
    return [a + b for a, b in zip(filled, filled[1:])]


def program_rows(program):
    return [dict(clk=0, ip=i, ci=c, ni=0 if i == len(program) - 1 else program[i + 1]) for i, c in enumerate(program)]


def instruction_table(regs, program):
    rows = sorted(program_rows(program) + [dict(r) for r in regs], key=lambda r: (r["ip"], r["clk"]))
    e = [[r["ip"], r["ci"], r["ni"], 0] for r in rows]
    last_ip = e[-1][0]
    e += [[last_ip, 0, 0, 1]] * (next_pow2(len(e)) - len(e))
    e.append([e[-1][0], 0, 0, 1])
    return [a + b for a, b in zip(e, e[1:])]


def program_table(program):
    e = [[r["ip"], r["ci"], r["ni"], 0] for r in program_rows(program)]
    return e + [[e[-1][0], 0, 0, 1]] * (next_pow2(len(e)) - len(e))


def entry(r):
    return [r["clk"], r["ip"], r["ci"], r["ni"], r["mp"], r["mv"], r["mvi"], 0]


def dummy(clk, ip):
    return [clk % P, ip, 0, 0, 0, 0, 0, 1]


def processor_table(regs):
    e = [entry(r) for r in regs]
    last = e[-1]
    e += [dummy(last[0] + i, last[1]) for i in range(1, next_pow2(len(e)) - len(e) + 1)]
    e.append(dummy(e[-1][0] + 1, e[-1][1]))
    return [a + [b[0]] for a, b in zip(e, e[1:])]


def paired_entries(regs, opcode):
    e = []
    for a, b in zip(regs, regs[1:]):
        if a["ci"] == opcode:
            e += [entry(a), entry(b)]
    last_clk, last_ip = (e[-1][0], e[-1][1]) if e else (0, 0)
    e += [dummy(last_clk + i, last_ip) for i in range(next_pow2(len(e)) - len(e))]
    pairs = [(e[i], e[i + 1]) for i in range(0, len(e) - 1, 2)]
    if len(e) % 2:
        pairs.append((e[-1], dummy(e[-1][0] + 1, e[-1][1])))
    return pairs


def instruction_sub_table(regs, opcode):
    return [a + [b[1], b[4], b[5]] for a, b in paired_entries(regs, opcode)]          # + next_ip, next_mp, next_mv


def jump_table(regs, opcode):
    rows = []
    for a, b in paired_entries(regs, opcode):
        assert a[7] == b[7]
        rows.append(a[:7] + [b[0], b[1], b[4], b[5], a[7], (1 - a[5] * a[6]) % P])
    return rows


def end_of_execution_table(regs):
    return [entry(r)[:7] for r in regs if r["ci"] == 0]


def build_table(comp, regs, program):
    if comp == 0:
        return memory_table(regs)
    if comp == 1:
        return instruction_table(regs, program)
    if comp == 2:
        return program_table(program)
    if comp == 3:
        return processor_table(regs)
    if comp in (4, 5):
        return jump_table(regs, ord("]" if comp == 4 else "["))
    if comp <= 11:
        return instruction_sub_table(regs, ord(",<-.+>"[comp - 6]))
    return end_of_execution_table(regs)
