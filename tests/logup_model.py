"""Pure-Python model of the LogUp interaction traces — a third, independent statement of what the reference's seven
`interaction_trace_evaluation` functions compute, used to check both the CPU oracle and the CUDA kernels at small sizes.

Follows, per component, the fractions the reference writes (file:line under crates/brainfuck_prover/src/components):
  memory/table.rs:485-518, instruction/table.rs:456-491, program/table.rs:233-267, processor/table.rs:456-533,
  processor/instructions/table.rs:466-507, .../jump/table.rs:436-477, .../end_of_execution/table.rs:220-257
and Stwo's `LogupTraceGenerator` as those functions drive it (constraint_framework/logup.rs @ 31e8dbc):
  write_frac(row, num, denom); finalize_col: column_k[row] = num/denom + column_{k-1}[row];
  finalize_last: the last column becomes its inclusive prefix sum in trace-coset order; claimed_sum = its last value.
The reference's own unit tests (e.g. memory/table.rs:811-878, processor/table.rs:1070-…) check exactly this equality
against Stwo's generator; here Python integers take the generator's place.  Test infrastructure only."""
P = (1 << 31) - 1
LOG_N_LANES = 4


# --- QM31 = CM31[u]/(u^2 - (2 + i)), CM31 = M31[i]/(i^2 + 1); values are 4-tuples (a, b, c, d) = (a + bi) + (c + di)u
def c_mul(x, y):
    return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def q_add(x, y):
    return tuple((s + t) % P for s, t in zip(x, y))


def q_sub(x, y):
    return tuple((s - t) % P for s, t in zip(x, y))


def q_mul(x, y):
    a, b, c, d = (x[0], x[1]), (x[2], x[3]), (y[0], y[1]), (y[2], y[3])
    ac, bd, ad, bc = c_mul(a, c), c_mul(b, d), c_mul(a, d), c_mul(b, c)
    rbd = c_mul((2, 1), bd)                                   # R = 2 + i
    return ((ac[0] + rbd[0]) % P, (ac[1] + rbd[1]) % P, (ad[0] + bc[0]) % P, (ad[1] + bc[1]) % P)


def q_inv(x):
    a, b = (x[0], x[1]), (x[2], x[3])
    b2 = c_mul(b, b)
    rb2 = c_mul((2, 1), b2)
    a2 = c_mul(a, a)
    den = ((a2[0] - rb2[0]) % P, (a2[1] - rb2[1]) % P)        # a^2 - R b^2  (CM31)
    n = (den[0] * den[0] + den[1] * den[1]) % P
    ninv = pow(n, P - 2, P)
    dinv = (den[0] * ninv % P, (-den[1]) * ninv % P)
    ra, rb = c_mul(a, dinv), c_mul(b, dinv)
    return (ra[0], ra[1], (-rb[0]) % P, (-rb[1]) % P)


class Q:
    """A QM31 value with Python operators, so that the integer formulas of tests/air_model.py can be evaluated on secure-field
    mask values as well (tests/py_verifier.py evaluates the constraints at the out-of-domain point)."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = tuple(x % P for x in t)

    @staticmethod
    def of(v):
        return v if isinstance(v, Q) else Q((v, 0, 0, 0))

    def __add__(self, o): return Q(q_add(self.t, Q.of(o).t))
    __radd__ = __add__
    def __sub__(self, o): return Q(q_sub(self.t, Q.of(o).t))
    def __rsub__(self, o): return Q(q_sub(Q.of(o).t, self.t))
    def __mul__(self, o): return Q(q_mul(self.t, Q.of(o).t))
    __rmul__ = __mul__
    def __neg__(self): return Q(q_sub((0, 0, 0, 0), self.t))
    def __eq__(self, o): return self.t == Q.of(o).t
    def __hash__(self): return hash(self.t)
    def inv(self): return Q(q_inv(self.t))
    def __repr__(self): return "Q%r" % (self.t,)


def q_from(v):
    if isinstance(v, Q):
        return v.t
    return (v % P, 0, 0, 0)


MEM, INS, PROC = 0, 1, 2   # relation order of the C ABI's `elements` (memory, instruction, processor)
ALL7 = [0, 1, 2, 3, 4, 5, 6]
# per component (C-ABI numbering): one entry per LogUp column = (sign, index of the dummy flag column or None, relation, columns)
#   sign +1: numerator 1 - d (the component provides the tuple);  -1: numerator d - 1 (it consumes it)
FRACTIONS = {
    0: [(-1, 3, MEM, [0, 1, 2])],                                                   # memory:      clk, mp, mv
    1: [(-1, 3, INS, [0, 1, 2])],                                                   # instruction: ip, ci, ni
    2: [(+1, 3, INS, [0, 1, 2])],                                                   # program
    3: [(+1, 7, PROC, ALL7), (+1, 7, INS, [1, 2, 3]), (+1, 7, MEM, [0, 4, 5])],     # processor
    4: [(-1, 11, PROC, ALL7)], 5: [(-1, 11, PROC, ALL7)],                           # ] and [   (JumpColumn::D = 11)
    **{k: [(-1, 7, PROC, ALL7)] for k in range(6, 12)},                             # , < - . + >
    12: [(-1, None, PROC, ALL7)],                                                   # end of execution: numerator -1
}


def combine(el, rel, vals):
    """Relation::combine: sum_i alpha^i * v_i - z.  el = 96 words, 3 x {z[4], alpha_powers[7][4]}."""
    base = 32 * rel
    acc = (0, 0, 0, 0)
    for i, v in enumerate(vals):
        acc = q_add(acc, q_mul(tuple(el[base + 4 + 4 * i: base + 8 + 4 * i]), q_from(v)))
    return q_sub(acc, tuple(el[base: base + 4]))


def bit_reverse(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def coset_order_storage_indices(log):
    """storage index (bit-reversed circle-domain order) of the k-th point of the trace coset, k = 0 .. 2^log - 1."""
    n = 1 << log
    return [bit_reverse(k // 2 if k % 2 == 0 else (2 * n - 1 - k) // 2, log) for k in range(n)]


def logup_columns(comp, rows, el):
    """rows: table rows (each a list of main-column values).  Returns (columns, claimed_sum): 4 coordinate columns per LogUp
    column, each of length 16 * len(rows) in storage order."""
    n_rows = len(rows)
    log = (n_rows - 1).bit_length() + LOG_N_LANES
    assert n_rows == 1 << (log - LOG_N_LANES)
    n = 1 << log
    prev = [(0, 0, 0, 0)] * n
    out = []
    for sign, d_idx, rel, cols in FRACTIONS[comp]:
        cur = []
        for r in rows:                                   # vec_row = table row; all 16 lanes carry the same value
            d = r[d_idx] if d_idx is not None else 0
            num = q_from(1 - d) if sign > 0 else q_from(d - 1)
            frac = q_mul(num, q_inv(combine(el, rel, [r[c] for c in cols])))
            cur.extend([frac] * 16)
        cur = [q_add(f, p) for f, p in zip(cur, prev)]
        out.append(cur)
        prev = cur
    last, acc = out[-1], (0, 0, 0, 0)
    summed = list(last)
    for s in coset_order_storage_indices(log):
        acc = q_add(acc, last[s])
        summed[s] = acc
    out[-1] = summed
    columns = [[v[k] for v in col] for col in out for k in range(4)]
    return columns, acc
