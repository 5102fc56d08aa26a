"""Pure-Python transcription of the 13 `evaluate()` bodies of the reference (which constraint, in which order, with which
numerator and relation) — an independent statement of the AIR, used to pin csrc/host/air.hpp (shared by the CUDA kernels and
the CPU oracle, so proof equality between those two cannot see a mistake in it).  Follows, under
crates/brainfuck_prover/src/components: memory/component.rs:62-137, instruction/component.rs, program/component.rs,
processor/component.rs:79-153, processor/instructions/{input,left,minus,output,plus,right}_component.rs,
.../jump/jump_if_{not_,}zero_component.rs, .../end_of_execution/component.rs.  The LogUp constraints are Stwo's
`LogupAtRow` as SURVEY.md A.8 restates it: per batch  (cur - prev_col) * denom - num,  the last batch with the cumulative
column's mask [-1, 0] and `+ is_first * total_sum`.  Test infrastructure only."""
import logup_model as M

P = M.P


def main_and_relations(comp, r, is_first):
    """The `add_constraint` values of one row (as QM31 tuples) and the `add_to_relation` entries (numerator, relation, values)."""
    out = []
    add = lambda v: out.append(M.q_from(v))   # ints (trace rows) or M.Q values (mask values at the out-of-domain point)
    one = 1
    if comp == 0:                                                   # memory/component.rs:62-137
        clk, mp, mv, d, nclk, nmp, nmv, nd = r
        add(is_first * clk); add(is_first * mp); add(is_first * mv); add(is_first * d)
        add(d * (d - one)); add(nd * (nd - one))
        add((nmp - mp) * (nmp - mp - one))
        add((nmp - mp - one) * (nclk - clk - one))
        add((nmp - mp) * nmv)
        add(d * (nmp - mp)); add(d * (nmv - mv))
        rel = [(d - one, M.MEM, [clk, mp, mv])]
    elif comp == 1:                                                 # instruction/component.rs
        ip, ci, ni, d, nip, nci, nni, nd = r
        add(is_first * ip)
        add(d * (d - one)); add(nd * (nd - one))
        add(d * ci); add(d * ni); add(nd * nci); add(nd * nni)
        add((nip - ip) * (nip - ip - one))
        add((nip - ip - one) * (nci - ci)); add((nip - ip - one) * (nni - ni))
        rel = [(d - one, M.INS, [ip, ci, ni])]
    elif comp == 2:                                                 # program/component.rs
        ip, ci, ni, d = r
        add(is_first * ip); add(d * (d - one)); add(d * ci); add(d * ni)
        rel = [(one - d, M.INS, [ip, ci, ni])]
    elif comp == 3:                                                 # processor/component.rs:79-153
        clk, ip, ci, ni, mp, mv, mvi, d, nclk = r
        add(is_first * clk); add(is_first * ip); add(is_first * mp); add(is_first * mv)
        add(mv * (mv * mvi - one)); add(mvi * (mv * mvi - one))
        add(nclk - clk - one)
        num = one - d
        rel = [(num, M.PROC, [clk, ip, ci, ni, mp, mv, mvi]), (num, M.INS, [ip, ci, ni]), (num, M.MEM, [clk, mp, mv])]
    elif comp in (4, 5):                                            # jump/jump_if_{not_zero,zero}_component.rs
        clk, ip, ci, ni, mp, mv, mvi, nclk, nip, nmp, nmv, d, is_mv_zero = r
        add(ci * (ci - ord("]" if comp == 4 else "[")))
        add(nclk - clk - one)
        add(d * (d - one)); add(d * mv); add(d * ci)
        if comp == 4:
            add((d - one) * (is_mv_zero * (nip - ip - 2) + mv * (nip - ni)))
        else:
            add((d - one) * (mv * (nip - ip - 2) + is_mv_zero * (nip - (ni + one))))
        add(nmp - mp); add(nmv - mv)
        rel = [(d - one, M.PROC, [clk, ip, ci, ni, mp, mv, mvi])]
    elif 6 <= comp <= 11:                                           # processor/instructions/*_component.rs
        clk, ip, ci, ni, mp, mv, mvi, d, nip, nmp, nmv = r
        add(ci * (ci - ord(",<-.+>"[comp - 6])))
        add(d * (d - one)); add(d * mv); add(d * ci)
        add((one - d) * (nip - ip - one))
        if comp == 6:                                               # ,  input: mp unchanged, mv free
            add(nmp - mp)
        elif comp == 7:                                             # <  left: mp decreases
            add((one - d) * (nmp - mp + one))
        elif comp == 8:                                             # -  minus
            add(nmp - mp); add((one - d) * (nmv - mv + one))
        elif comp == 9:                                             # .  output
            add(nmp - mp); add(nmv - mv)
        elif comp == 10:                                            # +  plus
            add(nmp - mp); add((one - d) * (nmv - mv - one))
        else:                                                       # >  right
            add((one - d) * (nmp - mp - one))
        rel = [(d - one, M.PROC, [clk, ip, ci, ni, mp, mv, mvi])]
    else:                                                           # end_of_execution/component.rs
        clk, ip, ci, ni, mp, mv, mvi = r
        add(ci)
        rel = [(-one, M.PROC, [clk, ip, ci, ni, mp, mv, mvi])]
    return out, rel


def constraints(comp, r, is_first, el, ext_cur, ext_prev_last, total_sum):
    """Values of all constraints of component `comp` on one row.
    r: main-trace values of the row; is_first: 0/1; ext_cur[b]: value of LogUp column b on this row (QM31 4-tuples);
    ext_prev_last: last LogUp column on the previous row in coset order; total_sum: the component's claimed sum."""
    out, rel = main_and_relations(comp, r, is_first)
    # LogupAtRow::finalize
    prev_col = (0, 0, 0, 0)
    for b, (num, relation, vals) in enumerate(rel):
        den, numq = M.combine(el, relation, vals), M.q_from(num)
        if b + 1 < len(rel):
            diff = M.q_sub(ext_cur[b], prev_col)
            prev_col = ext_cur[b]
        else:
            diff = M.q_sub(M.q_sub(ext_cur[b], ext_prev_last), prev_col)
            diff = M.q_add(diff, M.q_mul(total_sum, M.q_from(is_first)))
        out.append(M.q_sub(M.q_mul(diff, den), numq))
    return out


def constraint_values(comp, rows, el, nat):
    """All constraint values of `comp` at natural trace-domain row `nat`, LogUp columns generated from the table."""
    cols, total = M.logup_columns(comp, rows, el)
    log = (len(rows) * 16).bit_length() - 1
    s = M.bit_reverse(nat, log)
    order = M.coset_order_storage_indices(log)
    prev = order[(order.index(s) - 1) % len(order)]
    n_ext = len(cols) // 4
    ext = lambda b, i: tuple(cols[4 * b + k][i] for k in range(4))
    return constraints(comp, rows[s // 16], 1 if nat == 0 else 0, el, [ext(b, s) for b in range(n_ext)], ext(n_ext - 1, prev), total)


def first_violation(comp, rows, el):
    """assert_constraints in Python: (natural row, constraint index, value) of the first non-zero constraint, or None.
    Returns also the claimed sum."""
    cols, total = M.logup_columns(comp, rows, el)
    log = (len(rows) * 16).bit_length() - 1
    order = M.coset_order_storage_indices(log)
    coset_of = {s: k for k, s in enumerate(order)}
    n_ext = len(cols) // 4
    ext = lambda b, i: tuple(cols[4 * b + k][i] for k in range(4))
    for nat in range(1 << log):
        s = M.bit_reverse(nat, log)
        prev = order[(coset_of[s] - 1) % len(order)]
        vals = constraints(comp, rows[s // 16], 1 if nat == 0 else 0, el, [ext(b, s) for b in range(n_ext)], ext(n_ext - 1, prev), total)
        for k, v in enumerate(vals):
            if v != (0, 0, 0, 0):
                return (nat, k, v), total
    return None, total
