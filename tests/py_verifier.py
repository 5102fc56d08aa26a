"""An independent verifier for the wire-format proof (`sbf_proof_json`): pure Python integers + hashlib, written from the
protocol description (crates/brainfuck_prover/src/brainfuck_air/mod.rs:738-797 `verify_brainfuck`,
and stwo-prover 0.1.1 @ 31e8dbc core/{prover/mod.rs `verify`, pcs/verifier.rs, pcs/quotients.rs `fri_answers`, fri.rs
`FriVerifier`, vcs/verifier.rs `MerkleVerifier`, queries.rs, channel/blake2s.rs}).  It shares NO code with
stwo-brainfuck_b200/csrc/host/ — not the channel, not the mask layout, not the AIR (tests/air_model.py is a separate
transcription of the 13 `evaluate()` bodies), not the quotient formulas, not the FRI or Merkle walks — so a convention that
`prover.hpp` and `verifier.hpp` got wrong TOGETHER (transcript order, mask points, column order inside a tree, fold
positions, witness order, JSON shape) makes this verifier reject.  Test infrastructure only.

What it cannot establish: that upstream Stwo uses exactly these conventions (no Rust toolchain here or on the GPU box:
profiles/r2_gpu_box_probe.txt) — parity with the real reference stays unpinned; tools/make_reference_goldens.sh is the
recipe that closes it on a machine with cargo."""
import hashlib
import json
import struct

import air_model
import logup_model as M
from logup_model import P, Q

N_MAIN = [8, 8, 4, 9, 13, 13, 11, 11, 11, 11, 11, 11, 7]
N_LOGUP = [1, 1, 1, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1]
NAMES = ["memory", "instruction", "program", "processor", "jump_if_not_zero", "jump_if_zero", "input_instruction", "left_instruction",
         "minus_instruction", "output_instruction", "plus_instruction", "right_instruction", "end_of_execution"]
LOG_N_LANES = 4
M32 = 0xFFFFFFFF


class Reject(Exception):
    pass


def need(cond, why):
    if not cond:
        raise Reject(why)


# ------------------------------------------------------------------------------------------------ Blake2s F and the channel
IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
SIGMA = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
         [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
         [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
         [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
         [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0]]


def compress(h, m):
    """RFC 7693 F with zero counters and flags (Blake2sMerkleHasher::hash_node, Blake2sChannel::mix_u64)."""
    v = list(h) + list(IV)
    ror = lambda x, r: ((x >> r) | (x << (32 - r))) & M32

    def g(a, b, c, d, x, y):
        v[a] = (v[a] + v[b] + x) & M32; v[d] = ror(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & M32; v[b] = ror(v[b] ^ v[c], 12)
        v[a] = (v[a] + v[b] + y) & M32; v[d] = ror(v[d] ^ v[a], 8)
        v[c] = (v[c] + v[d]) & M32; v[b] = ror(v[b] ^ v[c], 7)
    for s in SIGMA:
        g(0, 4, 8, 12, m[s[0]], m[s[1]]); g(1, 5, 9, 13, m[s[2]], m[s[3]]); g(2, 6, 10, 14, m[s[4]], m[s[5]]); g(3, 7, 11, 15, m[s[6]], m[s[7]])
        g(0, 5, 10, 15, m[s[8]], m[s[9]]); g(1, 6, 11, 12, m[s[10]], m[s[11]]); g(2, 7, 8, 13, m[s[12]], m[s[13]]); g(3, 4, 9, 14, m[s[14]], m[s[15]])
    return [h[i] ^ v[i] ^ v[8 + i] for i in range(8)]


def hash_node(children, values):
    """state = 0^8; F(state, left || right) if there are children; then F over the column values, 16 at a time, zero padded."""
    st = [0] * 8
    if children is not None:
        st = compress(st, list(struct.unpack("<16I", children[0] + children[1])))
    for o in range(0, len(values), 16):
        chunk = list(values[o:o + 16])
        st = compress(st, chunk + [0] * (16 - len(chunk)))
    return struct.pack("<8I", *st)


class Channel:
    def __init__(self):
        self.digest, self.n_sent = bytes(32), 0

    def _set(self, d):
        self.digest, self.n_sent = d, 0

    def mix_root(self, root):
        self._set(hashlib.blake2s(self.digest + root).digest())

    def mix_felts(self, felts):
        self._set(hashlib.blake2s(self.digest + b"".join(struct.pack("<4I", *f) for f in felts)).digest())

    def mix_u64(self, v):
        self._set(struct.pack("<8I", *compress(list(struct.unpack("<8I", self.digest)), [v & M32, v >> 32] + [0] * 14)))

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
        return tuple(self.draw_base_felts()[:4])

    def draw_felts(self, n):
        out = []
        while len(out) < n:
            f = self.draw_base_felts()
            out += [tuple(f[:4]), tuple(f[4:])]
        return out[:n]

    def trailing_zeros(self):
        v = int.from_bytes(self.digest[:16], "little")
        return 128 if v == 0 else (v & -v).bit_length() - 1


# ------------------------------------------------------------------------------------------------ circle group
GEN = (2, 1268011823)


def padd(p, q):
    return ((p[0] * q[0] - p[1] * q[1]) % P, (p[0] * q[1] + p[1] * q[0]) % P)


def point_at(idx):
    idx %= 1 << 31
    r, b = (1, 0), GEN
    while idx:
        if idx & 1:
            r = padd(r, b)
        b = padd(b, b)
        idx >>= 1
    return r


def bit_reverse(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def canonic_domain_at(log, i):
    """CanonicCoset(log).circle_domain().at(i): half coset G^(2^(30-log)) * <G^(2^(32-log))>, then its conjugate."""
    half = 1 << (log - 1)
    if i < half:
        return point_at((1 << (30 - log)) + (i << (32 - log)))
    x, y = point_at((1 << (30 - log)) + ((i - half) << (32 - log)))
    return (x, (-y) % P)


def line_domain_x(log, i):
    """LineDomain(Coset::half_odds(log)).at(i): x of G^(2^(29-log)) * <G^(2^(31-log))>."""
    return point_at((1 << (29 - log)) + (i << (31 - log)))[0]


def qpadd(p, q):
    return (p[0] * q[0] - p[1] * q[1], p[0] * q[1] + p[1] * q[0])


def conj(q):
    """QM31 complex conjugate over CM31: u -> -u."""
    return Q((q.t[0], q.t[1], -q.t[2], -q.t[3]))


def cm31_inv(c):
    n = pow((c[0] * c[0] + c[1] * c[1]) % P, P - 2, P)
    return (c[0] * n % P, (-c[1]) * n % P)


def q_mul_cm31(q, c):
    a, b = M.c_mul((q.t[0], q.t[1]), c), M.c_mul((q.t[2], q.t[3]), c)
    return Q((a[0], a[1], b[0], b[1]))


# ------------------------------------------------------------------------------------------------ parsing
def parse(js):
    p = json.loads(js)
    need(list(p.keys()) == ["claim", "interaction_claim", "proof"], "top-level shape")
    need(list(p["claim"].keys()) == NAMES and list(p["interaction_claim"].keys()) == NAMES, "claim component order")
    q = lambda v: tuple(v[0]) + tuple(v[1])
    h = lambda v: bytes(v)
    out = {"log_size": [p["claim"][n]["log_size"] for n in NAMES],
           "claimed": [q(p["interaction_claim"][n]["claimed_sum"]) for n in NAMES]}
    s = p["proof"]
    dec = lambda d: {"hash_witness": [h(x) for x in d["hash_witness"]], "column_witness": list(d["column_witness"])}
    layer = lambda l: {"fri_witness": [q(x) for x in l["fri_witness"]], "decommitment": dec(l["decommitment"]), "commitment": h(l["commitment"])}
    out.update(commitments=[h(x) for x in s["commitments"]],
               sampled=[[[q(v) for v in col] for col in tree] for tree in s["sampled_values"]],
               decommitments=[dec(d) for d in s["decommitments"]],
               queried=s["queried_values"], pow=s["proof_of_work"],
               first=layer(s["fri_proof"]["first_layer"]), inner=[layer(l) for l in s["fri_proof"]["inner_layers"]],
               last=[q(x) for x in s["fri_proof"]["last_layer_poly"]["coeffs"]])
    need(len(out["last"]) == 1 << s["fri_proof"]["last_layer_poly"]["log_size"], "last layer polynomial size")
    return out


# ------------------------------------------------------------------------------------------------ Merkle
def merkle_verify(root, column_logs, queries_by_log, queried, dec):
    """MerkleVerifier::verify.  column_logs: log size of every column in commitment order; queries_by_log: sorted positions per
    log size; queried[c]: the values of column c at its queries, in order."""
    hw, cw = iter(dec["hash_witness"]), iter(dec["column_witness"])
    qv = [iter(v) for v in queried]
    prev = None          # [(node index, hash)] of the layer below
    top = max(column_logs)
    for lg in range(top, -1, -1):
        cols = [c for c, l in enumerate(column_logs) if l == lg]
        colq = list(queries_by_log.get(lg, []))
        nodes = sorted(set(colq) | ({i // 2 for i, _ in prev} if prev else set()))
        below = dict(prev) if prev else None
        cur = []
        for node in nodes:
            children = None
            if lg < top:
                kids = []
                for k in (2 * node, 2 * node + 1):
                    if below is not None and k in below:
                        kids.append(below[k])
                    else:
                        try:
                            kids.append(next(hw))
                        except StopIteration:
                            raise Reject("hash witness too short")
                children = kids
            vals = []
            for c in cols:
                try:
                    vals.append(next(qv[c]) if node in colq else next(cw))
                except StopIteration:
                    raise Reject("column values too short")
            cur.append((node, hash_node(children, vals)))
        prev = cur
    need(next(hw, None) is None and next(cw, None) is None, "witness too long")
    need(all(next(it, None) is None for it in qv), "queried values too long")
    need(prev and prev[0][1] == root, "Merkle root mismatch")


# ------------------------------------------------------------------------------------------------ queries
def generate_queries(ch, log_domain, n):
    out, cnt = set(), 0
    while True:
        for w in struct.unpack("<8I", ch.draw_random_bytes()):
            out.add(w & ((1 << log_domain) - 1))
            cnt += 1
            if cnt == n:
                return sorted(out)


def fold_queries(qs, k):
    return sorted({q >> k for q in qs})


# ------------------------------------------------------------------------------------------------ the verifier
def verify(js, log_max_rows, pow_bits=5, log_blowup=1, n_queries=3, log_last_layer=0):
    pr = parse(js)
    ls = pr["log_size"]
    need(all(LOG_N_LANES <= x <= log_max_rows for x in ls), "component log size out of range")
    ch = Channel()
    # ---- column log sizes of the three traces (BrainfuckClaim::log_sizes, brainfuck_air/mod.rs:118-143; IS_FIRST_LOG_SIZES :453-464)
    pre_logs = list(range(log_max_rows, LOG_N_LANES - 1, -1))
    main_logs = [ls[c] for c in range(13) for _ in range(N_MAIN[c])]
    int_logs = [ls[c] for c in range(13) for _ in range(4 * N_LOGUP[c])]
    comp_log = max(ls) + 1
    tree_logs = [pre_logs, main_logs, int_logs, [comp_log] * 4]
    need(len(pr["commitments"]) == 4 and len(pr["sampled"]) == 4 and len(pr["queried"]) == 4 and len(pr["decommitments"]) == 4, "tree count")
    # ---- transcript up to the composition commitment (verify_brainfuck, mod.rs:738-797)
    ch.mix_root(pr["commitments"][0])
    for x in ls:
        ch.mix_u64(x)
    ch.mix_root(pr["commitments"][1])
    el = []
    for _ in range(3):                     # memory, instruction, processor lookup elements: z, alpha -> alpha powers (mod.rs:149-165)
        z, alpha = ch.draw_felts(2)
        pw, cur = [], (1, 0, 0, 0)
        for _ in range(7):
            pw.append(cur)
            cur = M.q_mul(cur, alpha)
        el += list(z) + [w for t in pw for w in t]
    tot = (0, 0, 0, 0)
    for s in pr["claimed"]:
        tot = M.q_add(tot, s)
    need(tot == (0, 0, 0, 0), "InvalidLogupSum")
    for s in pr["claimed"]:
        ch.mix_felts([s])
    ch.mix_root(pr["commitments"][2])
    random_coeff = ch.draw_felt()
    ch.mix_root(pr["commitments"][3])
    # ---- out-of-domain point and mask points
    t = Q(ch.draw_felt())
    t2 = t * t
    inv = (t2 + 1).inv()
    oods = ((1 - t2) * inv, (t + t) * inv)
    masks = [[[] for _ in pre_logs], [], [], [[oods]] * 4]
    for c in range(13):
        masks[0][log_max_rows - ls[c]] = [oods]          # IsFirst(log_size) is the only preprocessed column a component reads
    for c in range(13):
        masks[1] += [[oods]] * N_MAIN[c]
        sx, sy = point_at(1 << (31 - ls[c]))             # step of CanonicCoset(log_size); offset -1 subtracts it
        prev = qpadd(oods, (Q.of(sx), Q.of((-sy) % P)))
        n = 4 * N_LOGUP[c]
        masks[2] += [[oods]] * (n - 4) + [[prev, oods]] * 4
    for tr in range(4):
        need(len(pr["sampled"][tr]) == len(tree_logs[tr]), "sampled values: column count")
        for col, m in zip(pr["sampled"][tr], masks[tr]):
            need(len(col) == len(m), "sampled values: mask size")
    # ---- composition polynomial at the point from the sampled mask values (Horner over all constraints in component order)
    acc = Q.of(0)
    rc = Q(random_coeff)
    mo = io = 0
    for c in range(13):
        vx = oods[0]
        for _ in range(1, ls[c]):
            vx = 2 * vx * vx - 1                          # coset_vanishing of the canonic coset: pi^(log-1)(x)
        dinv = vx.inv()
        row = [Q(pr["sampled"][1][mo + j][0]) for j in range(N_MAIN[c])]
        is_first = Q(pr["sampled"][0][log_max_rows - ls[c]][0])
        iu = [(1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)]
        ext = lambda b, s: sum((Q(pr["sampled"][2][io + 4 * b + k][s]) * Q(iu[k]) for k in range(4)), Q.of(0))   # from_partial_evals
        nb = N_LOGUP[c]
        cur = [ext(b, 0).t for b in range(nb - 1)] + [ext(nb - 1, 1).t]
        vals = air_model.constraints(c, row, is_first, el, cur, ext(nb - 1, 0).t, pr["claimed"][c])
        for v in vals:
            acc = acc * rc + dinv * Q(v)
        mo += N_MAIN[c]
        io += 4 * nb
    comp = sum((Q(pr["sampled"][3][k][0]) * Q(iu[k]) for k in range(4)), Q.of(0))
    need(comp == acc, "OodsNotMatching")
    # ---- verify_values: transcript
    ch.mix_felts([v for tr in pr["sampled"] for col in tr for v in col])
    alpha = Q(ch.draw_felt())
    lde_sizes = sorted({l + log_blowup for tl in tree_logs for l in tl}, reverse=True)
    top = lde_sizes[0]
    # ---- FRI commit phase
    ch.mix_root(pr["first"]["commitment"])
    circle_alpha = Q(ch.draw_felt())
    line_log = top - 1
    inner_alpha = []
    for L in pr["inner"]:
        ch.mix_root(L["commitment"])
        inner_alpha.append(Q(ch.draw_felt()))
    need(line_log - len(pr["inner"]) == log_last_layer + log_blowup, "InvalidNumFriLayers")
    need(len(pr["last"]) <= 1 << log_last_layer, "LastLayerDegreeInvalid")
    ch.mix_felts(pr["last"])
    ch.mix_u64(pr["pow"])
    need(ch.trailing_zeros() >= pow_bits, "ProofOfWork")
    queries = generate_queries(ch, top, n_queries)
    q_by_log = {lg: fold_queries(queries, top - lg) for lg in lde_sizes}
    # ---- decommitments of the four trees at the query positions
    for tr in range(4):
        need(len(pr["queried"][tr]) == len(tree_logs[tr]), "queried values: column count")
        merkle_verify(pr["commitments"][tr], [l + log_blowup for l in tree_logs[tr]], q_by_log, pr["queried"][tr], pr["decommitments"][tr])
    # ---- fri_answers: DEEP quotients at the query positions, per LDE size (descending), columns in tree order
    answers = {}
    for lg in lde_sizes:
        cols = [(tr, c) for tr in range(4) for c, l in enumerate(tree_logs[tr]) if l + log_blowup == lg]
        batches = {}
        for k, (tr, c) in enumerate(cols):
            for pt, val in zip(masks[tr][c], pr["sampled"][tr][c]):
                key = pt[0].t + pt[1].t                     # BTreeMap keyed by the point: lexicographic on the eight words
                batches.setdefault(key, (pt, []))[1].append((k, Q(val)))
        order = sorted(batches)
        out = []
        for qi, pos in enumerate(q_by_log[lg]):
            dx, dy = canonic_domain_at(lg, bit_reverse(pos, lg))
            row = [pr["queried"][tr][c][qi] for tr, c in cols]
            acc_q = Q.of(0)
            for key in order:
                (px, py), entries = batches[key]
                num, al = Q.of(0), Q.of(1)
                cc = conj(py) - py
                for k, v in entries:
                    al = al * alpha
                    a = conj(v) - v
                    b = v * cc - a * py
                    num = num + al * (cc * row[k] - (a * dy + b))
                prx, pix = (px.t[0], px.t[1]), (px.t[2], px.t[3])
                pry, piy = (py.t[0], py.t[1]), (py.t[2], py.t[3])
                d1 = M.c_mul(((prx[0] - dx) % P, prx[1]), piy)
                d2 = M.c_mul(((pry[0] - dy) % P, pry[1]), pix)
                den = ((d1[0] - d2[0]) % P, (d1[1] - d2[1]) % P)
                apow = Q.of(1)
                for _ in entries:
                    apow = apow * alpha
                acc_q = acc_q * apow + q_mul_cm31(num, cm31_inv(den))
            out.append(acc_q)
        answers[lg] = out
    # ---- FRI decommit
    def rebuild(qs, evals, witness, log):
        """fold cosets {2k, 2k+1} that contain a query: (positions to decommit, [(k, (e0, e1))])"""
        ev = dict(zip(qs, evals))
        pos, pairs = [], []
        for k in fold_queries(qs, 1):
            pair = []
            for p in (2 * k, 2 * k + 1):
                pos.append(p)
                if p in ev:
                    pair.append(ev[p])
                else:
                    try:
                        pair.append(Q(next(witness)))
                    except StopIteration:
                        raise Reject("FRI witness too short")
            pairs.append((k, pair))
        return pos, pairs

    # first layer: every quotient column (4 coordinates each), one tree
    wit = iter(pr["first"]["fri_witness"])
    first_pos, first_vals, first_pairs = {}, [], {}
    for lg in lde_sizes:
        pos, pairs = rebuild(q_by_log[lg], answers[lg], wit, lg)
        first_pos[lg] = pos
        first_pairs[lg] = pairs
        flat = [e for _, pair in pairs for e in pair]
        first_vals += [[e.t[k] for e in flat] for k in range(4)]
    need(next(wit, None) is None, "FRI witness too long")
    merkle_verify(pr["first"]["commitment"], [lg for lg in lde_sizes for _ in range(4)], first_pos, first_vals, pr["first"]["decommitment"])
    # inner layers
    lq = fold_queries(queries, 1)
    le = [Q.of(0)] * len(lq)
    for li, L in enumerate(pr["inner"]):
        lg = line_log - li
        if lg + 1 in first_pairs:      # a circle column of this size folds into the line here, with the first layer's alpha
            folded = []
            for k, (e0, e1) in first_pairs[lg + 1]:
                _, y = canonic_domain_at(lg + 1, bit_reverse(2 * k, lg + 1))
                f0, f1 = e0 + e1, (e0 - e1) * pow(y, P - 2, P)
                folded.append(f0 + circle_alpha * f1)
            need(len(folded) == len(le), "FRI: fold positions")
            a2 = circle_alpha * circle_alpha
            le = [x * a2 + f for x, f in zip(le, folded)]
        wit = iter(L["fri_witness"])
        pos, pairs = rebuild(lq, le, wit, lg)
        need(next(wit, None) is None, "FRI witness too long")
        flat = [e for _, pair in pairs for e in pair]
        merkle_verify(L["commitment"], [lg] * 4, {lg: pos}, [[e.t[k] for e in flat] for k in range(4)], L["decommitment"])
        nxt = []
        for k, (e0, e1) in pairs:
            x = line_domain_x(lg, bit_reverse(2 * k, lg))
            f0, f1 = e0 + e1, (e0 - e1) * pow(x, P - 2, P)
            nxt.append(f0 + inner_alpha[li] * f1)
        lq, le = fold_queries(lq, 1), nxt
    need(set(first_pairs) <= {line_log - li + 1 for li in range(len(pr["inner"]))}, "FRI: a column was never folded in")
    # last layer: a constant (degree bound 2^0)
    need(log_last_layer == 0 and len(pr["last"]) == 1, "only log_last_layer_degree_bound = 0 is supported")
    for e in le:
        need(e == Q(pr["last"][0]), "LastLayerEvaluationsInvalid")
    return True
