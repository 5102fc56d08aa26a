"""Generates tests/golden/proof_hashes.json: sha256 of the proof JSON produced by the CPU oracle prover for a set of
programs.  These are golden vectors of THIS repository's oracle (the reference holds none); they pin the
whole transcript (roots, claims, OODS values, FRI layers, decommitments) against regressions on both provers.
Run here: python tests/golden/make_proof_hashes.py"""
import ctypes, hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import proof_canon  # noqa: E402
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
L.orc_prove_json.restype = ctypes.c_void_p
L.orc_last_error.restype = ctypes.c_char_p
CASES = [("with_input", "+>,<[>+.<-]", "01", 10), ("no_input", "+++>++<[->+<]>.", "", 10), ("jump_mid", "++[>+<-]>[-]<", "", 10),
         ("a-bc", None, "61", 12), ("hello_kakarot", None, "", 17), ("collatz", None, "370a", 21)]
# fib19 at the reference's LOG_MAX_ROWS = 24 (BASELINE.json configs[1], 1.14 G LDE cells) takes the oracle ~6 minutes on 8 cores
# and ~20 GB: only with --full.  Its entry in proof_hashes.json was produced that way (349.8 s) and is otherwise carried over.
import sys
# synthetic_2p24 (BASELINE.json configs[3]: '+' * 262000 + '[-]', 786 002 steps, Processor = Memory = Instruction = log 24,
# programs/synthetic_2p24.bf) is about twice that; same rule.
FULL = [("fib19", None, "", 24), ("synthetic_2p24", None, "", 24)]
# The small shipped programs at the reference's shipped LOG_MAX_ROWS = 24 (brainfuck_air/mod.rs:427-428), where the 21-column
# preprocessed tree and a 2^25-row FRI dominate whatever the program is (BASELINE.json configs[0] and [2]): --shipped,
# about half a minute of oracle time each on 16 cores.  name@24 -> programs/name.bf
SHIPPED = [("hello_kakarot@24", None, "", 24), ("collatz@24", None, "370a", 24)]
if "--full" in sys.argv:
    CASES += FULL
if "--shipped" in sys.argv:
    CASES = SHIPPED
out = {}
if os.path.exists(os.path.join(HERE, "proof_hashes.json")):
    old = json.load(open(os.path.join(HERE, "proof_hashes.json")))
    keep = (FULL if "--full" not in sys.argv else []) + (SHIPPED if "--shipped" not in sys.argv else [])
    if "--shipped" in sys.argv:
        out = dict(old)              # everything else is carried over
    for name, _, _, _ in keep:
        if name in old:
            out[name] = old[name]
for name, code, stdin_hex, lmr in CASES:
    src = code.encode() if code else open(os.path.join(HERE, "programs", name.split("@")[0] + ".bf"), "rb").read()
    stdin = bytes.fromhex(stdin_hex)
    p = L.orc_prove_json(src, stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(lmr), 1)
    assert p, L.orc_last_error()
    js = ctypes.string_at(p)
    L.orc_free(ctypes.c_void_p(p))
    js = proof_canon.canonical(js)   # hashes are over the canonical text (tests/proof_canon.py), not the wire spelling
    out[name] = {"code": code, "stdin_hex": stdin_hex, "log_max_rows": lmr, "proof_bytes": len(js), "sha256": hashlib.sha256(js).hexdigest()}
    print(name, out[name]["sha256"][:16], len(js))
json.dump(out, open(os.path.join(HERE, "proof_hashes.json"), "w"), indent=1)
