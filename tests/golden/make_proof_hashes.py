"""Generates tests/golden/proof_hashes.json: sha256 of the proof JSON produced by the CPU oracle prover for a set of
programs.  These are golden vectors of THIS repository's oracle (the reference holds none); they pin the
whole transcript (roots, claims, OODS values, FRI layers, decommitments) against regressions on both provers.
Run here: python tests/golden/make_proof_hashes.py"""
import ctypes, hashlib, json, os
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
L.orc_prove_json.restype = ctypes.c_void_p
L.orc_last_error.restype = ctypes.c_char_p
CASES = [("with_input", "+>,<[>+.<-]", "01", 10), ("no_input", "+++>++<[->+<]>.", "", 10), ("jump_mid", "++[>+<-]>[-]<", "", 10),
         ("a-bc", None, "61", 12), ("hello_kakarot", None, "", 17), ("collatz", None, "370a", 21)]
out = {}
for name, code, stdin_hex, lmr in CASES:
    src = code.encode() if code else open(os.path.join(HERE, "programs", name + ".bf"), "rb").read()
    stdin = bytes.fromhex(stdin_hex)
    p = L.orc_prove_json(src, stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(lmr), 1)
    assert p, L.orc_last_error()
    js = ctypes.string_at(p)
    L.orc_free(ctypes.c_void_p(p))
    out[name] = {"code": code, "stdin_hex": stdin_hex, "log_max_rows": lmr, "proof_bytes": len(js), "sha256": hashlib.sha256(js).hexdigest()}
    print(name, out[name]["sha256"][:16], len(js))
json.dump(out, open(os.path.join(HERE, "proof_hashes.json"), "w"), indent=1)
