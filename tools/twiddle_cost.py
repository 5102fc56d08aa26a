"""What the per-context twiddle cache saves: fib19 proofs with SBF_NO_TWIDDLE_CACHE (the reference recomputes the tree of
half_odds(26) in every proof, brainfuck_air/mod.rs:480-484) next to the default."""
import ctypes, importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("stwo-brainfuck_b200")
be = pkg.CudaBackend(0)
code = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "programs", "fib19.bf"), "rb").read()
lib = be._lib


def prove(flags):
    h = ctypes.c_void_p()
    t = time.perf_counter()
    rc = lib.sbf_prove(be._ctx, ctypes.c_char_p(code), ctypes.c_char_p(b""), ctypes.c_size_t(0), ctypes.c_uint32(24), ctypes.c_uint32(flags), ctypes.byref(h))
    assert rc == 0
    p = pkg.Proof(lib, h)
    js = p.json()
    return time.perf_counter() - t, p.report(), js


out = {}
for name, flags in (("cached", 0), ("recomputed", 2), ("cached_again", 0), ("recomputed_again", 2)):
    runs = [prove(flags) for _ in range(4)][1:]
    out[name] = {"e2e_ms": sum(r[0] for r in runs) / len(runs) * 1e3, "twiddles_stage_ms": sum(r[1]["stages_ms"]["twiddles"] for r in runs) / len(runs)}
    out.setdefault("proofs", set()).add(runs[0][2])
out["same_proof"] = len(out.pop("proofs")) == 1
print(json.dumps(out))
