"""One fib19 proof on a fresh context — the target of the ncu captures under profiles/ (tools/ncu_capture.sh)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("stwo-brainfuck_b200")
be = pkg.CudaBackend(0)
pr = pkg.prove_brainfuck(be, open(os.path.join(ROOT, "tests/golden/programs/fib19.bf"), "rb").read(), b"", 24, overlap_host=False)
pr.verify()
print("ok", pr.report()["prove_ms"])
