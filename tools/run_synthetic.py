"""BASELINE.json configs[3]: synthetic looping program sized to a 2^24-row Processor column
('+' * 262000 + '[-]': 786002 steps; Processor = Memory = Instruction = log 24)."""
import importlib, sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("stwo-brainfuck_b200")
be = pkg.CudaBackend(0)
code = b"+" * 262000 + b"[-]"
for it in range(3):
    t = time.time()
    proof = pkg.prove_brainfuck(be, code, b"", 24)
    dt = time.time() - t
    r = proof.report()
    t = time.time(); proof.verify(); tv = time.time() - t
    print(json.dumps({"wall_s": dt, "verify_s": tv, "proof_bytes": len(proof.json()), **r}))
    del proof
