import importlib, sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("stwo-brainfuck_b200")
be = pkg.CudaBackend(0)
code = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/programs/fib19.bf'),'rb').read()
for it in range(3):
    p = pkg.prove_brainfuck_sharded(be, None, code, b"", 24)
    r = p.report()
print("sharded world1", json.dumps(r["stages_ms"]), r["prove_ms"])
be.profile(True); be.profile_report()
p = pkg.prove_brainfuck_sharded(be, None, code, b"", 24)
print({k: round(v[0],2) for k,v in be.profile_report().items()})
