import importlib, sys, time, json
sys.path.insert(0, '/root/repo')
pkg = importlib.import_module("stwo-brainfuck_b200")
be = pkg.CudaBackend(0)
code = open('/root/repo/tests/golden/programs/fib19.bf','rb').read()
for it in range(3):
    t = time.time()
    proof = pkg.prove_brainfuck(be, code, b"", 24)
    dt = time.time() - t
    r = proof.report()
    print("fib19 prove wall %.3fs" % dt, json.dumps(r))
    t = time.time(); proof.verify(); print("verify %.3fs" % (time.time()-t), "proof bytes", len(proof.json()))
    del proof
