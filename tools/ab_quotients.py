import importlib, json, os, sys
sys.path.insert(0, '/root/repo')
import torch
pkg = importlib.import_module("stwo-brainfuck_b200")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
be = pkg.CudaBackend(0, st.cuda_stream)
code = open('/root/repo/tests/golden/programs/fib19.bf','rb').read()
for _ in range(2): pkg.prove_brainfuck(be, code, b"", 24, overlap_host=False)
be.profile(True); be.profile_report()
for _ in range(3): pkg.prove_brainfuck(be, code, b"", 24, overlap_host=False)
r = be.profile_report()
print(os.path.basename(pkg.LIB_PATH), {k: round(v[0]/3, 3) for k, v in r.items() if k in ("accumulate_quotients", "eval_constraints", "merkle_commit_layer")})
