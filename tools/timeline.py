"""GPU timeline of one warm fib19 proof: where the device sits idle between profiled scopes.
usage: python tools/timeline.py [out.json]   (needs a GPU)"""
import importlib, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("stwo-brainfuck_b200")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
be = pkg.CudaBackend(0, st.cuda_stream)
code = open(os.path.join(ROOT, "tests/golden/programs/fib19.bf"), "rb").read()
for _ in range(3):
    pkg.prove_brainfuck(be, code, b"", 24, overlap_host=False)
be.profile(True); be.profile_report()
pr = pkg.prove_brainfuck(be, code, b"", 24, overlap_host=False)
tl = be.profile_timeline()
rep = pr.report()
be.profile_report(); be.profile(False)
busy = sum(d for _, _, d in tl)
span = tl[-1][1] + tl[-1][2] - tl[0][1]
gaps = collections.defaultdict(lambda: [0.0, 0])
big = []
for (t0, s0, d0), (t1, s1, d1) in zip(tl, tl[1:]):
    g = s1 - (s0 + d0)
    gaps[(t0, t1)][0] += g; gaps[(t0, t1)][1] += 1
    if g > 0.15: big.append((round(s0 + d0, 3), round(g, 3), t0, t1))
print("scopes %d  span %.2f ms  busy %.2f ms  idle %.2f ms" % (len(tl), span, busy, span - busy))
print("stages", json.dumps(rep["stages_ms"]))
print("idle by (previous scope -> next scope):")
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-28s -> %-28s %8.3f ms over %4d gaps" % (k[0], k[1], v[0], v[1]))
print("gaps > 0.15 ms (at ms, length, prev, next):")
for b in big: print("  ", b)
if len(sys.argv) > 1:
    json.dump({"timeline": tl, "stages_ms": rep["stages_ms"], "span_ms": span, "busy_ms": busy}, open(sys.argv[1], "w"))
