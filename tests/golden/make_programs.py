"""Copies the reference's example programs (workload inputs, BASELINE.json:configs) into tests/golden/programs/ so that the
GPU box (which has no /root/reference) can run them.  They are inputs, not source code of the reference.
Run here: python tests/golden/make_programs.py"""
import os, shutil
src = "/root/reference/brainfuck_programs"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "programs")
os.makedirs(dst, exist_ok=True)
for f in sorted(os.listdir(src)):
    if f.endswith(".bf"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
print(sorted(os.listdir(dst)))
