"""Goldens produced by the REAL reference (tools/make_reference_goldens.sh on a machine with cargo) — the only thing that can
turn "parity unpinned" into "pinned".  Skipped while tests/golden/ref/ holds no proof: neither this image nor the GPU box has a
Rust toolchain (profiles/r2_gpu_box_probe.txt)."""
import glob
import json
import os

import pytest

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")
PROOFS = sorted(glob.glob(os.path.join(REF, "*.proof.json")))
STDIN = {"collatz": b"7\n"}
pytestmark = pytest.mark.skipif(not PROOFS, reason="no reference-made goldens: run tools/make_reference_goldens.sh where cargo exists")


@pytest.mark.parametrize("path", PROOFS)
def test_reference_proofs_pass_the_independent_verifier(path):
    import py_verifier
    assert py_verifier.verify(open(path).read(), 24)


@pytest.mark.gpu
@pytest.mark.parametrize("path", PROOFS)
def test_cuda_proof_equals_the_reference_proof(pkg, be, path):
    name = os.path.basename(path)[:-len(".proof.json")]
    code = open(os.path.join(os.path.dirname(REF), "programs", name + ".bf"), "rb").read()
    want = json.loads(open(path).read())
    got = json.loads(pkg.prove_brainfuck(be, code, STDIN.get(name, b""), 24).json())
    assert got["claim"] == want["claim"]
    for t in range(3):   # name the first stage that differs
        assert got["proof"]["commitments"][t] == want["proof"]["commitments"][t], f"commitment of tree {t}"
    assert got["interaction_claim"] == want["interaction_claim"]
    assert got["proof"]["commitments"][3] == want["proof"]["commitments"][3], "composition commitment"
    assert got["proof"]["sampled_values"] == want["proof"]["sampled_values"]
    assert got["proof"]["fri_proof"]["first_layer"]["commitment"] == want["proof"]["fri_proof"]["first_layer"]["commitment"]
    for k, (a, b) in enumerate(zip(got["proof"]["fri_proof"]["inner_layers"], want["proof"]["fri_proof"]["inner_layers"])):
        assert a["commitment"] == b["commitment"], f"FRI inner layer {k}"
    assert got == want
