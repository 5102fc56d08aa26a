"""The C-ABI library loads and exports every symbol include/stwo_cuda.h declares (no compute calls: no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "stwo_cuda.h")).read()
    declared = set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", hdr))
    lib = pkg.load_library()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in stwo_cuda.h but not exported"
    assert declared <= set(pkg.ABI_SYMBOLS) | {"sc_status"}


def test_sharded_and_prover_header_symbols_exported(pkg):
    """include/stwo_cuda_sharded.h and include/stwo_brainfuck.h: every declared entry point is exported too."""
    lib = pkg.load_library()
    for hdr_name, prefix, listed in (("stwo_cuda_sharded.h", "sc_", pkg.SHARDED_SYMBOLS), ("stwo_brainfuck.h", "sbf_", pkg.PROVER_SYMBOLS)):
        hdr = open(os.path.join(ROOT, "include", hdr_name)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)                       # prose in comments may mention other calls
        declared = set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, hdr))
        assert declared, hdr_name
        for sym in sorted(declared):
            assert hasattr(lib, sym), f"{sym} declared in {hdr_name} but not exported"
        missing = declared - set(listed) - set(pkg.ABI_SYMBOLS) - set(pkg.SHARDED_SYMBOLS)
        assert not missing, f"{hdr_name}: {sorted(missing)} not listed in the Python mirror"


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.BackendError):
        pkg.CudaBackend(0)


def test_product_does_not_import_oracle():
    """No product file includes, links, loads or calls anything under oracle/ (comments may mention it)."""
    bad = re.compile(r"#\s*include[^\n]*(orc_|oracle/)|dlopen[^\n]*orc|CDLL[^\n]*(orc|oracle)|import[^\n]*oracle|\borc_[a-z_]+\s*\(|\borc::")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "stwo-brainfuck_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc")):
                src = open(os.path.join(dirpath, f)).read()
                m = bad.search(src)
                assert not m, f"{f}: {m.group(0)}"


def _declared_in_headers():
    out = set()
    for hdr_name in ("stwo_cuda.h", "stwo_cuda_sharded.h", "stwo_brainfuck.h"):
        hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", hdr_name)).read(), flags=re.S)
        out |= set(re.findall(r"\b((?:sc|sbf)_[a-z0-9_]+)\s*\(", hdr))
    return out


def test_every_export_is_declared(pkg):
    """The reverse direction: libstwo_cuda.so exports no sc_* / sbf_* entry point that the headers do not declare."""
    import subprocess
    so = os.path.join(ROOT, "stwo-brainfuck_b200", "libstwo_cuda.so")
    nm = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if re.search(r" T (sc|sbf)_[a-z0-9_]+$", ln)}
    assert exported, "no exports found"
    assert exported == _declared_in_headers()


RUST_CRATE = os.path.join(ROOT, "bindings", "rust", "stwo-cuda-backend")


def test_rust_ffi_is_generated_from_the_headers():
    """bindings/rust/stwo-cuda-backend/src/ffi.rs is exactly what tools/gen_rust_ffi.py makes of include/*.h today, and
    declares every entry point of the three headers (the crate itself cannot be compiled in this image: no rustc)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(RUST_CRATE, "src", "ffi.rs")).read()
    assert text == gen.generate(), "ffi.rs is stale: run python tools/gen_rust_ffi.py"
    assert set(re.findall(r"pub fn ((?:sc|sbf)_[a-z0-9_]+)\(", text)) == _declared_in_headers()


def test_rust_crate_calls_only_declared_entry_points():
    """Every `ffi::name` the hand-written Rust modules use exists in ffi.rs with the same number of arguments as the call
    passes — a consistency check that stands in for the compiler this image lacks."""
    ffi = open(os.path.join(RUST_CRATE, "src", "ffi.rs")).read()
    arity = {m.group(1): (0 if not m.group(2).strip() else m.group(2).count(":"))
             for m in re.finditer(r"pub fn ((?:sc|sbf)_[a-z0-9_]+)\(([^)]*)\)", ffi)}
    consts = set(re.findall(r"pub const ([A-Z_]+):", ffi))
    types = set(re.findall(r"pub struct ([A-Za-z]+)", ffi))
    used = 0
    for f in sorted(os.listdir(os.path.join(RUST_CRATE, "src"))):
        if f == "ffi.rs" or not f.endswith(".rs"):
            continue
        src = open(os.path.join(RUST_CRATE, "src", f)).read()
        src = re.sub(r"//[^\n]*", "", src)
        mods = re.findall(r"^(?:pub )?mod ([a-z_]+);", src, flags=re.M)
        for m in mods:
            assert os.path.exists(os.path.join(RUST_CRATE, "src", m + ".rs")), f"{f}: mod {m} has no file"
        for m in re.finditer(r"(?<!std::)(?<!core::)\bffi::([A-Za-z_0-9]+)", src):
            name = m.group(1)
            assert name in arity or name in consts or name in types, f"{f}: ffi::{name} is not declared"
            if name in arity and src[m.end():m.end() + 1] == "(":
                depth, i, commas, nonempty = 0, m.end(), 0, False
                while True:                                   # count top-level commas of the call's argument list
                    c = src[i]
                    if c in "([{":
                        depth += 1
                    elif c in ")]}":
                        depth -= 1
                        if depth == 0:
                            break
                    elif c == "," and depth == 1:
                        commas += 1
                    elif depth >= 1 and not c.isspace():
                        nonempty = True
                    i += 1
                tail = src[m.end():i].rstrip()
                n_args = 0 if not nonempty else commas + (0 if tail.endswith(",") else 1)
                assert n_args == arity[name], f"{f}: {name} called with {n_args} arguments, declared with {arity[name]}"
                used += 1
    assert used >= 30


def test_reference_patch_applies():
    """bindings/rust/reference-cuda-feature.patch applies cleanly to the reference tree (only where that tree exists)."""
    import shutil, subprocess
    ref = "/root/reference"
    if not os.path.isdir(ref) or not shutil.which("patch"):
        pytest.skip("reference tree or patch(1) not available")
    patch = os.path.join(ROOT, "bindings", "rust", "reference-cuda-feature.patch")
    r = subprocess.run(["patch", "-p1", "--dry-run", "--batch", "-F0", "-d", ref, "-o", "/dev/null", "-i", patch], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_column_counts_match_the_reference_source(pkg):
    """`TraceColumn::count()` of the seven column enums, read from the reference's source text (only where that tree exists),
    against the counts the ABI documents (`COMPONENT_COLUMNS`), the C++ constants (air_ids.hpp) and the Rust `ComponentId`."""
    ref = "/root/reference/crates/brainfuck_prover/src/components"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not available")

    def count(rel):
        src = open(os.path.join(ref, rel)).read()
        m = re.search(r"fn count\(\) -> \(usize, usize\) \{\s*\((\d+), (\d+)\)", src)
        return int(m.group(1)), int(m.group(2))
    instr = count("processor/instructions/table.rs")
    jump = count("processor/instructions/jump/table.rs")
    want = [count("memory/table.rs"), count("instruction/table.rs"), count("program/table.rs"), count("processor/table.rs"), jump, jump] + \
        [instr] * 6 + [count("processor/instructions/end_of_execution/table.rs")]
    assert pkg.CudaBackend.COMPONENT_COLUMNS == want
    ids = open(os.path.join(ROOT, "stwo-brainfuck_b200", "csrc", "host", "air_ids.hpp")).read()
    nums = lambda name: [int(x) for x in re.search(name + r"\[N_COMPONENTS\] = \{([^}]*)\}", ids).group(1).split(",")]
    assert nums("N_MAIN_COLS") == [w[0] for w in want] and nums("N_LOGUP_COLS") == [w[1] for w in want]


def test_constraint_counts_match_the_reference_source():
    """N_CONSTRAINTS of air_ids.hpp = `eval.add_constraint(` calls + `eval.add_to_relation(` calls (one LogUp constraint per
    relation entry, no batching) in each component's `evaluate()` body of the reference (only where that tree exists)."""
    ref = "/root/reference/crates/brainfuck_prover/src/components"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not available")
    files = ["memory/component.rs", "instruction/component.rs", "program/component.rs", "processor/component.rs",
             "processor/instructions/jump/jump_if_not_zero_component.rs", "processor/instructions/jump/jump_if_zero_component.rs",
             "processor/instructions/input_component.rs", "processor/instructions/left_component.rs",
             "processor/instructions/minus_component.rs", "processor/instructions/output_component.rs",
             "processor/instructions/plus_component.rs", "processor/instructions/right_component.rs",
             "processor/instructions/end_of_execution/component.rs"]
    want = []
    for f in files:
        src = open(os.path.join(ref, f)).read()
        body = src[src.index("fn evaluate<E: EvalAtRow>"):]
        body = body[:body.index("\n    }\n")]
        body = re.sub(r"//[^\n]*", "", body)
        want.append(body.count("eval.add_constraint(") + body.count("eval.add_to_relation("))
    ids = open(os.path.join(ROOT, "stwo-brainfuck_b200", "csrc", "host", "air_ids.hpp")).read()
    got = [int(x) for x in re.search(r"N_CONSTRAINTS\[N_COMPONENTS\] = \{([^}]*)\}", ids).group(1).split(",")]
    assert got == want and sum(got) == 103
