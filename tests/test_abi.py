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
        missing = declared - set(listed) - set(pkg.ABI_SYMBOLS)
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
