import ctypes
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build():
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def pkg():
    m = importlib.import_module("stwo-brainfuck_b200")
    sys.modules["stwo_brainfuck_b200"] = m
    return m


@pytest.fixture(scope="session")
def orc():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def be(pkg):
    b = pkg.CudaBackend(0)
    yield b
    b.close()
