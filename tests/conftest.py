import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


# several ranks of a sharded register may live in this process (tests of the sharded path on one GPU): give every
# stream its own hardware queue so that one rank's stream wait can never sit in front of another rank's kernels
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """the product package (directory `quromorphic.jl_b200/`, importable name quromorphic_jl_b200)"""
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(params=["records", "compiled_ptx"])
def emu_mode(request):
    """the host emulation of the fused-pass programs runs twice: over the encoded records (what the interpreter kernel reads)
    and over the straight-line PTX text generated for the compiled pass kernels (csrc/jit_codegen.hpp, executed by the
    PTX-subset interpreter of tests/native/ptx_interp.hpp)"""
    import emu
    emu.set_jit(request.param == "compiled_ptx")
    yield request.param
    emu.set_jit(0)
