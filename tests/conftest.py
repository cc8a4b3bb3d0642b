import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# new device allocations of the library are filled with 0xFF bytes in the tests: reads of never-written memory must not pass
# by the luck of zero pages (csrc/ctx.cuh: poison_allocations)
os.environ.setdefault("TRAJOPT_B200_POISON", "1")
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_ref():
    """the compiled reference (oracle/_ref); building it needs /root/reference, so it may be absent"""
    from oracle import oracle_api as oa
    if not os.path.exists(oa.REF_SO):
        if os.path.isdir("/root/reference/HighOrderCCD"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return oa.RefOracle()


@pytest.fixture(scope="session")
def oracle_any():
    from oracle import oracle_api as oa
    try:
        return oa.get()
    except FileNotFoundError:
        pytest.skip("no oracle library built")


@pytest.fixture(scope="session")
def hostsim():
    import ctypes
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libhostsim.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", d])
    return ctypes.CDLL(so)
