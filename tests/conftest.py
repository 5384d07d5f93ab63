"""pytest configuration: registers the `gpu` marker and puts the product package and the oracle on sys.path.

`-m "not gpu"`: oracle vs the reference's known-answer values, host logic, C-ABI symbol export (no device needed).
`-m gpu`: parity of the CUDA path (through the C ABI) against the oracle, on a real B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kernelabstractions.jl_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import ka_oracle

    ka_oracle.build()
    return ka_oracle
