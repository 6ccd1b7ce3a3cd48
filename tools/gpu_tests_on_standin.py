"""Runs GPU-marked test modules against tests/fake_engine.py (the test-only stand-in for the C ABI) -- a DRY RUN of their own
logic and of the host mirrors on a machine without a GPU, before spending GPU time on them.  It proves nothing about the device
code (the stand-in computes group operations with the oracle and runs a few kernel bodies on the host emulation); tests that need
entry points the stand-in does not implement fail with AttributeError.
usage: python tools/gpu_tests_on_standin.py tests/test_gpu_verifier.py tests/test_gpu_multiopen.py ... [-k expr]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fake_engine  # noqa: E402

cm = fake_engine.installed()
cm.__enter__()
import pytest  # noqa: E402

sys.exit(pytest.main(sys.argv[1:] + ["-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"]))
