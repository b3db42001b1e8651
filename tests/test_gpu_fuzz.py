"""A bounded slice of the differential campaign (tests/gpu_fuzz.py; the long runs are logged in profiles/r1_fuzz_campaign.txt)
as part of the GPU suite: fresh seeds, every method and output kind, bit patterns compared against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("args", [["400", "101"], ["60", "102", "nvrtc"], ["300", "103", "fast"]], ids=["builtin", "nvrtc", "fast-sanity"])
def test_differential_fuzz_slice(deq, args):
    out = subprocess.run([sys.executable, os.path.join(HERE, "gpu_fuzz.py")] + args, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "mismatches 0" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
