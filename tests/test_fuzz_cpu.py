"""A bounded slice of the restatement-vs-restatement campaign (tests/cpu_fuzz_pyref.py; the long runs are logged in
profiles/r1_fuzz_pyref_vs_oracle.txt): the pure-Python restatement and the C++ oracle must agree on bit patterns."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_python_and_cpp_restatements_agree_on_random_cases(oracle):
    out = subprocess.run([sys.executable, os.path.join(HERE, "cpu_fuzz_pyref.py"), "3000", "201"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "mismatches 0" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
