"""CPU: exhaustive proof-by-enumeration of the exact 3-operation division the circularity kernels use."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_three_operation_division_is_the_ieee_quotient(tmp_path):
    exe = tmp_path / "exact_div"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", os.path.join(ROOT, "tests", "helpers", "exact_div.c"), "-o", str(exe), "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "0", r.stdout
