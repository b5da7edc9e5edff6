"""The smoothing kernel divides by 3 and by small valences with a reciprocal + one correction step (csrc/smooth.cuh div3_small)
instead of the IEEE division sequence.  That is only legitimate if it returns the correctly rounded quotient for every input it is
used on: checked here exhaustively on the CPU (hardware FMA = the same IEEE operations the device executes)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_short_division_is_exact_for_every_significand(tmp_path):
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("no hardware FMA on this host (the software fmaf would take minutes)")
    exe = str(tmp_path / "div_exact_check")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "div_exact_check.c"), "-lm"], check=True)
    r = subprocess.run([exe, "63"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("bad=0"), r.stdout[-2000:]
