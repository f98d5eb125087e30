"""CPU suite: the host-side staging converters of the class call (schicexplorer_b200/csrc/hostconv.cpp) -- the only host
arithmetic of the product path -- compiled with g++ together with a stand-alone checker and run here."""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_hostconv_narrowing_contract(tmp_path):
    exe = str(tmp_path / "hostconv_check")
    src = [os.path.join(ROOT, "tests", "hostconv_check.cpp"), os.path.join(ROOT, "schicexplorer_b200", "csrc", "hostconv.cpp")]
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, *src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0 and "HOSTCONV_OK" in r.stdout, r.stdout + r.stderr
