"""Unit checks of small host-side building blocks (inline strings of the records, host SHA-1 / record ids)."""
import os
import subprocess

from conftest import ROOT


def test_host_units(tmp_path):
    exe = str(tmp_path / "host_units")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-Wall",
                    "-Wno-missing-field-initializers", "-o", exe, os.path.join(ROOT, "tests", "units", "host_units.cpp"), "-lz"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host units ok" in r.stdout
