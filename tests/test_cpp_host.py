"""The C++ host mirror (thesia_b200/host/thesia_host.hpp: SpecSetting, TrackList, TrackManager, encode_waveform_tile
with the reference's names and update rules) exercised by its own C++ test binary, which links only the C ABI."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BIN = ROOT / "tests" / "cpp" / "test_host_mirror"


def _build():
    r = subprocess.run(["make", "-C", str(ROOT / "tests" / "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_host_mirror_cpu():
    _build()
    r = subprocess.run([str(BIN), "--cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_gpu():
    _build()
    r = subprocess.run([str(BIN), "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    print(r.stdout)   # incl. the wall-clock line of the concurrent tile readers (std::thread, no interpreter lock)
