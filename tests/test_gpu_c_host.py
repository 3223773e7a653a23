"""GPU: the plain-C host program (examples/frame_roundtrip.c) through LizardF_compressFrame / LizardF_decompress."""
import subprocess

import pytest

from tests.test_abi_cpu import build_c_host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [10, 41])
def test_c_host_frame_round_trip(tmp_path, level):
    exe = build_c_host(tmp_path)
    r = subprocess.run([exe, "32", str(level)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "round trip ok" in r.stdout
