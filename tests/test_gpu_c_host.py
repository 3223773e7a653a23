"""GPU: the plain-C host programs -- examples/frame_roundtrip.c (LizardF_compressFrame / LizardF_decompress) and
examples/block_bench.c (the batch form of programs/bench.c: LizardB200_compress_blocks / _decompress_blocks)."""
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


@pytest.mark.parametrize("level", [10, 41])
def test_c_host_block_bench(tmp_path, level):
    exe = build_c_host(tmp_path, "block_bench")
    r = subprocess.run([exe, str(level), "64", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "512 blocks of 128 KiB" in r.stdout and "MB/s" in r.stdout
