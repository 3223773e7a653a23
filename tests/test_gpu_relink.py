"""GPU: the relink of INTEGRATION.md section 1 as a test.  oracle/_ref/relinked_frame is the reference's UNMODIFIED frame
layer (lib/lizard_frame.c + xxhash.c) linked against liblizard_b200.so, so LizardF_compressFrame / LizardF_decompress of
the reference drive our Lizard_createStream / Lizard_compress_extState / Lizard_decompress_safe one block per call.
The frame it writes must be byte-identical to the frame of the pure reference (parity build), no block may be stored
raw, and its own decode must give the input back."""
import os
import subprocess

import pytest

import lizard_b200 as lz
from tests import refs

pytestmark = pytest.mark.gpu
EXE = os.path.join(refs.REF_DIR, "relinked_frame")


@pytest.mark.parametrize("level", [10, 21, 41])
def test_reference_frame_layer_over_the_gpu_codec(tmp_path, level):
    ref = refs.ref_parity()
    if ref is None or not os.path.exists(EXE):
        pytest.skip("oracle/_ref not built")
    lz.bind_frame_api(ref)
    out = os.path.join(str(tmp_path), "f.liz")
    mib = 4
    r = subprocess.run([EXE, str(level), str(mib), out, "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "%d blocks, 0 stored raw; round trip ok" % (mib * 8) in r.stdout, r.stdout
    data = lz.datagen(mib << 20)
    want = lz.frame_compress(ref, data, lz.make_prefs(level, 1, True, True, 0))
    assert open(out, "rb").read() == want
