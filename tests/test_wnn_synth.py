"""Native witness synthesis (csrc/wnn_synth.cu, zg_wnn_* in the C ABI; host code) against the Python front-end that
mirrors the reference's gadgets (zg_b200/plonk/gadgets.py): the six advice columns must agree cell for cell, and the
class scores must be Wnn::predict's (tests/integration_test.rs snapshots).  Runs without a GPU."""
import os
import time

import numpy as np
import pytest

import bn254
from zg_b200.io import load_grayscale_image, load_wnn, synthetic_image, synthetic_wnn
from zg_b200.lib import NativeSynthesizer, ZgError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _python_columns(wnn, img, k):
    _, asm = wnn.synthesize(img, k)
    return [bn254.fr_to_limbs(col) for col in asm.advice], asm.usable_rows


@pytest.mark.parametrize("fname,k,expected", [
    ("model_28input_256entry_1hash_1bpi.hdf5", 14, [9, 6, 13, 10, 17, 10, 9, 26, 11, 16]),       # integration_test.rs:13-20
    ("model_28input_1024entry_2hash_2bpi.hdf5", 15, [17, 13, 25, 27, 29, 21, 15, 55, 27, 32]),    # :30-37
    ("model_28input_2048entry_2hash_3bpi.hdf5", 15, [29, 21, 40, 47, 45, 41, 28, 82, 35, 66]),    # :47-54
])
def test_native_witness_matches_python_front_end(fname, k, expected):
    wnn = load_wnn(os.path.join(GOLD, fname))
    ns = NativeSynthesizer(wnn)
    for img in (load_grayscale_image(os.path.join(GOLD, "example_image_7.png")), synthetic_image(3), np.zeros((28, 28), np.uint8),
                np.full((28, 28), 255, np.uint8)):
        ref, usable = _python_columns(wnn, img, k)
        cols, scores = ns.synthesize(img, k, usable)
        assert scores == wnn.predict(img)
        for c in range(6):
            bad = np.nonzero((cols[c] != ref[c]).any(axis=1))[0]
            assert bad.size == 0, "advice column %d differs first at row %d" % (c, bad[0])
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    assert ns.synthesize(img, k, usable)[1] == expected


def test_native_witness_large_shape_and_speed():
    wnn = synthetic_wnn()
    ns = NativeSynthesizer(wnn)
    img = synthetic_image(0)
    ref, usable = _python_columns(wnn, img, 17)
    t0 = time.perf_counter()
    cols, scores = ns.synthesize(img, 17, usable)
    dt = time.perf_counter() - t0
    assert scores == wnn.predict(img)
    for c in range(6):
        assert (cols[c] == ref[c]).all(), c
    print("native synthesis, k = 17 shape: %.1f ms" % (dt * 1e3))
    assert dt < 1.0


def test_not_enough_rows_is_reported():
    wnn = load_wnn(os.path.join(GOLD, "model_28input_1024entry_2hash_2bpi.hdf5"))
    ns = NativeSynthesizer(wnn)
    with pytest.raises(ZgError) as e:
        ns.synthesize(np.zeros((28, 28), np.uint8), 14, (1 << 14) - 6)      # the small model needs k = 15 (src/lib.rs:48-51)
    assert e.value.code == -5 and "not enough rows" in str(e.value)
