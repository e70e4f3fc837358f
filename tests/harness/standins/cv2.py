"""Harness stand-in for ``cv2`` (not installed offline): load/load_blender.py:7 imports it for the ``half_res``
resize only.  INTER_AREA at an integer factor is a box filter.  TEST HARNESS ONLY."""
import numpy as np

INTER_AREA = 3


def resize(img, dsize, interpolation=INTER_AREA):
    w, h = dsize
    a = np.asarray(img)
    fh, fw = a.shape[0] // h, a.shape[1] // w
    assert fh * h == a.shape[0] and fw * w == a.shape[1], "stand-in cv2.resize: integer factors only"
    return a.reshape(h, fh, w, fw, *a.shape[2:]).mean(axis=(1, 3))
