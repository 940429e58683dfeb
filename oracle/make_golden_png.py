"""Golden file for the result writer / reader (test infrastructure; run in the build container, needs OpenCV):
the reference stores predictions with `cv2.imwrite(path, uint16 HxWx3)` (evaluation/synburst/save_results.py:65-68) and reads
them with `cv2.imread(path, cv2.IMREAD_UNCHANGED)` (compute_score.py:101).  This script writes a deterministic 14-bit image with
exactly that call, so the tests can check `read_png16` against a file the reference's writer produced, without OpenCV.

    python oracle/make_golden_png.py
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_image(h=40, w=56):
    """deterministic 14-bit test image: smooth ramps (adaptive filters pick Sub / Up / Average / Paeth rows) + a hashed texture"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    tex = (yy * 7919 + xx * 104729 + (yy * xx) * 31) % 97
    img = np.stack([(yy * 400 + xx * 3 + tex) % 16385, (xx * 290 + tex * 5) % 16385, ((yy + xx) * 199 + tex) % 16385], axis=-1)
    return img.astype(np.uint16)


def golden_image4(h=24, w=32):
    """deterministic 14-bit four-channel frame (R, G, G, B planes of a packed mosaic)"""
    base = golden_image(h, w).astype(np.int64)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    fourth = (yy * 613 + xx * 389 + (yy ^ xx) * 17) % 16385
    return np.concatenate([base, fourth[:, :, None]], axis=-1).astype(np.uint16)


if __name__ == '__main__':
    import cv2          # only the generator needs OpenCV; the tests import golden_image / golden_image4
    out = os.path.join(ROOT, 'tests', 'golden', 'pred_u16_cv2.png')
    img = golden_image()
    assert cv2.imwrite(out, img)
    back = cv2.imread(out, cv2.IMREAD_UNCHANGED)
    assert back.dtype == np.uint16 and np.array_equal(back, img)
    print('wrote', out, os.path.getsize(out), 'bytes')
    # a packed RGGB burst frame as the SyntheticBurst validation set stores it: uint16 HxWx4 through the same cv2.imwrite
    # (dataset/synthetic_burst_val_set.py:44 reads it back with cv2.IMREAD_UNCHANGED)
    out4 = os.path.join(ROOT, 'tests', 'golden', 'burst_u16x4_cv2.png')
    img4 = golden_image4()
    assert cv2.imwrite(out4, img4)
    back4 = cv2.imread(out4, cv2.IMREAD_UNCHANGED)
    assert back4.dtype == np.uint16 and np.array_equal(back4, img4)
    print('wrote', out4, os.path.getsize(out4), 'bytes')
