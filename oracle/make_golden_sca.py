"""Golden vectors of the spatial + colour alignment path from the UNMODIFIED reference modules (build container only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_sca.py

TEST INFRASTRUCTURE ONLY.  Shims: those of make_golden.py (cupy stub, pure-torch cost volume) plus `torch.lstsq`, which
models/loss/spatial_color_alignment.py:40 calls and torch >= 2.0 removed: `torch.lstsq(B, A)` is restated as
`torch.linalg.lstsq(A, B)` (same minimiser, argument order reversed).  `lpips` (imported by image_quality_v2.py:21, not
installed, not on this path) is stubbed so that AlignedL2 can be imported.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import sca_oracle as S  # noqa: E402
from oracle.make_golden import import_reference  # noqa: E402

CASES = [('sca_b2_192', 0, 1.0, 0, 2, 192), ('sca_b1_128_gain2', 1, 2.0, 1, 1, 128)]


def main():
    ref = import_reference()
    # torch 2.x keeps `torch.lstsq` only as a stub that raises "removed": replace it
    torch.lstsq = lambda B, A: types.SimpleNamespace(solution=torch.linalg.lstsq(A, B).solution)
    sys.modules.setdefault('lpips', types.ModuleType('lpips'))
    from models.loss.spatial_color_alignment import SpatialColorAlignment
    from models.loss.image_quality_v2 import AlignedL2
    outdir = os.path.join(ROOT, 'tests', 'golden')
    for name, wseed, gain, iseed, B, size in CASES:
        pwc = ref['PWCNet'](load_pretrained=False)
        pwc.load_state_dict(S.pwc_state_dict(wseed, gain), strict=True)
        pwc.eval()
        pred, gt, burst = S.make_sca_inputs(iseed, B, size)
        sca = SpatialColorAlignment(pwc, sr_factor=4)
        with torch.no_grad():
            pred_m, valid = sca(pred, gt, burst)
            flow = pwc(pred / (pred.max() + 1e-6), gt / (gt.max() + 1e-6))
            l2 = AlignedL2(pwc, sr_factor=4, boundary_ignore=16)(pred, gt, burst)
        np.savez_compressed(os.path.join(outdir, name + '.npz'),
                            meta=np.array([wseed, iseed, B, size], dtype=np.int64), gain=np.array([gain]),
                            pred_m=pred_m.numpy().astype(np.float32), valid=valid.numpy(),
                            flow=flow.numpy().astype(np.float32), aligned_l2=np.array([float(l2)]))
        print(name, 'valid frac', float(valid.float().mean()), 'max|flow|', float(flow.abs().max()), 'l2', float(l2))


if __name__ == '__main__':
    main()
