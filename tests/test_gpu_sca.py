"""GPU: the BurstSR spatial + colour alignment metric path -- drop-in `SpatialColorAlignment`,
`match_colors`, `AlignedL2`, `PSNR` on top of the sm_100a PWC-Net / warp kernels -- against the golden vectors produced by
the reference's own modules (oracle/make_golden_sca.py) and against the CPU oracle.

Tolerances: fp32 PWC-Net: aligned prediction <= 1e-4 max-abs, validity mask mismatch < 0.2 % of the pixels (it is a
threshold on a smooth error map), AlignedL2 within 1e-3 relative.  bf16 tensor-core PWC-Net: flow within 0.1 px,
AlignedL2 within 5 % (the metric moves with the flow)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sca_oracle as S  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def _pwc(sd, dev, precision='fp32'):
    from deep_rawburst_sr_b200.models.alignment.pwcnet import PWCNet
    net = PWCNet(load_pretrained=False)
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval().set_precision(precision)


@pytest.mark.parametrize('name', ['sca_b2_192', 'sca_b1_128_gain2'])
def test_sca_against_reference_golden(dev, golden_dir, name):
    from deep_rawburst_sr_b200.models.loss.image_quality_v2 import AlignedL2
    from deep_rawburst_sr_b200.models.loss.spatial_color_alignment import SpatialColorAlignment
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    wseed, iseed, B, size = [int(v) for v in g['meta']]
    sd = S.pwc_state_dict(wseed, float(g['gain'][0]))
    pred, gt, burst = S.make_sca_inputs(iseed, B, size)
    pwc = _pwc(sd, dev)
    sca = SpatialColorAlignment(pwc, sr_factor=4)
    assert sca.to(dev) is None          # reference signature
    pm, valid = sca(pred.to(dev), gt.to(dev), burst.to(dev))
    assert pm.shape == pred.shape and valid.shape == (B, 1, size, size) and valid.dtype == torch.bool
    flow = pwc(pred.to(dev) / (pred.max() + 1e-6), gt.to(dev) / (gt.max() + 1e-6))
    assert np.abs(flow.cpu().numpy() - g['flow']).max() < 1e-3
    assert np.abs(pm.cpu().numpy() - g['pred_m']).max() <= 1e-4
    assert (valid.cpu().numpy() != g['valid']).mean() < 2e-3
    l2 = float(AlignedL2(pwc, sr_factor=4, boundary_ignore=16)(pred.to(dev), gt.to(dev), burst.to(dev)))
    ref = float(g['aligned_l2'][0])
    assert abs(l2 - ref) <= 1e-3 * ref, (l2, ref)
    with pytest.raises(NotImplementedError):
        sca(pred, gt, burst)            # CPU tensors are refused, there is no CPU path


def test_sca_bf16_alignment_net_and_psnr(dev):
    """PWC-Net on the tensor cores for the output-resolution alignment: flow within 0.1 px of the fp32 oracle, the metric
    within 5 %; PSNR with a validity mask follows the reference formula (image_quality_v2.py:47-66, 75-101)"""
    from deep_rawburst_sr_b200.models.loss.image_quality_v2 import PSNR, AlignedL2
    sd = S.pwc_state_dict(0)
    pred, gt, burst = S.make_sca_inputs(5, 2, 192)
    ref_pm, ref_valid, aux = S.spatial_color_alignment(pred, gt, burst, sd)
    ref_l2 = float(S.aligned_l2(pred, gt, burst, sd, boundary_ignore=16))
    pwc = _pwc(sd, dev, 'bf16')
    flow = pwc(pred.to(dev) / (pred.max() + 1e-6), gt.to(dev) / (gt.max() + 1e-6))
    assert (flow.cpu() - aux['flow']).abs().max() < 0.1
    l2 = float(AlignedL2(pwc, sr_factor=4, boundary_ignore=16)(pred.to(dev), gt.to(dev), burst.to(dev)))
    assert abs(l2 - ref_l2) <= 0.05 * ref_l2, (l2, ref_l2)
    # PSNR over the valid pixels, per image then averaged
    b = 16
    want = []
    for p, t, v in zip(ref_pm, gt, ref_valid):
        mse = (((p - t) ** 2)[..., b:-b, b:-b] * v[..., b:-b, b:-b].float()).sum() / (v[..., b:-b, b:-b].float().sum() * 3 + 1e-12)
        want.append(-10.0 * torch.log10(mse))
    got = PSNR(boundary_ignore=b)(ref_pm.to(dev), gt.to(dev), ref_valid.to(dev))
    assert abs(float(got) - float(sum(want) / len(want))) < 1e-3


def test_sca_burstsr_size(dev):
    """the BurstSR evaluation shape (evaluation/burstsr/compute_score.py:73,118): 640^2 prediction, 80^2 RAW burst"""
    from deep_rawburst_sr_b200.models.loss.spatial_color_alignment import SpatialColorAlignment
    pred, gt, burst = S.make_sca_inputs(9, 1, 640)
    sca = SpatialColorAlignment(_pwc(S.pwc_state_dict(0), dev, 'bf16'), sr_factor=4)
    sca.to(dev)
    pm, valid = sca(pred.to(dev), gt.to(dev), burst.to(dev))
    assert pm.shape == (1, 3, 640, 640) and valid.shape == (1, 1, 640, 640)
    assert torch.isfinite(pm).all() and 0.3 < float(valid.float().mean()) < 0.95
