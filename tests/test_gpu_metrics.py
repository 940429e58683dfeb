"""GPU: the evaluation metrics -- drop-in `msssim.ssim / msssim / SSIM / MSSSIM`,
`image_quality_v2.SSIM`, `PSNR` on the fused sm_100a kernels (`dbsr_ssim`, `dbsr_avgpool2_pair`, `dbsr_mse_per_image`) --
against the values produced by the reference's own modules (oracle/make_golden_metrics.py) and against the CPU oracle.

Tolerances (fp32 path; the kernel applies the 11 x 11 window separably, the reference as one 121-tap window, so the moments
differ in the last bits and `E[x^2] - mu^2` amplifies that on flat patches): means (SSIM, contrast term, MS-SSIM, metric
values) <= 5e-6 abs, 5e-5 on 0..255 images, 4 x that for MS-SSIM (a product over five levels); SSIM map <= 1e-3 pointwise and <= 1e-5 on average; PSNR <= 1e-4 dB.
Determinism: every reduction has a fixed order, so repeated calls and a batch vs its images alone are bit-identical."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import metrics_oracle as M  # noqa: E402

CASES = ['metrics_rgb_b2_176', 'metrics_ragged_b3_97x131', 'metrics_gray_255_b1_200', 'metrics_signed_b2_180']


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def _case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    seed, n, c, h, w, bi = [int(v) for v in g['meta']]
    noise, scale, offset = [float(v) for v in g['gen']]
    pred, gt = M.make_image_pair(seed, n, c, h, w, noise, scale, offset)
    valid = torch.rand(n, 1, h, w, generator=torch.Generator().manual_seed(int(g['valid_seed']))) > 0.3
    return g, pred, gt, valid, bi, scale


@pytest.mark.parametrize('name', CASES)
def test_ssim_family_against_reference_golden(dev, golden_dir, name):
    from deep_rawburst_sr_b200.models.loss import msssim as ms
    from deep_rawburst_sr_b200.models.loss.image_quality_v2 import SSIM
    g, pred, gt, valid, bi, scale = _case(golden_dir, name)
    p, q = pred.to(dev), gt.to(dev)
    tol = 5e-6 if scale <= 2 else 5e-5
    s_mean, cs = ms.ssim(p, q, full=True)
    assert abs(float(s_mean) - float(g['ssim_mean'])) <= tol and abs(float(cs) - float(g['cs'])) <= tol
    assert np.abs(ms.ssim(p, q, size_average=False).cpu().numpy() - g['ssim_per_image']).max() <= tol
    smap = ms.ssim(p, q, spatial_out=True).cpu()
    ref_map, _ = M.ssim_map(pred, gt)
    assert smap.shape == ref_map.shape
    assert np.abs(smap[..., :24, :24].numpy() - g['ssim_map_corner']).max() <= 1e-3
    assert np.abs(smap[..., -16:, -16:].numpy() - g['ssim_map_tail']).max() <= 1e-3
    d = (smap - ref_map).abs()
    assert float(d.max()) <= 1e-3 and float(d.mean()) <= 1e-5, (float(d.max()), float(d.mean()))
    assert abs(float(ms.ssim(p, q, val_range=1.0)) - float(g['ssim_val_range1'])) <= tol
    assert abs(float(ms.msssim(p, q)) - float(g['msssim'])) <= 4 * tol            # product of five levels' terms
    assert abs(float(ms.msssim(p, q, normalize=True)) - float(g['msssim_normalized'])) <= 4 * tol
    assert abs(float(ms.SSIM()(p, q)) - float(g['ssim_class'])) <= tol
    assert abs(float(ms.MSSSIM()(p, q)) - float(g['msssim_class'])) <= 4 * tol
    # the reference's own window tensor is accepted, any other window is refused
    assert abs(float(ms.ssim(p, q, window=ms.create_window(11, p.shape[1]).to(dev))) - float(g['ssim_mean'])) <= tol
    with pytest.raises(NotImplementedError):
        ms.ssim(p, q, window=torch.ones(p.shape[1], 1, 11, 11, device=dev) / 121)
    assert abs(float(SSIM(boundary_ignore=bi)(p, q)) - float(g['iq_ssim_loss'])) <= tol
    assert abs(float(SSIM(boundary_ignore=bi, use_for_loss=False)(p, q)) - float(g['iq_ssim'])) <= tol
    assert abs(float(SSIM(boundary_ignore=bi, use_for_loss=False)(p, q, valid.to(dev))) - float(g['iq_ssim_valid'])) <= tol
    assert abs(float(SSIM(boundary_ignore=None, use_for_loss=False)(p[0], q[0])) - float(g['iq_ssim_single'])) <= tol
    small = ms.ssim(p[..., :7, :9].contiguous(), q[..., :7, :9].contiguous(), full=True)      # 7-tap window, one row of positions
    assert abs(float(small[0]) - float(g['small_ssim'])) <= tol and abs(float(small[1]) - float(g['small_cs'])) <= tol
    with pytest.raises(NotImplementedError):
        ms.ssim(pred, gt)                       # CPU tensors are refused, there is no CPU path


@pytest.mark.parametrize('name', CASES)
def test_psnr_against_reference_golden(dev, golden_dir, name):
    from deep_rawburst_sr_b200.models.loss.image_quality_v2 import PSNR
    g, pred, gt, _, bi, scale = _case(golden_dir, name)
    m = PSNR(boundary_ignore=bi, max_value=max(scale, 1.0))
    p, q = pred.to(dev), gt.to(dev)
    assert abs(float(m(p, q)) - float(g['psnr'])) <= 1e-4
    assert np.abs(m.psnr_per_image(p, q).cpu().numpy() - g['psnr_per_image']).max() <= 1e-4
    assert abs(float(PSNR(boundary_ignore=None)(p, q)) - float(M.psnr(pred, gt, None))) <= 1e-4
    # identical images: the reference drops the inf and returns 0 when nothing is left (image_quality_v2.py:97-99)
    assert float(m(p, p.clone())) == 0.0
    mixed_q = q.clone()
    mixed_q[0] = p[0]
    if p.shape[0] > 1:
        assert abs(float(m(p, mixed_q)) - float(g['psnr_per_image'][1:].mean())) <= 1e-4


def test_avgpool_pair_is_exact(dev):
    from deep_rawburst_sr_b200 import ops
    for shape in [(2, 3, 64, 48), (1, 1, 37, 53), (3, 2, 2, 2)]:
        a = torch.rand(*shape, generator=torch.Generator().manual_seed(1))
        b = torch.rand(*shape, generator=torch.Generator().manual_seed(2))
        oa, ob = ops.avgpool2_pair(a.to(dev), b.to(dev))
        assert torch.equal(oa.cpu(), M.avg_pool2(a)) and torch.equal(ob.cpu(), M.avg_pool2(b))


def test_metrics_full_size_properties(dev):
    """BASELINE configs[1] output size (32 x 3 x 384 x 384): SSIM(x, x) == 1 and the contrast term == 1 exactly, symmetry in the
    arguments, run-to-run and batch-composition bit-identity, PSNR consistent with the MSE of a known perturbation, and
    agreement with the CPU oracle on a few images of the batch."""
    from deep_rawburst_sr_b200.models.loss import msssim as ms
    from deep_rawburst_sr_b200.models.loss.image_quality_v2 import PSNR, SSIM
    pred, gt = M.make_image_pair(11, 32, 3, 384, 384, 0.03)
    p, q = pred.to(dev), gt.to(dev)
    s_pp = ms.ssim(p, p, size_average=False, full=True)
    assert torch.all(s_pp[0] == 1.0) and float(s_pp[1]) == 1.0
    a, b = ms.ssim(p, q, size_average=False), ms.ssim(q, p, size_average=False)
    assert torch.allclose(a, b, rtol=0, atol=1e-6)
    assert torch.equal(a, ms.ssim(p, q, size_average=False))
    assert torch.equal(a[5:9], ms.ssim(p[5:9].contiguous(), q[5:9].contiguous(), size_average=False))
    ref = M.ssim(pred[:3], gt[:3], size_average=False)
    assert float((a[:3].cpu() - ref).abs().max()) <= 5e-6
    assert abs(float(ms.msssim(p[:2].contiguous(), q[:2].contiguous())) - float(M.msssim(pred[:2], gt[:2]))) <= 2e-5
    m = SSIM(boundary_ignore=40, use_for_loss=False)
    assert abs(float(m(p[:3].contiguous(), q[:3].contiguous())) - float(M.ssim_metric(pred[:3], gt[:3], 40, False))) <= 5e-6
    shifted = (p + 0.125).contiguous()
    per = PSNR(boundary_ignore=40).psnr_per_image(p, shifted)
    assert torch.allclose(per, torch.full_like(per, -10.0 * np.log10(0.125 ** 2)), rtol=0, atol=1e-4)
    per = PSNR(boundary_ignore=40).psnr_per_image(p, q)
    assert float((per.cpu() - M.psnr_per_image(pred, gt, 40)).abs().max()) <= 1e-4
    assert torch.equal(per, PSNR(boundary_ignore=40).psnr_per_image(p, q))
