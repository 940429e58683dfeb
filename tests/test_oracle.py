"""CPU: the oracle restatement (oracle/dbsr_oracle.py) against vectors produced by the reference's own
modules (oracle/make_golden.py -> tests/golden/*.npz).  fp32 tolerance: the oracle re-associates a few
sums (explicit gathers vs ATen kernels), so 2e-5 abs on pred in [0, 0.13] and 1e-4 px on the flow."""
import os

import numpy as np
import pytest
import torch

from oracle import dbsr_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + '.npz'))


def test_state_dict_contract(golden_dir):
    lines = open(os.path.join(golden_dir, 'state_dict_keys.txt')).read().split('\n')
    ref = [l.split(' ') for l in lines if l]
    spec = O.state_dict_spec()
    assert len(spec) == len(ref) == 231
    for (k, s), (rk, rs) in zip(spec, ref):
        assert k == rk
        assert 'x'.join(map(str, s)) == rs
    sd = O.make_state_dict(0)
    assert sum(v.numel() for v in sd.values()) == 13011237  # SURVEY.md 8c parameter count


def test_ops_against_reference(golden_dir):
    g = _load(golden_dir, 'ops')
    feat = torch.from_numpy(g['feat'])
    flow = torch.from_numpy(g['flow'])
    assert np.abs(O.warp(feat, flow).numpy() - g['warp']).max() < 2e-5
    assert np.abs(O.backwarp(feat, flow).numpy() - g['backwarp']).max() < 2e-5
    assert np.abs(O.resize_bilinear(feat, 64, 64).numpy() - g['interp_up']).max() < 1e-5
    assert np.abs(O.resize_bilinear(feat, 5, 7).numpy() - g['interp_down']).max() < 1e-5
    f1 = torch.from_numpy(g['corr_f1'])
    f2 = torch.from_numpy(g['corr_f2'])
    assert np.abs(O.correlation81(f1, f2).numpy() - g['corr']).max() < 1e-6
    assert np.abs(O.correlation81_bruteforce(f1[:1, :, :3, :4], f2[:1, :, :3, :4]).numpy()
                  - O.correlation81(f1[:1, :, :3, :4], f2[:1, :, :3, :4]).numpy()).max() < 1e-6
    m = torch.remainder(torch.tensor([-0.25, 1.75, -1e-9, 0.0, 3.0]), 1.0).numpy()
    assert np.array_equal(m, g['mod'])


def test_semantic_kats():
    """SURVEY.md Appendix D(7)."""
    g = torch.Generator().manual_seed(3)
    img = torch.rand(1, 2, 6, 7, generator=g)
    # integer flow == shift
    flow = torch.zeros(1, 2, 6, 7)
    flow[:, 0] = 1.0
    w = O.warp(img, flow)
    assert torch.allclose(w[..., :-1], img[..., 1:], atol=1e-6) and float(w[..., -1].abs().max()) == 0.0
    # backwarp: flow (W-1)/W shifts by exactly one pixel
    flow[:, 0] = (7 - 1.0) / 7
    bw = O.backwarp(img, flow)
    assert torch.allclose(bw[..., :-1], img[..., 1:], atol=1e-5)
    # pixel shuffle index map
    x = torch.arange(64 * 2 * 3, dtype=torch.float32).view(1, 64, 2, 3)
    ps = O.pixel_shuffle(x, 8)
    assert torch.equal(ps, torch.nn.functional.pixel_shuffle(x, 8))
    assert ps[0, 0, 8 * 1 + 3, 8 * 2 + 5] == x[0, 3 * 8 + 5, 1, 2]
    # deconv restatement
    xin = torch.randn(2, 5, 3, 4, generator=g)
    wt = torch.randn(5, 2, 4, 4, generator=g)
    b = torch.randn(2, generator=g)
    ref = torch.nn.functional.conv_transpose2d(xin, wt, b, stride=2, padding=1)
    assert torch.allclose(O.deconv4x4s2(xin, wt, b), ref, atol=1e-5)
    K = O.gauss_kernel3()
    assert abs(float(K[1, 1]) - 0.2042) < 1e-4 and abs(float(K[0, 0]) - 0.0751) < 1e-4


@pytest.mark.parametrize('name', ['tiny_b1n3_16x16', 'rect_b2n4_24x40', 'stress_b1n5_32x32', 'cfg1_b1n14_48x48'])
def test_forward_against_reference(golden_dir, name):
    g = _load(golden_dir, name)
    wseed, bseed, B, N, H, W = [int(v) for v in g['meta']]
    sd = O.make_state_dict(wseed, pwc_gain=float(g['pwc_gain'][0]))
    burst = O.make_burst(bseed, B, N, H, W)
    pred, aux = O.dbsr_forward(burst, sd)
    flow_tol = 1e-4 if float(g['pwc_gain'][0]) == 1.0 else 5e-3
    assert np.abs(aux['offsets'].numpy() - g['offsets']).max() < flow_tol
    if 'pred' in g:
        assert np.abs(pred.numpy() - g['pred']).max() < 2e-5
    else:
        assert np.abs(pred[:, :, ::2, ::2].numpy() - g['pred_sub']).max() < 2e-5
        assert abs(pred.double().sum().item() - g['pred_sum'][0]) < 1e-3 * abs(g['pred_sum'][0])
    fw = aux['fusion_weights']
    assert np.abs(fw[:, :, ::37, ::3, ::3].numpy() - g['fusion_weights_sub']).max() < 2e-5
    assert np.abs(fw.mean(dim=(2, 3, 4)).numpy() - g['fusion_weights_mean']).max() < 1e-5


def test_library_op_restatement_matches_explicit_oracle():
    """the timed CPU arm (same library-op sequence as the reference) == the explicit-arithmetic oracle"""
    sd = O.make_state_dict(1)
    burst = O.make_burst(4, 1, 3, 16, 24)
    p0, a0 = O.dbsr_forward(burst, sd)
    p1, a1 = O.dbsr_forward_fast(burst, sd)
    assert (p0 - p1).abs().max().item() < 2e-5
    assert (a0['offsets'] - a1['offsets']).abs().max().item() < 1e-4


@pytest.mark.parametrize('name', ['sca_b2_192', 'sca_b1_128_gain2'])
def test_sca_oracle_matches_reference_golden(golden_dir, name):
    """spatial + colour alignment: oracle/sca_oracle.py against vectors produced by the reference's
    SpatialColorAlignment / AlignedL2 themselves (oracle/make_golden_sca.py)"""
    from oracle import sca_oracle as S
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    wseed, iseed, B, size = [int(v) for v in g['meta']]
    sd = S.pwc_state_dict(wseed, float(g['gain'][0]))
    pred, gt, burst = S.make_sca_inputs(iseed, B, size)
    pm, valid, aux = S.spatial_color_alignment(pred, gt, burst, sd)
    assert np.abs(aux['flow'].numpy() - g['flow']).max() < 1e-4
    assert np.abs(pm.numpy() - g['pred_m']).max() < 1e-4
    assert (valid.numpy() != g['valid']).mean() < 1e-3
    assert 0.02 < valid.float().mean() < 0.9          # the mask is neither empty nor full
    l2 = float(S.aligned_l2(pred, gt, burst, sd, boundary_ignore=16))
    assert abs(l2 - float(g['aligned_l2'][0])) <= 1e-4 * float(g['aligned_l2'][0])


METRIC_CASES = ['metrics_rgb_b2_176', 'metrics_ragged_b3_97x131', 'metrics_gray_255_b1_200', 'metrics_signed_b2_180']


@pytest.mark.parametrize('name', METRIC_CASES)
def test_metrics_oracle_matches_reference_golden(golden_dir, name):
    """oracle/metrics_oracle.py against the values the reference's own msssim.py / image_quality_v2.py produced
    (oracle/make_golden_metrics.py): SSIM mean / per image / map, contrast term, MS-SSIM (incl. the small-map levels with
    a min(11, h, w) window), the boundary_ignore + valid-mask SSIM metric, PSNR.  Tolerance 2e-6 abs on the means (summation
    order of the 121-tap window), 10 x that on 0..255 images (E[x^2] - mu^2 cancels ~4 more digits there), 1e-4 dB on PSNR."""
    from oracle import metrics_oracle as M
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    seed, n, c, h, w, bi = [int(v) for v in g['meta']]
    noise, scale, offset = [float(v) for v in g['gen']]
    pred, gt = M.make_image_pair(seed, n, c, h, w, noise, scale, offset)
    valid = torch.rand(n, 1, h, w, generator=torch.Generator().manual_seed(int(g['valid_seed']))) > 0.3
    tol = 2e-6 if scale <= 2 else 2e-5
    mtol = 10 * tol
    s_mean, cs = M.ssim(pred, gt, full=True)
    assert abs(float(s_mean) - float(g['ssim_mean'])) <= tol and abs(float(cs) - float(g['cs'])) <= tol
    assert np.abs(M.ssim(pred, gt, size_average=False).numpy() - g['ssim_per_image']).max() <= tol
    smap = M.ssim(pred, gt, spatial_out=True)
    assert np.abs(smap[..., :24, :24].numpy() - g['ssim_map_corner']).max() <= mtol
    assert np.abs(smap[..., -16:, -16:].numpy() - g['ssim_map_tail']).max() <= mtol
    assert abs(float(M.ssim(pred, gt, val_range=1.0)) - float(g['ssim_val_range1'])) <= tol
    assert abs(float(M.msssim(pred, gt)) - float(g['msssim'])) <= tol
    assert abs(float(M.msssim(pred, gt, normalize=True)) - float(g['msssim_normalized'])) <= tol
    assert float(g['ssim_class']) == float(g['ssim_mean']) and float(g['msssim_class']) == float(g['msssim'])
    assert abs(float(M.ssim_metric(pred, gt, bi)) - float(g['iq_ssim_loss'])) <= tol
    assert abs(float(M.ssim_metric(pred, gt, bi, use_for_loss=False)) - float(g['iq_ssim'])) <= tol
    assert abs(float(M.ssim_metric(pred, gt, bi, use_for_loss=False, valid=valid)) - float(g['iq_ssim_valid'])) <= tol
    assert abs(float(M.ssim_metric(pred[0], gt[0], None, use_for_loss=False)) - float(g['iq_ssim_single'])) <= tol
    small = M.ssim(pred[..., :7, :9], gt[..., :7, :9], full=True)
    assert abs(float(small[0]) - float(g['small_ssim'])) <= tol and abs(float(small[1]) - float(g['small_cs'])) <= tol
    mv = max(scale, 1.0)
    assert abs(float(M.psnr(pred, gt, bi, mv)) - float(g['psnr'])) <= 1e-4
    assert np.abs(M.psnr_per_image(pred, gt, bi, mv).numpy() - g['psnr_per_image']).max() <= 1e-4


@pytest.mark.parametrize('name', ['camera_s0_64x96', 'camera_s1_48x40'])
def test_camera_oracle_matches_reference_golden(golden_dir, name):
    """oracle/camera_oracle.py against the reference's own data/camera_pipeline.py functions (oracle/make_golden_camera.py):
    the inverse pipeline within 1e-6 (torch.mm vs explicit sums), mosaic and the noisy burst bit for bit."""
    from oracle import camera_oracle as C
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    seed, h, w, n = [int(v) for v in g['meta']]
    image, rgb2cam, gains, burst_rgb, (shot, read) = C.make_inputs(seed, h, w, n)
    assert np.abs(C.unprocess(image, rgb2cam, *gains).numpy() - g['linear']).max() <= 1e-6
    assert np.abs(C.unprocess(image, rgb2cam, *gains, gamma=False).numpy() - g['linear_nogamma']).max() <= 1e-6
    assert np.array_equal(C.mosaic(burst_rgb).numpy(), g['raw'])
    torch.manual_seed(500 + seed)
    z = torch.FloatTensor(*g['raw'].shape).normal_()
    assert np.array_equal(C.mosaic_add_noise(burst_rgb, shot, read, z).numpy(), g['noisy'])


LRBURST_CASES = ['lrburst_default_b14_432', 'lrburst_shear_scale_b5_200x264', 'lrburst_factor2_b3_96x80']


@pytest.mark.parametrize('name', LRBURST_CASES)
def test_lrburst_oracle_is_bit_exact_against_reference_golden(golden_dir, name):
    """oracle/lrburst_oracle.py (OpenCV's fixed-point warpAffine / resize restated in integer arithmetic) against the bursts
    the reference's own `single2lrburst` produced with OpenCV (oracle/make_golden_lrburst.py): every byte equal; flow vectors
    within 1e-5 px (fp32 matrix products)."""
    from oracle import lrburst_oracle as L
    from oracle.make_golden_lrburst import make_image
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    seed, H, W, n, f, crop = [int(v) for v in g['meta']]
    burst, flow = L.single2lrburst(make_image(seed, H, W), list(g['t_mats']), f, None if crop < 0 else crop)
    assert np.array_equal(burst.numpy(), np.float32(g['burst_u8']) / np.float32(255.0))
    assert np.abs(flow.numpy() - g['flow']).max() <= 1e-5
    # the restated cv2.getRotationMatrix2D / get_tmat reproduce the recorded matrices to the last bits
    assert np.abs(L.get_tmat((H, W), (1.5, 1.5), 0.0, (0.0, 0.0), (1.0, 1.0)) - g['t_mats'][0]).max() <= 1e-12 or f != 4
