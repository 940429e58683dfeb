"""GPU: inverse camera pipeline of the synthetic burst generator (`dbsr_unprocess_rgb`,
`dbsr_mosaic_noise`) against the values of the reference's own data/camera_pipeline.py functions
(oracle/make_golden_camera.py) and the CPU oracle.  Tolerances: the inverse pipeline <= 2e-6 abs on [0, 1] (CUDA powf / asinf /
sinf against the host's libm, a few ulp); mosaic and mosaic + noise + clamp are bit-exact (IEEE mul / add / sqrt only)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import camera_oracle as C  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


@pytest.mark.parametrize('name', ['camera_s0_64x96', 'camera_s1_48x40'])
def test_camera_pipeline_against_reference_golden(dev, golden_dir, name):
    from deep_rawburst_sr_b200.data import camera_pipeline as cp
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    seed, h, w, n = [int(v) for v in g['meta']]
    image, rgb2cam, gains, burst_rgb, (shot, read) = C.make_inputs(seed, h, w, n)
    lin = cp.unprocess(image.to(dev), rgb2cam, *gains)
    assert lin.shape == image.shape and np.abs(lin.cpu().numpy() - g['linear']).max() <= 2e-6
    lin2 = cp.unprocess(image.to(dev), rgb2cam, *gains, gamma=False)
    assert np.abs(lin2.cpu().numpy() - g['linear_nogamma']).max() <= 2e-6
    # a batch sharing the parameters equals its images alone
    both = cp.unprocess(torch.stack([image, image.flip(-1)]).to(dev), rgb2cam, *gains)
    assert torch.equal(both[0], lin) and torch.equal(both[1], cp.unprocess(image.flip(-1).contiguous().to(dev), rgb2cam, *gains))
    assert np.array_equal(cp.mosaic(burst_rgb.to(dev)).cpu().numpy(), g['raw'])
    assert np.array_equal(cp.mosaic(burst_rgb[0].to(dev)).cpu().numpy(), g['raw'][0])
    torch.manual_seed(500 + seed)
    z = torch.FloatTensor(*g['raw'].shape).normal_()
    noisy = cp.mosaic_add_noise(burst_rgb.to(dev), shot, read, noise=z.to(dev))
    assert np.array_equal(noisy.cpu().numpy(), g['noisy'])
    drawn = cp.mosaic_add_noise(burst_rgb.to(dev), shot, read)           # device-side draw: same statistics, in range
    assert drawn.shape == noisy.shape and float(drawn.min()) >= 0.0 and float(drawn.max()) <= 1.0
    assert abs(float(drawn.mean()) - float(noisy.mean())) < 5e-3
    with pytest.raises(NotImplementedError):
        cp.unprocess(image, rgb2cam, *gains)                              # CPU tensors are refused


def test_camera_pipeline_full_size(dev):
    """generator-sized inputs (a 448 x 448 crop -> 14-frame 96 x 96 RGB burst -> 48 x 48 packed RAW, default_synthetic.py) against
    the CPU oracle; the burst it produces is a valid network input (packed RGGB in [0, 1])"""
    from deep_rawburst_sr_b200.data import camera_pipeline as cp
    image, rgb2cam, gains, _, (shot, read) = C.make_inputs(3, 448, 448, 1)
    lin = cp.unprocess(image.to(dev), rgb2cam, *gains)
    assert float((lin.cpu() - C.unprocess(image, rgb2cam, *gains)).abs().max()) <= 2e-6
    burst_rgb = torch.nn.functional.avg_pool2d(lin, 4)[None].expand(14, -1, -1, -1)[..., :96, :96].contiguous()
    z = torch.randn(14, 4, 48, 48, generator=torch.Generator().manual_seed(1))
    raw = cp.mosaic_add_noise(burst_rgb, shot, read, noise=z.to(dev))
    ref = C.mosaic_add_noise(burst_rgb.cpu(), shot, read, z)
    d = (raw.cpu() - ref).abs()
    # (against the reference-generated goldens above the kernel is bit-exact; here allow the last bit)
    assert float(d.max()) <= 1.2e-7 and float((d > 0).float().mean()) < 1e-2, (float(d.max()), float((d > 0).float().mean()))
    assert raw.shape == (14, 4, 48, 48)


@pytest.mark.parametrize('name', ['lrburst_default_b14_432', 'lrburst_shear_scale_b5_200x264', 'lrburst_factor2_b3_96x80'])
def test_single2lrburst_is_bit_exact_against_opencv_golden(dev, golden_dir, name):
    """`dbsr_single2lrburst` (uint8 quantisation + cv2.warpAffine + border crop + cv2.resize + / 255, fused) against the bursts
    the reference's own `single2lrburst` produced with OpenCV: BIT-EXACT (byte work); flow vectors within 1e-4 px.  Then the
    drop-in `single2lrburst` under the same `random` seed (its own transform sampling) reproduces the same burst."""
    import random
    from deep_rawburst_sr_b200.data import synthetic_burst_generation as G
    from oracle.make_golden_lrburst import CASES, make_image
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    seed, H, W, n, f, crop = [int(v) for v in g['meta']]
    image = make_image(seed, H, W)
    ref = np.float32(g['burst_u8']) / np.float32(255.0)
    burst, flow = G.lrburst_from_transforms(image.to(dev), list(g['t_mats']), f, None if crop < 0 else crop)
    assert burst.shape == ref.shape and np.array_equal(burst.cpu().numpy(), ref)
    assert np.abs(flow.cpu().numpy() - g['flow']).max() <= 1e-4
    params = dict([c for c in CASES if c[0] == name][0][7])
    if crop >= 0:
        params['border_crop'] = crop
    random.seed(seed)
    burst2, flow2 = G.single2lrburst(image.to(dev), n, downsample_factor=f, transformation_params=params)
    # (the sampled matrices equal the recorded ones to 1e-15: a pixel could only differ on an exact rounding tie)
    assert float((burst2 != burst).float().mean()) < 1e-5 and np.abs(flow2.cpu().numpy() - g['flow']).max() <= 1e-4
    with pytest.raises(NotImplementedError):
        G.single2lrburst(image.to(dev), n, downsample_factor=f, transformation_params=params, interpolation_type='lanczos')


def test_rgb2rawburst_end_to_end(dev):
    """the whole generator on the device (default_synthetic.py settings: 14 frames, x4, +-24 px, +-1 deg, border crop 24) against
    the CPU oracle fed with the same sampled parameters; the burst is a valid DBSR input shape"""
    import random
    from deep_rawburst_sr_b200.data import synthetic_burst_generation as G
    from oracle import lrburst_oracle as L
    from oracle.make_golden_lrburst import make_image
    image = make_image(7, 432, 432)
    params = {'max_translation': 24.0, 'max_rotation': 1.0, 'max_shear': 0.0, 'max_scale': 0.0, 'border_crop': 24}
    random.seed(3)
    torch.manual_seed(3)
    raw, lin, burst_rgb, flow, meta = G.rgb2rawburst(image.to(dev), 14, 4, dict(params))
    assert raw.shape == (14, 4, 48, 48) and burst_rgb.shape == (14, 3, 96, 96) and flow.shape == (14, 2, 96, 96)
    assert float(raw.min()) >= 0.0 and float(raw.max()) <= 1.0 and float(flow[0].abs().max()) == 0.0
    # replay on the oracle with the same draws
    random.seed(3)
    torch.manual_seed(3)
    rgb2cam = G.rgb2raw.random_ccm()
    gains = G.rgb2raw.random_gains()
    assert torch.equal(rgb2cam, meta['rgb2cam']) and gains == (meta['rgb_gain'], meta['red_gain'], meta['blue_gain'])
    ref_lin = C.unprocess(image, rgb2cam, *gains)
    assert float((lin.cpu() - ref_lin).abs().max()) <= 2e-6
    t_mats = G.sample_transforms((432, 432), 14, 4, params)
    ref_rgb, ref_flow = L.single2lrburst(lin.cpu(), t_mats, 4, 24)          # from the device's linear image: byte-exact stage
    assert torch.equal(burst_rgb.cpu(), ref_rgb) and float((flow.cpu() - ref_flow).abs().max()) <= 1e-4
    shot, read = G.rgb2raw.random_noise_levels()
    assert (shot, read) == (meta['shot_noise_level'], meta['read_noise_level'])
    z = torch.FloatTensor(14, 4, 48, 48).normal_()
    d = (raw.cpu() - C.mosaic_add_noise(ref_rgb, shot, read, z)).abs()
    assert float(d.max()) <= 1.2e-7


def test_rgb2rawburst_batch_equals_per_burst_calls(dev):
    """rgb2rawburst_batch (one launch per kernel for the whole batch, parameters sampled for all bursts up front and uploaded
    once) under the same `random` / torch seeds == `rgb2rawburst` called burst after burst (the reference's schedule,
    data/synthetic_burst_generation.py:23-103): every output bit for bit with noise='host'; with noise='device' everything
    but the noise realisation is identical and the noisy burst has the same statistics."""
    import random
    from deep_rawburst_sr_b200.data import synthetic_burst_generation as G
    g = torch.Generator().manual_seed(2)
    images = torch.rand(3, 3, 240, 208, generator=g).to(dev)
    tp = {'max_translation': 24.0, 'max_rotation': 1.0, 'max_shear': 0.0, 'max_scale': 0.0, 'border_crop': 24}
    random.seed(77); torch.manual_seed(77)
    per = [G.rgb2rawburst(images[i], 5, 4, dict(tp), None) for i in range(3)]
    random.seed(77); torch.manual_seed(77)
    raw, lin, rgb, flow, metas = G.rgb2rawburst_batch(images, 5, 4, dict(tp), None, noise='host')
    assert tuple(raw.shape) == (3, 5, 4, 24, 20) and tuple(rgb.shape) == (3, 5, 3, 48, 40) and tuple(flow.shape) == (3, 5, 2, 48, 40)
    for i, (p_raw, p_lin, p_rgb, p_flow, p_meta) in enumerate(per):
        assert torch.equal(lin[i], p_lin) and torch.equal(rgb[i], p_rgb) and torch.equal(flow[i], p_flow), i
        assert torch.equal(raw[i], p_raw), i
        assert torch.equal(metas[i]['rgb2cam'], p_meta['rgb2cam']) and metas[i]['shot_noise_level'] == p_meta['shot_noise_level']
        assert metas[i]['red_gain'] == p_meta['red_gain'] and metas[i]['read_noise_level'] == p_meta['read_noise_level']
    # device-side noise: the host stream no longer contains the normal_ draws, so only burst 0 shares its parameters
    random.seed(77); torch.manual_seed(77)
    raw_d, lin_d, rgb_d, flow_d, _ = G.rgb2rawburst_batch(images, 5, 4, dict(tp), None, noise='device')
    assert torch.equal(lin_d[0], lin[0]) and torch.equal(rgb_d[0], rgb[0]) and torch.equal(flow_d, flow)
    assert float(raw_d.min()) >= 0.0 and float(raw_d.max()) <= 1.0 and abs(float(raw_d[0].mean()) - float(raw[0].mean())) < 5e-3
    with pytest.raises(NotImplementedError):
        G.rgb2rawburst_batch(images.cpu(), 5, 4, dict(tp), None)
