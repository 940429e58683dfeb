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
