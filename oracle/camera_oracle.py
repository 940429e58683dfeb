"""CPU oracle of the inverse camera pipeline -- TEST INFRASTRUCTURE ONLY.

Restates reference data/camera_pipeline.py (`invert_smoothstep` :78-81, `gamma_expansion` :84-87, `apply_ccm` :96-107,
`safe_invert_gains` :121-136, `mosaic` :139-150, `add_noise` :178-183) and their order of use in
data/synthetic_burst_generation.py:59-99, with explicit per-channel arithmetic.  Pin: `oracle/make_golden_camera.py` runs the
reference's own functions and commits `tests/golden/camera_*.npz`; `tests/test_oracle.py` checks this file against them."""
import torch


def make_inputs(seed: int, h: int, w: int, n: int):
    g = torch.Generator().manual_seed(12000 + seed)
    image = torch.rand(3, h, w, generator=g)
    image[:, : h // 8] = (0.9 + 0.2 * torch.rand(3, h // 8, w, generator=g)).clamp(0, 1.05)      # saturated band: gain masking
    w4 = torch.rand(4, 1, 1, generator=g)
    base = torch.tensor([[[1.0234, -0.2969, -0.2266], [-0.5625, 1.6328, -0.0469], [-0.0703, 0.2188, 0.6406]],
                         [[0.4913, -0.0541, -0.0202], [-0.613, 1.3513, 0.2906], [-0.1564, 0.2151, 0.7183]],
                         [[0.838, -0.263, -0.0639], [-0.2887, 1.0725, 0.2496], [-0.0627, 0.1427, 0.5438]],
                         [[0.6596, -0.2079, -0.0562], [-0.4782, 1.3016, 0.1933], [-0.097, 0.1581, 0.5181]]])
    xyz2cam = (base * w4).sum(0) / w4.sum()
    rgb2xyz = torch.tensor([[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750], [0.0193339, 0.1191920, 0.9503041]])
    rgb2cam = xyz2cam @ rgb2xyz
    rgb2cam = rgb2cam / rgb2cam.sum(dim=-1, keepdim=True)
    gains = (1.0 / (0.8 + 0.05 * seed), 1.9 + 0.1 * seed, 1.5 + 0.1 * seed)                          # rgb, red, blue
    burst_rgb = torch.rand(n, 3, h // 2, w // 2, generator=g)
    noise_levels = (0.0001 * (3.0 ** seed), 0.00002 * (4.0 ** seed))                                   # shot, read
    return image, rgb2cam, gains, burst_rgb, noise_levels


def unprocess(image, rgb2cam, rgb_gain, red_gain, blue_gain, smoothstep=True, gamma=True):
    """synthetic_burst_generation.py:59-79"""
    x = image
    if smoothstep:
        x = x.clamp(0.0, 1.0)
        x = 0.5 - torch.sin(torch.asin(1.0 - 2.0 * x) / 3.0)
    if gamma:
        x = x.clamp(1e-8) ** 2.2
    cam = torch.stack([rgb2cam[c, 0] * x[0] + rgb2cam[c, 1] * x[1] + rgb2cam[c, 2] * x[2] for c in range(3)])
    gains = (torch.tensor([1.0 / red_gain, 1.0, 1.0 / blue_gain]) / rgb_gain).view(3, 1, 1)
    gray = cam.mean(dim=0, keepdim=True)
    mask = ((gray - 0.9).clamp(0.0) / (1.0 - 0.9)) ** 2.0
    safe = torch.max(mask + (1.0 - mask) * gains, gains)
    return (cam * safe).clamp(0.0, 1.0)


def mosaic(burst_rgb):
    """camera_pipeline.py:139-150, 'rggb'"""
    return torch.stack((burst_rgb[:, 0, 0::2, 0::2], burst_rgb[:, 1, 0::2, 1::2], burst_rgb[:, 1, 1::2, 0::2],
                        burst_rgb[:, 2, 1::2, 1::2]), dim=1)


def mosaic_add_noise(burst_rgb, shot, read, z):
    """synthetic_burst_generation.py:88-99 with the standard-normal draw `z` of add_noise made explicit"""
    raw = mosaic(burst_rgb)
    return (raw + z * (raw * shot + read).sqrt()).clamp(0.0, 1.0)
