"""CPU oracle for the DBSR burst forward pass -- TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 restatement (torch CPU tensors used as an array library: explicit gathers,
explicit coordinate arithmetic, `F.conv2d` for the dense contractions) of the algorithm the reference
executes on the hot path `DBSRNet.forward` (reference models/dbsr/dbsrnet.py:33-38).  It is NOT part of
the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it, and only as the checker / the timed CPU arm.  The product path
(`deep_rawburst_sr_b200`) never imports anything from `oracle/` and raises when the CUDA extension is
missing.

Parity pin: the reference has no tests or golden vectors for this path, so the
oracle is pinned against the reference's OWN modules imported in the build container
(`oracle/make_golden.py`, run against /root/reference with a cupy stub and a pure-torch cost volume as
BASELINE.json config 1 sanctions); the resulting vectors are committed under `tests/golden/` and
`tests/test_oracle.py` checks this file against them on every CPU run.

Every function cites the reference file:line it restates.  The un-vendored dependency is PyTorch itself
(install.sh:24 pins pytorch+cudatoolkit 10.2): the semantics restated by hand here are
`F.interpolate(bilinear, align_corners=False)`, `F.grid_sample(bilinear, zeros, align_corners=False)`,
`nn.PixelShuffle`, `F.softmax`, `torch.remainder`, `nn.ConvTranspose2d(4,2,1)`.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------
# Architecture constants (reference train_settings/dbsr/default_synthetic.py:73-82 and the factory
# defaults models/dbsr/dbsrnet.py:44-56)
# ----------------------------------------------------------------------------------------------------
ENC_INIT_DIM = 64
ENC_NUM_RES = 9
ENC_OUT_DIM = 512
DEC_INIT_DIM = 64
DEC_NUM_PRE_RES = 5
DEC_POST_DIM = 32
DEC_NUM_POST_RES = 4
UPSAMPLE = 8
OFFSET_FEAT_DIM = 64
PROJ_DIM = 64
NUM_OFFSET_RES = 1
NUM_WP_RES = 3
OFFSET_MODULO = 1.0
GAUSS_SD = 1.0
GAUSS_KSZ = 3

PWC_EXT_CH = [3, 16, 32, 64, 96, 128, 196]
PWC_LEVEL_NAMES = ['One', 'Two', 'Thr', 'Fou', 'Fiv', 'Six']
PWC_DEC_OUT = [128, 128, 96, 64, 32, 2]
# decoder input channels per level (pwcnet.py:116-117)
PWC_DEC_CURRENT = {6: 81, 5: 81 + 128 + 2 + 2, 4: 81 + 96 + 2 + 2, 3: 81 + 64 + 2 + 2, 2: 81 + 32 + 2 + 2}
PWC_BACKWARP_SCALE = {5: 0.625, 4: 1.25, 3: 2.5, 2: 5.0}  # pwcnet.py:121
PWC_REFINER = [(565, 128, 1), (128, 128, 2), (128, 128, 4), (128, 96, 8), (96, 64, 16), (64, 32, 1), (32, 2, 1)]


def state_dict_spec():
    """(key, shape) list in the order of the reference's state_dict."""
    spec = []
    pre = 'encoder.alignment_net.net.'
    for l, name in enumerate(PWC_LEVEL_NAMES):
        cin, cout = PWC_EXT_CH[l], PWC_EXT_CH[l + 1]
        for j, idx in enumerate((0, 2, 4)):
            ci = cin if j == 0 else cout
            spec.append((f'{pre}netExtractor.net{name}.{idx}.weight', (cout, ci, 3, 3)))
            spec.append((f'{pre}netExtractor.net{name}.{idx}.bias', (cout,)))
    for lvl in (2, 3, 4, 5, 6):  # module registration order netTwo..netSix (pwcnet.py:211-215)
        name = PWC_LEVEL_NAMES[lvl - 1]
        cur = PWC_DEC_CURRENT[lvl]
        if lvl < 6:
            prev = PWC_DEC_CURRENT[lvl + 1]
            spec.append((f'{pre}net{name}.netUpflow.weight', (2, 2, 4, 4)))
            spec.append((f'{pre}net{name}.netUpflow.bias', (2,)))
            spec.append((f'{pre}net{name}.netUpfeat.weight', (prev + 128 + 128 + 96 + 64 + 32, 2, 4, 4)))
            spec.append((f'{pre}net{name}.netUpfeat.bias', (2,)))
        cin = cur
        for j, cout in enumerate(PWC_DEC_OUT):
            sub = PWC_LEVEL_NAMES[j]
            spec.append((f'{pre}net{name}.net{sub}.0.weight', (cout, cin, 3, 3)))
            spec.append((f'{pre}net{name}.net{sub}.0.bias', (cout,)))
            cin += cout
    for j, (ci, co, _d) in enumerate(PWC_REFINER):
        spec.append((f'{pre}netRefiner.netMain.{2 * j}.weight', (co, ci, 3, 3)))
        spec.append((f'{pre}netRefiner.netMain.{2 * j}.bias', (co,)))
    spec.append(('encoder.init_layer.0.weight', (ENC_INIT_DIM, 4, 3, 3)))
    spec.append(('encoder.init_layer.0.bias', (ENC_INIT_DIM,)))
    for i in range(ENC_NUM_RES):
        for c in ('conv1', 'conv2'):
            spec.append((f'encoder.res_layers.{i}.{c}.0.weight', (ENC_INIT_DIM, ENC_INIT_DIM, 3, 3)))
            spec.append((f'encoder.res_layers.{i}.{c}.0.bias', (ENC_INIT_DIM,)))
    spec.append(('encoder.out_layer.0.weight', (ENC_OUT_DIM, ENC_INIT_DIM, 3, 3)))
    spec.append(('encoder.out_layer.0.bias', (ENC_OUT_DIM,)))
    spec.append(('merging.feat_project_layer.0.weight', (PROJ_DIM, ENC_OUT_DIM, 1, 1)))
    spec.append(('merging.feat_project_layer.0.bias', (PROJ_DIM,)))
    spec.append(('merging.offset_feat_extractor.0.0.weight', (OFFSET_FEAT_DIM, 2, 3, 3)))
    spec.append(('merging.offset_feat_extractor.0.0.bias', (OFFSET_FEAT_DIM,)))
    for i in range(NUM_OFFSET_RES):
        for c in ('conv1', 'conv2'):
            spec.append((f'merging.offset_feat_extractor.{i + 1}.{c}.0.weight', (OFFSET_FEAT_DIM, OFFSET_FEAT_DIM, 3, 3)))
            spec.append((f'merging.offset_feat_extractor.{i + 1}.{c}.0.bias', (OFFSET_FEAT_DIM,)))
    wp = 2 * PROJ_DIM
    spec.append(('merging.weight_predictor.0.0.weight', (wp, 2 * PROJ_DIM + OFFSET_FEAT_DIM, 3, 3)))
    spec.append(('merging.weight_predictor.0.0.bias', (wp,)))
    for i in range(NUM_WP_RES):
        for c in ('conv1', 'conv2'):
            spec.append((f'merging.weight_predictor.{i + 1}.{c}.0.weight', (wp, wp, 3, 3)))
            spec.append((f'merging.weight_predictor.{i + 1}.{c}.0.bias', (wp,)))
    spec.append((f'merging.weight_predictor.{NUM_WP_RES + 1}.0.weight', (ENC_OUT_DIM, wp, 3, 3)))
    spec.append((f'merging.weight_predictor.{NUM_WP_RES + 1}.0.bias', (ENC_OUT_DIM,)))
    spec.append(('decoder.init_layer.0.weight', (DEC_INIT_DIM, ENC_OUT_DIM, 3, 3)))
    spec.append(('decoder.init_layer.0.bias', (DEC_INIT_DIM,)))
    for i in range(DEC_NUM_PRE_RES):
        for c in ('conv1', 'conv2'):
            spec.append((f'decoder.pre_res_layers.{i}.{c}.0.weight', (DEC_INIT_DIM, DEC_INIT_DIM, 3, 3)))
            spec.append((f'decoder.pre_res_layers.{i}.{c}.0.bias', (DEC_INIT_DIM,)))
    spec.append(('decoder.upsample_layer.conv_layer.0.weight', (DEC_POST_DIM * UPSAMPLE ** 2, DEC_INIT_DIM, 1, 1)))
    for i in range(DEC_NUM_POST_RES):
        for c in ('conv1', 'conv2'):
            spec.append((f'decoder.post_res_layers.{i}.{c}.0.weight', (DEC_POST_DIM, DEC_POST_DIM, 3, 3)))
            spec.append((f'decoder.post_res_layers.{i}.{c}.0.bias', (DEC_POST_DIM,)))
    spec.append(('decoder.predictor.0.weight', (3, DEC_POST_DIM, 1, 1)))
    spec.append(('decoder.predictor.0.bias', (3,)))
    return spec


def make_state_dict(seed: int = 0, pwc_gain: float = 1.0, dbsr_gain: float = 1.0):
    """Deterministic random-init weights with the reference's key names and shapes.

    The weights contract is the state_dict, not the RNG stream: this recipe draws every
    tensor from its own seeded generator so it can be regenerated bit-identically anywhere (torch CPU
    RNG), instead of shipping 52 MB.  Scale follows PyTorch's default Conv2d init (uniform with bound
    1/sqrt(fan_in)); the upsampler weight has the ICNR structure (reference
    models/layers/initializations.py:21-38: all 64 sub-pixel kernels of an output channel identical).
    `pwc_gain` > 1 scales PWC-Net weights to produce multi-pixel flows (stress the warps / masks).
    """
    sd = OrderedDict()
    for i, (key, shape) in enumerate(state_dict_spec()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        is_pwc = 'alignment_net' in key
        gain = pwc_gain if is_pwc else dbsr_gain
        if key == 'decoder.upsample_layer.conv_layer.0.weight':
            sub = torch.randn(DEC_POST_DIM, DEC_INIT_DIM, 1, 1, generator=g) * math.sqrt(2.0 / DEC_INIT_DIM)
            sd[key] = sub.repeat_interleave(UPSAMPLE ** 2, dim=0).contiguous() * gain
            continue
        if key.endswith('.weight'):
            if 'netUpflow' in key or 'netUpfeat' in key:
                fan_in = shape[1] * shape[2] * shape[3]  # ConvTranspose2d: weight [Cin, Cout, k, k]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound * gain
        else:
            wshape = sd[key[:-4] + 'weight'].shape
            fan_in = wshape[1] * wshape[2] * wshape[3]
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def make_burst(seed: int, B: int, N: int, H: int, W: int):
    """Seeded synthetic burst in [0,1) (dataset value range: dataset/synthetic_burst_val_set.py:45)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(B, N, 4, H, W, generator=g)


# ----------------------------------------------------------------------------------------------------
# Elementary ops restated by hand
def make_realistic_burst(seed: int, B: int, N: int, H: int, W: int, max_shift: float = 24.0, max_rot_deg: float = 1.0):
    """Data-free restatement of the reference's synthetic burst generator (SURVEY.md 8(d) "realistic" inputs):
    a low-pass random RGB scene of (8H+48) x (8W+48) pixels; per frame a random translation of up to +-max_shift HR pixels and
    rotation of up to +-max_rot_deg (train_settings/dbsr/default_synthetic.py:37-41; frame 0 is the untransformed base
    frame, data/synthetic_burst_generation.py:150-160) resampled bilinearly; centre crop 8H x 8W; x4 box downsample to the
    2H x 2W RGB image; RGGB mosaic (data/camera_pipeline.py:139-150); shot / read noise with the levels of
    random_noise_levels / add_noise (camera_pipeline.py:165-182: log-uniform shot in [1e-4, 1.2e-2],
    log read = 2.18 log shot + 1.2 + N(0, 0.26)); clamp to [0, 1]; 14-bit quantisation.  Returns [B, N, 4, H, W]."""
    g = torch.Generator().manual_seed(77000 + seed)
    hr_h, hr_w = 8 * H + 48, 8 * W + 48
    coarse = torch.rand(B, 3, hr_h // 16 + 2, hr_w // 16 + 2, generator=g)
    scene = F.interpolate(coarse, size=(hr_h, hr_w), mode='bicubic', align_corners=False).clamp(0.0, 1.0)
    ys, xs = torch.meshgrid(torch.arange(hr_h, dtype=torch.float32), torch.arange(hr_w, dtype=torch.float32), indexing='ij')
    cy, cx = (hr_h - 1) / 2.0, (hr_w - 1) / 2.0
    frames = []
    for n in range(N):
        if n == 0:
            t = torch.zeros(B, 2)
            th = torch.zeros(B)
        else:
            t = (torch.rand(B, 2, generator=g) * 2 - 1) * max_shift
            th = (torch.rand(B, generator=g) * 2 - 1) * math.radians(max_rot_deg)
        c, s_ = torch.cos(th).view(B, 1, 1), torch.sin(th).view(B, 1, 1)
        u = c * (xs - cx) - s_ * (ys - cy) + cx + t[:, 0].view(B, 1, 1)
        v = s_ * (xs - cx) + c * (ys - cy) + cy + t[:, 1].view(B, 1, 1)
        moved = bilinear_sample_zeros(scene, u, v)
        crop = moved[:, :, 24:24 + 8 * H, 24:24 + 8 * W]
        rgb = F.avg_pool2d(crop, 4)                                    # [B, 3, 2H, 2W]
        raw = torch.stack([rgb[:, 0, 0::2, 0::2], rgb[:, 1, 0::2, 1::2], rgb[:, 1, 1::2, 0::2], rgb[:, 2, 1::2, 1::2]], 1)
        frames.append(raw)
    burst = torch.stack(frames, 1)                                     # [B, N, 4, H, W]
    log_shot = torch.rand(B, generator=g) * (math.log(1.2e-2) - math.log(1e-4)) + math.log(1e-4)
    log_read = 2.18 * log_shot + 1.2 + 0.26 * torch.randn(B, generator=g)
    shot, read = torch.exp(log_shot).view(B, 1, 1, 1, 1), torch.exp(log_read).view(B, 1, 1, 1, 1)
    burst = burst + torch.randn(burst.shape, generator=g) * torch.sqrt(burst * shot + read)
    burst = burst.clamp(0.0, 1.0)
    return (torch.round(burst * 2 ** 14) / 2 ** 14).contiguous()


# ----------------------------------------------------------------------------------------------------
def lrelu(x):
    return torch.where(x > 0, x, 0.1 * x)


def rggb_to_rgb(x):
    """models/dbsr/encoders.py:52: [R, mean(G1,G2), B]."""
    return torch.stack((x[..., 0, :, :], (x[..., 1, :, :] + x[..., 2, :, :]) / 2.0, x[..., 3, :, :]), dim=-3)


def _resize_coords(n_in: int, n_out: int):
    """F.interpolate(mode='bilinear', align_corners=False) source coordinates along one axis."""
    scale = n_in / n_out
    dst = torch.arange(n_out, dtype=torch.float32)
    src = (dst + 0.5) * scale - 0.5
    src = torch.clamp(src, min=0.0)
    i0 = src.floor().long()
    i0 = torch.clamp(i0, max=n_in - 1)
    i1 = torch.clamp(i0 + 1, max=n_in - 1)
    w1 = src - i0.float()
    return i0, i1, w1


def resize_bilinear(x, Ho: int, Wo: int):
    """F.interpolate(bilinear, align_corners=False), used at models/alignment/pwcnet.py:266-275."""
    H, W = x.shape[-2:]
    y0, y1, wy = _resize_coords(H, Ho)
    x0, x1, wx = _resize_coords(W, Wo)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    wy = wy.view(-1, 1)
    rows = top * (1 - wy) + bot * wy
    left = rows[..., x0]
    right = rows[..., x1]
    return left * (1 - wx) + right * wx


def bilinear_sample_zeros(img, u, v, with_mask: bool = False):
    """Bilinear sample of img[n,c,:,:] at pixel coords (u horizontal, v vertical) [n,h,w]; taps outside
    the image contribute 0 (grid_sample padding_mode='zeros', align_corners=False, after the
    normalise / un-normalise round trip of warp.py:28-44 and pwcnet.py:20-31).
    Returns [n,c,h,w] (and the sampled all-ones channel when with_mask)."""
    n, c, H, W = img.shape
    x0 = torch.floor(u)
    y0 = torch.floor(v)
    ax = u - x0
    ay = v - y0
    x0 = x0.long()
    y0 = y0.long()
    out = torch.zeros(n, c, *u.shape[-2:], dtype=img.dtype)
    mask = torch.zeros(n, *u.shape[-2:], dtype=img.dtype)
    flat = img.reshape(n, c, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx
            yi = y0 + dy
            w = (ax if dx else 1 - ax) * (ay if dy else 1 - ay)
            valid = ((xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)).to(img.dtype)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).view(n, 1, -1).expand(-1, c, -1)
            tap = torch.gather(flat, 2, idx).view(n, c, *u.shape[-2:])
            ww = (w * valid).unsqueeze(1)
            out = out + tap * ww
            mask = mask + w * valid
    if with_mask:
        return out, mask
    return out


def warp(feat, flow):
    """models/layers/warp.py:19-46: sample feat at pixel (x + flow_x, y + flow_y), zeros outside."""
    n, c, H, W = feat.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    # reference: grid = (x+0.5+flow); norm = 2*grid/W-1; grid_sample un-normalises ((norm+1)*W-1)/2
    gx = 2.0 * (xs + 0.5 + flow[:, 0]) / W - 1.0
    gy = 2.0 * (ys + 0.5 + flow[:, 1]) / H - 1.0
    u = ((gx + 1.0) * W - 1.0) / 2.0
    v = ((gy + 1.0) * H - 1.0) / 2.0
    return bilinear_sample_zeros(feat, u, v)


def backwarp(f2, flow):
    """models/alignment/pwcnet.py:16-38: sample at x + flow_x*W/(W-1) (the linspace grid is pixel
    centres, the flow is divided by (W-1)/2), zero the pixels whose sampled ones-channel <= 0.999."""
    n, c, H, W = f2.shape
    hor = torch.linspace(-1.0 + 1.0 / W, 1.0 - 1.0 / W, W).view(1, 1, W).expand(1, H, W)
    ver = torch.linspace(-1.0 + 1.0 / H, 1.0 - 1.0 / H, H).view(1, H, 1).expand(1, H, W)
    gx = hor + flow[:, 0] / ((W - 1.0) / 2.0)
    gy = ver + flow[:, 1] / ((H - 1.0) / 2.0)
    u = ((gx + 1.0) * W - 1.0) / 2.0
    v = ((gy + 1.0) * H - 1.0) / 2.0
    out, mask = bilinear_sample_zeros(f2, u, v, with_mask=True)
    mask = (mask > 0.999).to(f2.dtype)  # :34-36 (>0.999 -> 1 ; then <1 -> 0)
    return out * mask.unsqueeze(1)


def correlation81(f1, f2):
    """external/pwcnet/correlation/correlation.py:69-100: out[b, 9*(dy+4)+(dx+4), y, x] =
    mean_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx], zero padded by 4."""
    n, c, H, W = f1.shape
    f2p = F.pad(f2, (4, 4, 4, 4))
    out = torch.empty(n, 81, H, W, dtype=f1.dtype)
    for dy in range(9):
        for dx in range(9):
            out[:, dy * 9 + dx] = (f1 * f2p[:, :, dy:dy + H, dx:dx + W]).sum(1) / c
    return out


def correlation81_bruteforce(f1, f2):
    """Literal loop restatement of the CUDA kernel indexing (correlation.py:46-100) for tiny tensors."""
    n, c, H, W = f1.shape
    out = torch.zeros(n, 81, H, W, dtype=f1.dtype)
    for b in range(n):
        for y in range(H):
            for x in range(W):
                for top in range(81):
                    s2o = top % 9 - 4
                    s2p = top // 9 - 4
                    y2, x2 = y + s2p, x + s2o
                    if 0 <= y2 < H and 0 <= x2 < W:
                        out[b, top, y, x] = (f1[b, :, y, x] * f2[b, :, y2, x2]).sum() / c
    return out


def deconv4x4s2(x, w, b):
    """nn.ConvTranspose2d(k=4, s=2, p=1) (pwcnet.py:119-120): out[oc, 2*iy-1+ky, 2*ix-1+kx] +=
    x[ic, iy, ix] * w[ic, oc, ky, kx]."""
    n, cin, H, W = x.shape
    cout = w.shape[1]
    out = torch.zeros(n, cout, 2 * H + 2, 2 * W + 2, dtype=x.dtype)  # padded canvas, crop 1 at the end
    for ky in range(4):
        for kx in range(4):
            contrib = torch.einsum('nihw,io->nohw', x, w[:, :, ky, kx])
            out[:, :, ky:ky + 2 * H:2, kx:kx + 2 * W:2] += contrib
    return out[:, :, 1:1 + 2 * H, 1:1 + 2 * W] + b.view(1, -1, 1, 1)


def pixel_shuffle(x, r: int):
    """nn.PixelShuffle: out[c, r*h+i, r*w+j] = in[c*r*r + i*r + j, h, w]."""
    n, c, H, W = x.shape
    co = c // (r * r)
    x = x.view(n, co, r, r, H, W).permute(0, 1, 4, 2, 5, 3)
    return x.reshape(n, co, H * r, W * r)


def gauss_kernel3(sd: float = GAUSS_SD):
    """models/layers/filtering.py:20-40 + upsampling.py:24-29: normalised ksz=3 Gaussian."""
    k = torch.arange(-1.0, 2.0)
    g = torch.exp(-1.0 / (2 * sd ** 2) * k ** 2) / (math.sqrt(2 * math.pi) * sd)
    K = g.view(1, -1) * g.view(-1, 1)
    return K / K.sum()


def conv(x, sd, key, stride=1, padding=1, dilation=1):
    b = sd.get(key + '.bias')
    return F.conv2d(x, sd[key + '.weight'], b, stride=stride, padding=padding, dilation=dilation)


def resblock(x, sd, key):
    """models/layers/blocks.py:63-96 (no BN, relu): relu(x + conv2(relu(conv1(x))))."""
    out = torch.relu(conv(x, sd, key + '.conv1.0'))
    out = conv(out, sd, key + '.conv2.0')
    return torch.relu(out + x)


# ----------------------------------------------------------------------------------------------------
# PWC-Net (models/alignment/pwcnet.py)
# ----------------------------------------------------------------------------------------------------
def pwc_extractor(x, sd, pre):
    """pwcnet.py:45-111."""
    feats = []
    for name in PWC_LEVEL_NAMES:
        k = f'{pre}netExtractor.net{name}'
        x = lrelu(conv(x, sd, k + '.0', stride=2))
        x = lrelu(conv(x, sd, k + '.2'))
        x = lrelu(conv(x, sd, k + '.4'))
        feats.append(x)
    return feats


def pwc_decoder(lvl, f1, f2, prev, sd, pre, trace=None):
    """pwcnet.py:153-184."""
    name = PWC_LEVEL_NAMES[lvl - 1]
    k = f'{pre}net{name}'
    if prev is None:
        vol = lrelu(correlation81(f1, f2))
        feat = vol
    else:
        upflow = deconv4x4s2(prev['flow'], sd[k + '.netUpflow.weight'], sd[k + '.netUpflow.bias'])
        upfeat = deconv4x4s2(prev['feat'], sd[k + '.netUpfeat.weight'], sd[k + '.netUpfeat.bias'])
        f2w = backwarp(f2, upflow * PWC_BACKWARP_SCALE[lvl])
        vol = lrelu(correlation81(f1, f2w))
        feat = torch.cat([vol, f1, upflow, upfeat], 1)
        if trace is not None:
            trace[f'pwc_upflow{lvl}'] = upflow
            trace[f'pwc_upfeat{lvl}'] = upfeat
    if trace is not None:
        trace[f'pwc_vol{lvl}'] = vol
    for sub in PWC_LEVEL_NAMES[:5]:
        feat = torch.cat([lrelu(conv(feat, sd, f'{k}.net{sub}.0')), feat], 1)  # new channels in front
    flow = conv(feat, sd, f'{k}.netSix.0')
    if trace is not None:
        trace[f'pwc_flow{lvl}'] = flow
    return {'flow': flow, 'feat': feat}


def pwc_refiner(x, sd, pre):
    """pwcnet.py:186-207."""
    n = len(PWC_REFINER)
    for j, (_ci, _co, d) in enumerate(PWC_REFINER):
        x = conv(x, sd, f'{pre}netRefiner.netMain.{2 * j}', padding=d, dilation=d)
        if j < n - 1:
            x = lrelu(x)
    return x


def pwc_network(first, second, sd, pre, trace=None):
    """pwcnet.py:221-231."""
    f1 = pwc_extractor(first, sd, pre)
    f2 = pwc_extractor(second, sd, pre)
    est = None
    for lvl in (6, 5, 4, 3, 2):
        est = pwc_decoder(lvl, f1[lvl - 1], f2[lvl - 1], est, sd, pre, trace)
    return est['flow'] + pwc_refiner(est['feat'], sd, pre)


def pwcnet_forward(source, target, sd, pre='encoder.alignment_net.net.', trace=None):
    """PWCNet.forward, pwcnet.py:248-281. flow maps target -> source pixels, (x, y) order."""
    H, W = source.shape[-2:]
    source = source.reshape(-1, 3, H, W)
    target = target.reshape(-1, 3, H, W)
    Hp = int(math.floor(math.ceil(H / 64.0) * 64.0))
    Wp = int(math.floor(math.ceil(W / 64.0) * 64.0))
    src_re = resize_bilinear(source, Hp, Wp)
    tgt_re = resize_bilinear(target, Hp, Wp)
    flow = pwc_network(tgt_re, src_re, sd, pre, trace)
    if trace is not None:
        trace['pwc_flow_quarter'] = flow
    flow = 20.0 * resize_bilinear(flow, H, W)
    sx = float(W) / float(Wp)
    sy = float(H) / float(Hp)
    return torch.stack((flow[:, 0] * sx, flow[:, 1] * sy), dim=1)


# ----------------------------------------------------------------------------------------------------
# DBSR (models/dbsr/*.py)
# ----------------------------------------------------------------------------------------------------
def encoder_forward(x, sd, trace=None):
    """ResEncoderWarpAlignnet.forward, models/dbsr/encoders.py:48-86."""
    assert x.dim() == 5
    B, N, _, H, W = x.shape
    x_rgb = rggb_to_rgb(x)
    x_ref = x_rgb[:, :1].repeat(1, N - 1, 1, 1, 1)
    x_oth = x_rgb[:, 1:]
    offsets = pwcnet_forward(x_oth.reshape(-1, 3, H, W), x_ref.reshape(-1, 3, H, W), sd, trace=trace)
    out = torch.relu(conv(x.reshape(-1, 4, H, W), sd, 'encoder.init_layer.0'))
    for i in range(ENC_NUM_RES):
        out = resblock(out, sd, f'encoder.res_layers.{i}')
    feat = torch.relu(conv(out, sd, 'encoder.out_layer.0'))
    feat = feat.view(B, N, ENC_OUT_DIM, H, W)
    ref_feat = feat[:, :1]
    oth_feat = warp(feat[:, 1:].reshape(-1, ENC_OUT_DIM, H, W), offsets).view(B, N - 1, ENC_OUT_DIM, H, W)
    offsets = offsets.view(B, N - 1, 2, H, W)
    if trace is not None:
        trace['enc_feat'] = feat
    return {'ref_feat': ref_feat, 'oth_feat': oth_feat, 'offsets': offsets}


def merging_forward(enc, sd, trace=None):
    """WeightedSum.forward, models/dbsr/merging.py:61-127 (softmax, use_base_frame, offset_modulo=1)."""
    ref_feat, oth_feat, offsets = enc['ref_feat'], enc['oth_feat'], enc['offsets']
    B, _, C, H, W = ref_feat.shape
    all_feat = torch.cat((ref_feat[:, :1], oth_feat), dim=1)
    N = all_feat.shape[1]
    proj = torch.relu(conv(all_feat.reshape(-1, C, H, W), sd, 'merging.feat_project_layer.0', padding=0))
    proj = proj.view(B, N, PROJ_DIM, H, W)
    base = proj[:, :1]
    diff = (proj - base).reshape(-1, PROJ_DIM, H, W)
    base = base.expand(-1, N, -1, -1, -1).reshape(-1, PROJ_DIM, H, W)
    offs = torch.cat((torch.zeros(B, 1, 2, H, W), offsets), dim=1).reshape(-1, 2, H, W)
    offs = torch.remainder(offs, OFFSET_MODULO)  # floor-mod, merging.py:104-105
    of = torch.relu(conv(offs, sd, 'merging.offset_feat_extractor.0.0'))
    for i in range(NUM_OFFSET_RES):
        of = resblock(of, sd, f'merging.offset_feat_extractor.{i + 1}')
    wp = torch.cat([base, diff, of], dim=1)
    wp = torch.relu(conv(wp, sd, 'merging.weight_predictor.0.0'))
    for i in range(NUM_WP_RES):
        wp = resblock(wp, sd, f'merging.weight_predictor.{i + 1}')
    logits = conv(wp, sd, f'merging.weight_predictor.{NUM_WP_RES + 1}.0').view(B, N, C, H, W)
    m = logits.max(dim=1, keepdim=True).values
    e = torch.exp(logits - m)
    w = e / e.sum(dim=1, keepdim=True)
    fused = (all_feat * w).sum(dim=1)
    if trace is not None:
        trace['logits'] = logits
        trace['all_feat'] = all_feat
    return {'fused_enc': fused, 'fusion_weights': w}


def decoder_forward(mer, sd, trace=None):
    """ResPixShuffleConv.forward, models/dbsr/decoders.py:54-62 + PixShuffleUpsampler upsampling.py:51-66."""
    x = mer['fused_enc']
    out = torch.relu(conv(x, sd, 'decoder.init_layer.0'))
    for i in range(DEC_NUM_PRE_RES):
        out = resblock(out, sd, f'decoder.pre_res_layers.{i}')
    up = torch.relu(F.conv2d(out, sd['decoder.upsample_layer.conv_layer.0.weight']))
    up = pixel_shuffle(up, UPSAMPLE)
    n, c, Hh, Wh = up.shape
    up = F.conv2d(up.reshape(-1, 1, Hh, Wh), gauss_kernel3().view(1, 1, 3, 3), padding=1).view(n, c, Hh, Wh)
    if trace is not None:
        trace['dec_up'] = up
    out = up
    for i in range(DEC_NUM_POST_RES):
        out = resblock(out, sd, f'decoder.post_res_layers.{i}')
    pred = torch.relu(conv(out, sd, 'decoder.predictor.0', padding=0))  # conv_block default act = relu
    return {'pred': pred}


@torch.no_grad()
def dbsr_forward(im, sd, trace=None):
    """DBSRNet.forward, models/dbsr/dbsrnet.py:33-38."""
    enc = encoder_forward(im, sd, trace)
    mer = merging_forward(enc, sd, trace)
    dec = decoder_forward(mer, sd, trace)
    return dec['pred'], {'offsets': enc['offsets'], 'fusion_weights': mer['fusion_weights']}


def psnr(pred, gt, boundary_ignore: int = 40, max_value: float = 1.0):
    """models/loss/image_quality_v2.py:47-51,75-101: per image 20log10(max) - 10log10(mse), then mean."""
    if boundary_ignore:
        pred = pred[..., boundary_ignore:-boundary_ignore, boundary_ignore:-boundary_ignore]
        gt = gt[..., boundary_ignore:-boundary_ignore, boundary_ignore:-boundary_ignore]
    vals = []
    for p, g in zip(pred, gt):
        mse = ((p.double() - g.double()) ** 2).mean()
        vals.append(20 * math.log10(max_value) - 10.0 * math.log10(float(mse)))
    return sum(vals) / len(vals)


# ----------------------------------------------------------------------------------------------------
# Library-op restatement: the SAME op sequence the reference modules execute (ATen grid_sample /
# interpolate / conv_transpose2d / pixel_shuffle / softmax), used as the timed CPU arm of bench.py
# (`cpu_baseline`, `--impl reference`) so that the baseline runs at the reference's own speed rather than
# at the speed of the explicit-gather restatement above.  tests/test_oracle.py pins it to the explicit one.
# ----------------------------------------------------------------------------------------------------
def _ref_backwarp(f2, flow):
    """pwcnet.py:16-38 verbatim op sequence."""
    n, c, H, W = f2.shape
    hor = torch.linspace(-1.0 + 1.0 / W, 1.0 - 1.0 / W, W).view(1, 1, 1, -1).expand(-1, -1, H, -1)
    ver = torch.linspace(-1.0 + 1.0 / H, 1.0 - 1.0 / H, H).view(1, 1, -1, 1).expand(-1, -1, -1, W)
    grid = torch.cat([hor, ver], 1)
    flow = torch.cat([flow[:, 0:1] / ((W - 1.0) / 2.0), flow[:, 1:2] / ((H - 1.0) / 2.0)], 1)
    inp = torch.cat([f2, f2.new_ones(n, 1, H, W)], 1)
    out = F.grid_sample(inp, (grid + flow).permute(0, 2, 3, 1), mode='bilinear', padding_mode='zeros', align_corners=False)
    mask = out[:, -1:]
    mask = (mask > 0.999).to(out.dtype)
    return out[:, :-1] * mask


def _ref_corr(f1, f2):
    n, c, H, W = f1.shape
    f2p = F.pad(f2, (4, 4, 4, 4))
    return torch.cat([(f1 * f2p[:, :, dy:dy + H, dx:dx + W]).mean(1, keepdim=True) for dy in range(9) for dx in range(9)], 1)


def _ref_warp(feat, flow):
    """models/layers/warp.py:19-46 verbatim op sequence."""
    B, C, H, W = feat.shape
    rowv, colv = torch.meshgrid([torch.arange(0.5, H + 0.5), torch.arange(0.5, W + 0.5)], indexing='ij')
    grid = torch.stack((colv, rowv), dim=0).unsqueeze(0).float() + flow
    gn = torch.stack((2.0 * grid[:, 0] / W - 1.0, 2.0 * grid[:, 1] / H - 1.0), dim=1).permute(0, 2, 3, 1)
    return F.grid_sample(feat, gn, mode='bilinear', padding_mode='zeros', align_corners=False)


@torch.no_grad()
def dbsr_forward_fast(im, sd):
    """DBSRNet.forward with the reference's own library-op sequence (CPU timing arm)."""
    B, N, _, H, W = im.shape
    pre = 'encoder.alignment_net.net.'
    x_rgb = torch.stack((im[:, :, 0], im[:, :, 1:3].mean(dim=2), im[:, :, 3]), dim=2)
    x_ref = x_rgb[:, :1].repeat(1, N - 1, 1, 1, 1).contiguous().view(-1, 3, H, W)
    x_oth = x_rgb[:, 1:].contiguous().view(-1, 3, H, W)
    Hp = int(math.floor(math.ceil(H / 64.0) * 64.0))
    Wp = int(math.floor(math.ceil(W / 64.0) * 64.0))
    src_re = F.interpolate(x_oth, size=(Hp, Wp), mode='bilinear', align_corners=False)
    tgt_re = F.interpolate(x_ref, size=(Hp, Wp), mode='bilinear', align_corners=False)
    f1 = pwc_extractor(tgt_re, sd, pre)
    f2 = pwc_extractor(src_re, sd, pre)
    est = None
    for lvl in (6, 5, 4, 3, 2):
        name = PWC_LEVEL_NAMES[lvl - 1]
        k = f'{pre}net{name}'
        a, b = f1[lvl - 1], f2[lvl - 1]
        if est is None:
            feat = F.leaky_relu(_ref_corr(a, b), 0.1)
        else:
            upflow = F.conv_transpose2d(est['flow'], sd[k + '.netUpflow.weight'], sd[k + '.netUpflow.bias'], stride=2, padding=1)
            upfeat = F.conv_transpose2d(est['feat'], sd[k + '.netUpfeat.weight'], sd[k + '.netUpfeat.bias'], stride=2, padding=1)
            vol = F.leaky_relu(_ref_corr(a, _ref_backwarp(b, upflow * PWC_BACKWARP_SCALE[lvl])), 0.1)
            feat = torch.cat([vol, a, upflow, upfeat], 1)
        for sub in PWC_LEVEL_NAMES[:5]:
            feat = torch.cat([F.leaky_relu(conv(feat, sd, f'{k}.net{sub}.0'), 0.1), feat], 1)
        est = {'flow': conv(feat, sd, f'{k}.netSix.0'), 'feat': feat}
    x = est['feat']
    for j, (_ci, _co, d) in enumerate(PWC_REFINER):
        x = conv(x, sd, f'{pre}netRefiner.netMain.{2 * j}', padding=d, dilation=d)
        if j < len(PWC_REFINER) - 1:
            x = F.leaky_relu(x, 0.1)
    flow = 20.0 * F.interpolate(est['flow'] + x, size=(H, W), mode='bilinear', align_corners=False)
    offsets = torch.stack((flow[:, 0] * (float(W) / Wp), flow[:, 1] * (float(H) / Hp)), dim=1)
    out = torch.relu(conv(im.view(-1, 4, H, W), sd, 'encoder.init_layer.0'))
    for i in range(ENC_NUM_RES):
        out = resblock(out, sd, f'encoder.res_layers.{i}')
    feat = torch.relu(conv(out, sd, 'encoder.out_layer.0')).view(B, N, ENC_OUT_DIM, H, W)
    oth = _ref_warp(feat[:, 1:].contiguous().view(-1, ENC_OUT_DIM, H, W), offsets).view(B, N - 1, ENC_OUT_DIM, H, W)
    offsets = offsets.view(B, N - 1, 2, H, W)
    all_feat = torch.cat((feat[:, :1].contiguous(), oth), dim=1)
    proj = torch.relu(conv(all_feat.view(-1, ENC_OUT_DIM, H, W), sd, 'merging.feat_project_layer.0', padding=0)).view(B, N, -1, H, W)
    base = proj[:, :1].contiguous()
    diff = (proj - base).view(-1, PROJ_DIM, H, W)
    base = base.expand(-1, N, -1, -1, -1).contiguous().view(-1, PROJ_DIM, H, W)
    offs = torch.cat((torch.zeros(B, 1, 2, H, W), offsets), dim=1).view(-1, 2, H, W) % OFFSET_MODULO
    of = torch.relu(conv(offs, sd, 'merging.offset_feat_extractor.0.0'))
    for i in range(NUM_OFFSET_RES):
        of = resblock(of, sd, f'merging.offset_feat_extractor.{i + 1}')
    wp = torch.relu(conv(torch.cat([base, diff, of], dim=1), sd, 'merging.weight_predictor.0.0'))
    for i in range(NUM_WP_RES):
        wp = resblock(wp, sd, f'merging.weight_predictor.{i + 1}')
    logits = conv(wp, sd, f'merging.weight_predictor.{NUM_WP_RES + 1}.0').view(B, N, ENC_OUT_DIM, H, W)
    w = F.softmax(logits, dim=1)
    fused = (all_feat * w).sum(dim=1)
    out = torch.relu(conv(fused, sd, 'decoder.init_layer.0'))
    for i in range(DEC_NUM_PRE_RES):
        out = resblock(out, sd, f'decoder.pre_res_layers.{i}')
    up = F.pixel_shuffle(torch.relu(F.conv2d(out, sd['decoder.upsample_layer.conv_layer.0.weight'])), UPSAMPLE)
    shp = up.shape
    up = F.conv2d(up.view(-1, 1, *shp[-2:]), gauss_kernel3().view(1, 1, 3, 3), padding=1).view(shp)
    for i in range(DEC_NUM_POST_RES):
        up = resblock(up, sd, f'decoder.post_res_layers.{i}')
    pred = torch.relu(conv(up, sd, 'decoder.predictor.0', padding=0))
    return pred, {'offsets': offsets, 'fusion_weights': w}
