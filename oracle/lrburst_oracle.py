"""CPU oracle of `single2lrburst` for given frame transforms -- TEST INFRASTRUCTURE ONLY.

Restates reference data/synthetic_burst_generation.py:131-246 (`get_tmat` :106-128, uint8 quantisation :153-157,
`cv2.warpAffine(..., INTER_LINEAR, BORDER_CONSTANT)` :211-212, `border_crop` :220-224, `cv2.resize(fx = fy = 1 / factor,
INTER_LINEAR)` :227-232, `/ 255` :236-237, the sampling-position maps and flow vectors :214-218, 229-246) in integer / numpy
arithmetic.  The two OpenCV calls are byte work; their algorithm lives in the un-vendored dependency OpenCV (opencv-python,
unpinned by the reference's install.sh; 4.13.0 in this image) and is restated from its published fixed-point scheme:
  warpAffine, 8-bit, bilinear: the inverse map is evaluated in 1/1024 px (AB_BITS = 10) with per-column / per-row terms
    rounded separately (cvRound, half to even) plus a rounding offset of 16, reduced to 1/32 px (INTER_BITS = 5); the four tap
    weights are (32 - fx)(32 - fy) * 32 ... (sum 2^15, INTER_REMAP_COEF_BITS = 15), result (sum + 2^14) >> 15; taps outside the
    image contribute the border value 0.
  resize, 8-bit, bilinear, integer down-scale f: sample position (dx + 0.5) f - 0.5, i.e. for even f the two pixels
    f dx + f/2 - 1, f dx + f/2 with weights 1024 / 2048 each and (sum of the 2x2 block + 2) >> 2 after the vertical pass
    (INTER_RESIZE_COEF_BITS = 11); for odd f the single pixel f dx + (f - 1) / 2.
Pin: `oracle/make_golden_lrburst.py` runs the reference's own `single2lrburst` (OpenCV itself, Python `random` seeded) and
commits `tests/golden/lrburst_*.npz` with the sampled transforms; `tests/test_oracle.py` requires BIT-EXACT bursts."""
import math

import numpy as np
import torch


def rotation_matrix_2d(center, angle_deg):
    """cv2.getRotationMatrix2D(center, angle, 1.0): [[a, b, (1-a) cx - b cy], [-b, a, b cx + (1-a) cy]], double precision"""
    a = math.cos(angle_deg * math.pi / 180.0)
    b = math.sin(angle_deg * math.pi / 180.0)
    cx, cy = center
    return np.array([[a, b, (1 - a) * cx - b * cy], [-b, a, b * cx + (1 - a) * cy]], dtype=np.float64)


def get_tmat(image_shape, translation, theta, shear_values, scale_factors):
    """synthetic_burst_generation.py:106-128"""
    im_h, im_w = image_shape
    t_mat = np.identity(3)
    t_mat[0, 2], t_mat[1, 2] = translation
    t_rot = np.concatenate((rotation_matrix_2d((im_w * 0.5, im_h * 0.5), theta), np.array([[0.0, 0.0, 1.0]])))
    t_shear = np.array([[1.0, shear_values[0], -shear_values[0] * 0.5 * im_w],
                        [shear_values[1], 1.0, -shear_values[1] * 0.5 * im_h], [0.0, 0.0, 1.0]])
    t_scale = np.array([[scale_factors[0], 0.0, 0.0], [0.0, scale_factors[1], 0.0], [0.0, 0.0, 1.0]])
    return (t_scale @ t_rot @ t_shear @ t_mat)[:2, :]


def invert_affine(t_mat):
    """the double-precision inversion cv::warpAffine applies to a forward matrix (no WARP_INVERSE_MAP)"""
    M = np.array(t_mat, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def warp_affine_u8(img, t_mat):
    """cv2.warpAffine(img [H, W, C] uint8, t_mat, (W, H), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT)"""
    H, W, _ = img.shape
    M = invert_affine(t_mat)
    rint = lambda a: np.rint(a).astype(np.int64)          # cvRound: half to even
    xs, ys = np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64)
    adelta, bdelta = rint(M[0] * xs * 1024), rint(M[3] * xs * 1024)
    X0, Y0 = rint((M[1] * ys + M[2]) * 1024) + 16, rint((M[4] * ys + M[5]) * 1024) + 16
    X, Y = (X0[:, None] + adelta[None, :]) >> 5, (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    fx, fy = X & 31, Y & 31

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        return img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64) * ok[..., None]

    acc = (tap(sy, sx) * ((32 - fx) * (32 - fy) * 32)[..., None] + tap(sy, sx + 1) * (fx * (32 - fy) * 32)[..., None] +
           tap(sy + 1, sx) * ((32 - fx) * fy * 32)[..., None] + tap(sy + 1, sx + 1) * (fx * fy * 32)[..., None])
    return np.clip((acc + 16384) >> 15, 0, 255).astype(np.uint8)


def downsample_u8(img, f):
    """cv2.resize(img, None, fx=1/f, fy=1/f, interpolation=INTER_LINEAR) for uint8 and sizes divisible by the integer f"""
    H, W, _ = img.shape
    assert H % f == 0 and W % f == 0
    v = img.astype(np.int64)
    if f % 2 == 1:
        return img[(f - 1) // 2::f, (f - 1) // 2::f].copy()
    o = f // 2 - 1
    s = v[o::f, o::f] + v[o::f, o + 1::f] + v[o + 1::f, o::f] + v[o + 1::f, o + 1::f]
    return ((s + 2) >> 2).astype(np.uint8)


def downsample_f32(a, f):
    """the same resize on a float32 map (the sampling positions): 0.5 a + 0.5 b horizontally, then vertically"""
    assert f % 2 == 0
    o = f // 2 - 1
    h = a[:, o::f] * np.float32(0.5) + a[:, o + 1::f] * np.float32(0.5)
    return h[o::f] * np.float32(0.5) + h[o + 1::f] * np.float32(0.5)


def single2lrburst(image, t_mats, downsample_factor=4, border_crop=None):
    """synthetic_burst_generation.py:131-246 for given 2x3 forward matrices (frame 0 first) -> (burst [n, 3, h, w] fp32,
    flow_vectors [n, 2, h, w] fp32)"""
    normalize = bool(image.max() < 2.0)
    if normalize:
        image = image * 255.0
    img = image.permute(1, 2, 0).numpy().astype(np.uint8)            # float -> uint8 truncates (:157)
    H, W, _ = img.shape
    rvs, cvs = torch.meshgrid([torch.arange(0, H), torch.arange(0, W)], indexing='ij')
    grid = torch.stack((cvs, rvs, torch.ones_like(cvs)), dim=-1).float()
    burst, pos = [], []
    for t_mat in t_mats:
        image_t = warp_affine_u8(img, t_mat)
        t3 = torch.cat((torch.from_numpy(np.asarray(t_mat)).float(), torch.tensor([0.0, 0.0, 1.0]).view(1, 3)), dim=0)
        t_inv = t3.inverse()[:2, :].contiguous()
        sample_pos_inv = torch.mm(grid.view(-1, 3), t_inv.t().float()).view(H, W, 2)
        if border_crop is not None:
            image_t = image_t[border_crop:-border_crop, border_crop:-border_crop, :]
            sample_pos_inv = sample_pos_inv[border_crop:-border_crop, border_crop:-border_crop, :]
        image_t = downsample_u8(image_t, downsample_factor)
        sp = downsample_f32(sample_pos_inv.numpy(), downsample_factor)
        frame = torch.from_numpy(image_t).float().permute(2, 0, 1)
        burst.append(frame / 255.0 if normalize else frame)
        pos.append(torch.from_numpy(sp).permute(2, 0, 1) / downsample_factor)
    pos = torch.stack(pos)
    return torch.stack(burst), pos - pos[:1]
