"""Thin torch-tensor wrappers over the C ABI (include/dbsr_b200.h).  PyTorch is used for device memory and
streams only; every op below launches a hand-written sm_100a kernel from libdbsr_b200.so on the current
torch stream.  Nothing here has a CPU / eager-PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, DBSR_BF16, DBSR_F32, ConvDesc, NhwcView, ResBlockDesc  # noqa: F401

_checked_devices = set()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _first_cuda_tensor(args):
    for a in args:
        if torch.is_tensor(a):
            if a.is_cuda:
                return a
        elif isinstance(a, dict):
            t = _first_cuda_tensor(a.values())
            if t is not None:
                return t
    return None


def tensor_device_guard(fn):
    """decorator of the module `forward`s: run with the device of the first CUDA tensor argument current.  Every launch of
    this package goes to `torch.cuda.current_stream()` of the CURRENT device and the library caches per current device, so a
    module living on cuda:1 must not run while cuda:0 is current (multi-GPU in one process, nn.DataParallel style)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(args)
        if t is None:
            t = _first_cuda_tensor(kwargs.values())
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


def require_device(t: torch.Tensor) -> None:
    """Fail loudly on anything but a CUDA tensor on an sm_100 device (reference behaviour for the cost volume on
    CPU input is NotImplementedError, external/pwcnet/correlation/correlation.py:324-325)."""
    if not t.is_cuda:
        raise NotImplementedError('deep_rawburst_sr_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path')
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        raise RuntimeError(f'tensor on cuda:{idx} but cuda:{torch.cuda.current_device()} is the current device: launches of this '
                           'package go to the current device\'s stream (run under `torch.cuda.device(tensor.device)`; the module '
                           'forwards do that themselves)')
    if idx not in _checked_devices:
        _lib.check(_lib.load_library().dbsr_device_check(idx), 'dbsr_device_check')
        _checked_devices.add(idx)


class Act:
    """Channels-last activation view: a channel slice [c_off, c_off + c) of a [n, h, w, c_pitch] torch buffer."""

    __slots__ = ('buf', 'n', 'h', 'w', 'c', 'c_off')

    def __init__(self, buf: torch.Tensor, c_off: int = 0, c: Optional[int] = None):
        assert buf.dim() == 4 and buf.is_contiguous(), 'Act needs a contiguous [n,h,w,c] buffer'
        assert buf.dtype in (torch.float32, torch.bfloat16)
        self.buf = buf
        self.n, self.h, self.w = buf.shape[0], buf.shape[1], buf.shape[2]
        self.c_off = c_off
        self.c = buf.shape[3] - c_off if c is None else c
        assert 0 <= self.c_off and self.c_off + self.c <= buf.shape[3]

    @staticmethod
    def empty(n, h, w, c, dtype, device, zero: bool = False):
        f = torch.zeros if zero else torch.empty
        return Act(f((n, h, w, c), dtype=dtype, device=device))

    def slice(self, c_off: int, c: int) -> 'Act':
        return Act(self.buf, self.c_off + c_off, c)

    def images(self, start: int, count: int) -> 'Act':
        return Act(self.buf[start:start + count], self.c_off, self.c)

    @property
    def dtype(self):
        return self.buf.dtype

    def view(self) -> NhwcView:
        return NhwcView(self.buf.data_ptr(), self.n, self.h, self.w, self.c, self.c_off, self.buf.shape[3],
                        DBSR_F32 if self.buf.dtype == torch.float32 else DBSR_BF16)

    def to_nchw(self) -> torch.Tensor:
        out = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.buf.device)
        v = self.view()
        _lib.check(_lib.load_library().dbsr_nhwc_to_nchw(ctypes.byref(v), out.data_ptr(), _stream()), 'nhwc_to_nchw')
        return out

    def from_nchw(self, src: torch.Tensor) -> 'Act':
        assert src.dtype == torch.float32 and src.is_contiguous() and tuple(src.shape) == (self.n, self.c, self.h, self.w)
        v = self.view()
        _lib.check(_lib.load_library().dbsr_nchw_to_nhwc(src.data_ptr(), ctypes.byref(v), _stream()), 'nchw_to_nhwc')
        return self


_NULL_VIEW = NhwcView(None, 0, 0, 0, 0, 0, 0, 0)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def conv2d(x: Act, w: torch.Tensor, bias: Optional[torch.Tensor], y: Act, ksize: int, stride: int = 1, dilation: int = 1,
           act: int = ACT_NONE, residual: Optional[Act] = None, shuffle_r: int = 0, tensor_core: bool = False,
           grid_limit: int = 0, residual_group: int = 0, static_weights: bool = False) -> Act:
    """grid_limit (tensor-core path): cap of the persistent grid of THIS launch, 0 = one CTA per SM.
    static_weights (tensor-core path): w / bias were uploaded earlier and no kernel queued before this call writes them, so
    the kernel may fetch them while the preceding kernel still runs (DBSR_CONV_STATIC_WEIGHTS).  Default off: safe for callers
    that pack weights on the device right before the call.
    residual_group g > 1 (tensor-core path): output image i adds residual image i // g (a per-burst map broadcast over frames)"""
    d = ConvDesc(x.view(), y.view(), residual.view() if residual is not None else _NULL_VIEW, w.data_ptr(),
                 _ptr(bias), ksize, stride, dilation, act, shuffle_r, int(grid_limit), int(residual_group), 1 if static_weights else 0)
    lib = _lib.load_library()
    if tensor_core:
        _lib.check(lib.dbsr_conv2d_tc(ctypes.byref(d), _stream()), 'dbsr_conv2d_tc')
    else:
        _lib.check(lib.dbsr_conv2d_direct(ctypes.byref(d), _stream()), 'dbsr_conv2d_direct')
    return y


def conv2d_tc_predictor(x: Act, w: torch.Tensor, bias: Optional[torch.Tensor], y: Act, ksize: int, act: int,
                        residual: Optional[Act], pred_w, pred_b, pred: torch.Tensor, grid_limit: int = 0,
                        static_weights: bool = False) -> torch.Tensor:
    """tcgen05 conv whose epilogue applies the 1x1 predictor + ReLU and writes `pred` [n, k, h, w] directly (the conv output
    map `y` is not written; it only describes the geometry): fp32, or -- when `pred` is an int16 tensor -- the reference's
    14-bit quantisation (min(value, 1) * 2^14, truncated).  pred_w [k][Cout] / pred_b [k]: HOST values (CPU
    tensors, lists or ctypes float arrays) -- they are copied into the kernel parameters."""
    if not isinstance(pred_w, ctypes.Array):
        pw = torch.as_tensor(pred_w, dtype=torch.float32).cpu().reshape(-1, y.c)
        pred_w = (ctypes.c_float * pw.numel())(*pw.reshape(-1).tolist())
    if not isinstance(pred_b, ctypes.Array):
        pb = torch.as_tensor(pred_b, dtype=torch.float32).cpu().reshape(-1)
        pred_b = (ctypes.c_float * pb.numel())(*pb.tolist())
    k = len(pred_b)
    assert len(pred_w) == k * y.c
    assert pred.dtype in (torch.float32, torch.int16) and pred.is_contiguous() and tuple(pred.shape) == (x.n, k, x.h, x.w)
    d = ConvDesc(x.view(), y.view(), residual.view() if residual is not None else _NULL_VIEW, w.data_ptr(),
                 _ptr(bias), ksize, 1, 1, act, 0, int(grid_limit), 0, 1 if static_weights else 0)
    _lib.check(_lib.load_library().dbsr_conv2d_tc_predictor(ctypes.byref(d), ctypes.cast(pred_w, ctypes.c_void_p),
                                                            ctypes.cast(pred_b, ctypes.c_void_p), k, pred.data_ptr(),
                                                            1 if pred.dtype == torch.int16 else 0, _stream()),
               'dbsr_conv2d_tc_predictor')
    return pred


def _host_floats(v, n=None):
    if isinstance(v, ctypes.Array):
        return v
    t = torch.as_tensor(v, dtype=torch.float32).cpu().reshape(-1)
    assert n is None or t.numel() == n
    return (ctypes.c_float * t.numel())(*t.tolist())


def _resblock_desc(x: Act, y: Optional[Act], w1, b1, w2, b2, pred_w=None, pred_b=None, pred=None, grid_limit=0, static_weights=False):
    keep = []
    pw = pb = None
    k = 0
    if pred is not None:
        pb = _host_floats(pred_b)
        k = len(pb)
        pw = _host_floats(pred_w, k * 32)
        keep = [pw, pb]
        assert pred.dtype in (torch.float32, torch.int16) and pred.is_contiguous() and tuple(pred.shape) == (x.n, k, x.h, x.w)
    d = ResBlockDesc(x.view(), y.view() if y is not None else _NULL_VIEW, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                     ctypes.cast(pw, ctypes.c_void_p) if pw is not None else None,
                     ctypes.cast(pb, ctypes.c_void_p) if pb is not None else None,
                     None if pred is None else pred.data_ptr(), k, 1 if (pred is not None and pred.dtype == torch.int16) else 0,
                     int(grid_limit), 1 if static_weights else 0)
    return d, keep


def resblock32_tc(x: Act, y: Optional[Act], w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                  pred_w=None, pred_b=None, pred: Optional[torch.Tensor] = None, grid_limit: int = 0,
                  static_weights: bool = False):
    """y = relu(x + conv2(relu(conv1(x) + b1)) + b2) for 32-channel bf16 maps in one tcgen05 launch (the intermediate map stays
    on the SM); w1 / w2: packed like `conv2d(..., tensor_core=True)` weights.  With `pred` (fp32 or int16 [n, k, h, w]) the 1x1
    predictor + ReLU is applied in the epilogue and `y` is not written."""
    d, _keep = _resblock_desc(x, y, w1, b1, w2, b2, pred_w, pred_b, pred, grid_limit, static_weights)
    _lib.check(_lib.load_library().dbsr_resblock32_tc(ctypes.byref(d), _stream()), 'dbsr_resblock32_tc')
    return pred if pred is not None else y


def resblock32_tc_supported(x: Act, y: Optional[Act], w1, b1, w2, b2, with_pred: bool = False) -> bool:
    d = ResBlockDesc(x.view(), y.view() if y is not None else _NULL_VIEW, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                     None, None, None, 0, 0, 0, 0)
    if with_pred:      # geometry / alignment only: the predictor arguments are checked by the call itself
        d.y = _NULL_VIEW
        d.pred, d.pred_w, d.pred_b, d.pred_c = 1, 1, 1, 3
    return bool(_lib.load_library().dbsr_resblock32_tc_supported(ctypes.byref(d)))


def quantize_q14(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst = (src.clamp(0, 1) * 2 ** 14).short()  -- the reference's 14-bit output quantisation (compute_score.py:110-111)"""
    require_device(src)
    assert src.dtype == torch.float32 and dst.dtype == torch.int16 and src.is_contiguous() and dst.is_contiguous()
    assert src.numel() == dst.numel()
    _lib.check(_lib.load_library().dbsr_quantize_q14(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), 'dbsr_quantize_q14')
    return dst


def conv2d_tc_supported(x: Act, w: torch.Tensor, bias, y: Act, ksize: int, stride: int = 1, dilation: int = 1,
                        residual: Optional[Act] = None, shuffle_r: int = 0, residual_group: int = 0) -> bool:
    d = ConvDesc(x.view(), y.view(), residual.view() if residual is not None else _NULL_VIEW, w.data_ptr(),
                 _ptr(bias), ksize, stride, dilation, 0, shuffle_r, 0, int(residual_group), 0)
    return bool(_lib.load_library().dbsr_conv2d_tc_supported(ctypes.byref(d)))


def conv2d_tc_geometry(cin: int, cout: int):
    """(ck, kpad, n_tile, cout_pad) the tcgen05 kernel uses for a Cin -> Cout convolution (single source of truth for
    the weight packer)."""
    a, b, c, d = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib.load_library().dbsr_conv2d_tc_geometry(cin, cout, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c),
                                                           ctypes.byref(d)), 'dbsr_conv2d_tc_geometry')
    return a.value, b.value, c.value, d.value


def space_to_depth2(x: Act, y: Act) -> Act:
    xv, yv = x.view(), y.view()
    _lib.check(_lib.load_library().dbsr_space_to_depth2(ctypes.byref(xv), ctypes.byref(yv), _stream()), 'dbsr_space_to_depth2')
    return y


def deconv4x4s2(x: Act, w: torch.Tensor, bias: torch.Tensor, y: Act, y2: Optional[Act] = None) -> Act:
    xv, yv = x.view(), y.view()
    y2v = y2.view() if y2 is not None else _NULL_VIEW
    _lib.check(_lib.load_library().dbsr_deconv4x4s2(ctypes.byref(xv), w.data_ptr(), bias.data_ptr(), ctypes.byref(yv),
                                                    ctypes.byref(y2v), _stream()), 'dbsr_deconv4x4s2')
    return y


def deconv_col2im(taps: Act, bias_t: torch.Tensor, y_t: Act, flow: Optional[Act] = None, wf: Optional[torch.Tensor] = None,
                  bias_f: Optional[torch.Tensor] = None, y_f: Optional[Act] = None, y_f2: Optional[Act] = None,
                  flow_taps: Optional[Act] = None, flow_bias: Optional[torch.Tensor] = None) -> Act:
    """scatter half of ConvTranspose2d(4, 2, 1) after a 1x1 conv produced the 32 (tap, oc) planes (+ the 2-channel
    flow deconvolution of the same level when `flow` is given -- or `flow_taps` [n, h, w, 18] + `flow_bias`: the coarser
    flow as the tap planes of its 3x3 convolution, summed on the fly)"""
    tv, yv = taps.view(), y_t.view()
    fv = flow.view() if flow is not None else _NULL_VIEW
    ftv = flow_taps.view() if flow_taps is not None else _NULL_VIEW
    yfv = y_f.view() if y_f is not None else _NULL_VIEW
    yf2v = y_f2.view() if y_f2 is not None else _NULL_VIEW
    _lib.check(_lib.load_library().dbsr_deconv_col2im_ftaps(ctypes.byref(tv), bias_t.data_ptr(), ctypes.byref(yv), ctypes.byref(fv),
                                                            ctypes.byref(ftv), _ptr(flow_bias), _ptr(wf), _ptr(bias_f),
                                                            ctypes.byref(yfv), ctypes.byref(yf2v), _stream()),
               'dbsr_deconv_col2im_ftaps')
    return y_t


def flow_from_taps(flow_taps: Act, bias: torch.Tensor, y: Act) -> Act:
    """y[n, y, x, oc] = bias[oc] + sum_taps flow_taps[n, y + ky - 1, x + kx - 1, (ky * 3 + kx) * 2 + oc] (zero outside the map):
    the second half of a 3x3 convolution with two output channels whose channel contraction ran as a 1x1 convolution"""
    fv, yv = flow_taps.view(), y.view()
    _lib.check(_lib.load_library().dbsr_flow_from_taps(ctypes.byref(fv), bias.data_ptr(), ctypes.byref(yv), _stream()),
               'dbsr_flow_from_taps')
    return y


def corr81(f1: Act, f2: Act, out: Act, pairs: int, group: int = 0, flow: Optional[Act] = None, flow_scale: float = 0.0,
           act: int = ACT_NONE, tensor_core: bool = True, f1_copy: Optional[Act] = None) -> Act:
    """tensor_core=False (per call) keeps bf16 maps on the CUDA-core kernels instead of the mma.sync banded product (A/B).
    f1_copy: [pairs, h, w, C] view that also receives the first map of every pair (the `tenFirst` slice of the decoder's concat
    buffer, pwcnet.py:173) -- written by the same launch where the kernel stages whole channel groups"""
    a, b, o = f1.view(), f2.view(), out.view()
    fl = flow.view() if flow is not None else _NULL_VIEW
    cp = f1_copy.view() if f1_copy is not None else _NULL_VIEW
    _lib.check(_lib.load_library().dbsr_corr81_copy(ctypes.byref(a), ctypes.byref(b), ctypes.byref(fl), float(flow_scale),
                                                    ctypes.byref(o), ctypes.byref(cp), pairs, group, act,
                                                    _lib.CORR_AUTO if tensor_core else _lib.CORR_CUDA_CORES, _stream()),
               'dbsr_corr81_copy')
    return out


def copy_channels(src: Act, dst: Act, group: int = 0, src_group: int = 0, src_first: int = 0) -> Act:
    s, d = src.view(), dst.view()
    _lib.check(_lib.load_library().dbsr_copy_channels(ctypes.byref(s), ctypes.byref(d), group, src_group, src_first,
                                                      _stream()), 'dbsr_copy_channels')
    return dst


def prep_burst(burst: torch.Tensor, enc_in: Act, pwc_in: Act) -> None:
    assert burst.dtype == torch.float32 and burst.is_contiguous() and burst.dim() == 5 and burst.shape[2] == 4
    frames = burst.shape[0] * burst.shape[1]
    e, p = enc_in.view(), pwc_in.view()
    _lib.check(_lib.load_library().dbsr_prep_burst(burst.data_ptr(), frames, burst.shape[3], burst.shape[4],
                                                   ctypes.byref(e), ctypes.byref(p), _stream()), 'dbsr_prep_burst')


def prep_burst_s2d(burst: torch.Tensor, enc_in: Act, pwc_s2d: Act, Hp: int, Wp: int) -> None:
    """prep_burst for the bf16 PWC-Net path: the resized RGB image is written directly in the space-to-depth layout the
    extractor's first (stride-2) convolution consumes, bf16 [frames, Hp/2, Wp/2, 12 (+4 pad)]"""
    assert burst.dtype == torch.float32 and burst.is_contiguous() and burst.dim() == 5 and burst.shape[2] == 4
    frames = burst.shape[0] * burst.shape[1]
    e, p = enc_in.view(), pwc_s2d.view()
    _lib.check(_lib.load_library().dbsr_prep_burst_s2d(burst.data_ptr(), frames, burst.shape[3], burst.shape[4], Hp, Wp,
                                                       ctypes.byref(e), ctypes.byref(p), _stream()), 'dbsr_prep_burst_s2d')


def flow_head(flow4: Act, offsets: torch.Tensor, H: int, W: int, Hp: int, Wp: int) -> torch.Tensor:
    assert offsets.dtype == torch.float32 and offsets.is_contiguous() and tuple(offsets.shape) == (flow4.n, 2, H, W)
    v = flow4.view()
    _lib.check(_lib.load_library().dbsr_flow_head(ctypes.byref(v), offsets.data_ptr(), H, W, Hp, Wp, _stream()),
               'dbsr_flow_head')
    return offsets


def warp(feat: Act, offsets: torch.Tensor, out: Act, frames: int = 0) -> Act:
    assert offsets.dtype == torch.float32 and offsets.is_contiguous()
    f, o = feat.view(), out.view()
    _lib.check(_lib.load_library().dbsr_warp(ctypes.byref(f), offsets.data_ptr(), ctypes.byref(o), frames, _stream()),
               'dbsr_warp')
    return out


def offsets_mod(offsets: torch.Tensor, out: Act, bursts: int, frames: int, modulo: float) -> Act:
    assert offsets.dtype == torch.float32 and offsets.is_contiguous()
    o = out.view()
    _lib.check(_lib.load_library().dbsr_offsets_mod(offsets.data_ptr(), ctypes.byref(o), bursts, frames, float(modulo),
                                                    _stream()), 'dbsr_offsets_mod')
    return out


def build_wp_input(proj: Act, wp_in: Act, frames: int) -> Act:
    p, w = proj.view(), wp_in.view()
    _lib.check(_lib.load_library().dbsr_build_wp_input(ctypes.byref(p), ctypes.byref(w), frames, _stream()),
               'dbsr_build_wp_input')
    return wp_in


def warp_proj(q: Act, bias: torch.Tensor, wp_in: Act, frames: int, offsets: Optional[torch.Tensor] = None) -> Act:
    if offsets is not None:
        assert offsets.dtype == torch.float32 and offsets.is_contiguous()
    qv, wv = q.view(), wp_in.view()
    _lib.check(_lib.load_library().dbsr_warp_proj(ctypes.byref(qv), bias.data_ptr(), _ptr(offsets), ctypes.byref(wv),
                                                  frames, _stream()), 'dbsr_warp_proj')
    return wp_in


def warp_proj_split(q: Act, bias: torch.Tensor, wp_in: Act, p0: Act, frames: int, offsets: Optional[torch.Tensor] = None) -> Act:
    """wp_in[:, 0:C] = p_n = relu(warp(q_n) + bias) for every frame, p0[b] = p_0 of every burst (split weight predictor input)"""
    if offsets is not None:
        assert offsets.dtype == torch.float32 and offsets.is_contiguous()
    qv, wv, pv = q.view(), wp_in.view(), p0.view()
    _lib.check(_lib.load_library().dbsr_warp_proj_split(ctypes.byref(qv), bias.data_ptr(), _ptr(offsets), ctypes.byref(wv),
                                                        ctypes.byref(pv), frames, _stream()), 'dbsr_warp_proj_split')
    return wp_in


def softmax_wsum(feat: Act, logits: Act, fused: Act, frames: int, offsets: Optional[torch.Tensor] = None,
                 weights_out: Optional[torch.Tensor] = None) -> Act:
    if offsets is not None:
        assert offsets.dtype == torch.float32 and offsets.is_contiguous()
    if weights_out is not None:
        assert weights_out.dtype == torch.float32 and weights_out.is_contiguous()
    f, l, o = feat.view(), logits.view(), fused.view()
    _lib.check(_lib.load_library().dbsr_softmax_wsum(ctypes.byref(f), ctypes.byref(l), _ptr(offsets), ctypes.byref(o),
                                                     _ptr(weights_out), frames, _stream()), 'dbsr_softmax_wsum')
    return fused


def blur3x3(x: Act, y: Act, k9) -> Act:
    arr = (ctypes.c_float * 9)(*[float(v) for v in k9])
    xv, yv = x.view(), y.view()
    _lib.check(_lib.load_library().dbsr_blur3x3(ctypes.byref(xv), ctypes.byref(yv), arr, _stream()), 'dbsr_blur3x3')
    return y


def predictor(x: Act, w: torch.Tensor, bias: Optional[torch.Tensor], pred: torch.Tensor) -> torch.Tensor:
    cout = w.shape[0]
    assert pred.dtype == torch.float32 and pred.is_contiguous() and tuple(pred.shape) == (x.n, cout, x.h, x.w)
    v = x.view()
    _lib.check(_lib.load_library().dbsr_predictor(ctypes.byref(v), w.data_ptr(), _ptr(bias), cout, pred.data_ptr(),
                                                  _stream()), 'dbsr_predictor')
    return pred


# ---- evaluation metrics (include/dbsr_b200.h "evaluation metrics") ------------------------------------------------------
def _check_image_pair(a: torch.Tensor, b: torch.Tensor) -> None:
    require_device(a)
    require_device(b)
    if a.dim() != 4 or a.shape != b.shape:
        raise ValueError(f'expected two [n, c, h, w] tensors of the same shape, got {tuple(a.shape)} / {tuple(b.shape)}')
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise TypeError('the metric kernels take fp32 images')
    assert a.is_contiguous() and b.is_contiguous()


def _mask_bytes(valid: Optional[torch.Tensor], like: torch.Tensor) -> Optional[torch.Tensor]:
    """[n, 1, h, w] bool / uint8 mask -> contiguous uint8 view for the metric kernels"""
    if valid is None:
        return None
    require_device(valid)
    n, _, h, w = like.shape
    if tuple(valid.shape) != (n, 1, h, w) or valid.dtype not in (torch.bool, torch.uint8):
        raise ValueError(f'valid mask must be a bool / uint8 [n, 1, h, w] tensor, got {valid.dtype} {tuple(valid.shape)}')
    return valid.contiguous().view(torch.uint8)


def ssim_stats(img1: torch.Tensor, img2: torch.Tensor, window1d, crop: int = 0, val_range: Optional[float] = None,
               want_map: bool = False, valid: Optional[torch.Tensor] = None):
    """Fused SSIM of two fp32 NCHW batches: returns (stats [n, 2] = per-image mean ssim / mean contrast term, map or None).
    window1d: the 1-D Gaussian of msssim.gaussian (python floats, len 1..11); val_range None -> derived on the device.
    valid ([n, 1, h, w] mask, window 11): stats become (sum ssim * valid, sum valid) per image, un-normalised."""
    _check_image_pair(img1, img2)
    vb = _mask_bytes(valid, img1)
    n, c, h, w = img1.shape
    k = len(window1d)
    lib = _lib.load_library()
    floats = lib.dbsr_ssim_workspace_floats(n, c, h, w, crop, k)
    if floats < 0:
        raise ValueError(f'ssim: image {h}x{w} with boundary_ignore {crop} and a {k}-tap window has no valid window position')
    ws = torch.empty(floats, dtype=torch.float32, device=img1.device)
    stats = torch.empty(n, 2, dtype=torch.float32, device=img1.device)
    smap = torch.empty(n, c, h - 2 * crop - k + 1, w - 2 * crop - k + 1, dtype=torch.float32, device=img1.device) if want_map else None
    arr = (ctypes.c_float * k)(*[float(v) for v in window1d])
    _lib.check(lib.dbsr_ssim(img1.data_ptr(), img2.data_ptr(), n, c, h, w, crop, arr, k, float(val_range) if val_range else 0.0,
                             _ptr(vb), ws.data_ptr(), stats.data_ptr(), _ptr(smap), _stream()), 'dbsr_ssim')
    return stats, smap


def avgpool2_pair(img1: torch.Tensor, img2: torch.Tensor):
    """(F.avg_pool2d(img1, (2, 2)), F.avg_pool2d(img2, (2, 2))) in one launch (msssim.py:88-89)."""
    _check_image_pair(img1, img2)
    n, c, h, w = img1.shape
    o1 = torch.empty(n, c, h // 2, w // 2, dtype=torch.float32, device=img1.device)
    o2 = torch.empty_like(o1)
    _lib.check(_lib.load_library().dbsr_avgpool2_pair(img1.data_ptr(), img2.data_ptr(), o1.data_ptr(), o2.data_ptr(), n * c, h, w,
                                                      _stream()), 'dbsr_avgpool2_pair')
    return o1, o2


def mse_per_image(pred: torch.Tensor, gt: torch.Tensor, crop: int = 0, valid: Optional[torch.Tensor] = None, raw: bool = False) -> torch.Tensor:
    """[n] mean squared error of each image over its interior (boundary_ignore = crop), one launch for the batch.
    valid ([n, 1, h, w] mask): the masked form (err * valid).sum() / (valid.sum() * C + 1e-12) of image_quality_v2.py:60-64;
    raw=True returns its two sums per image, [n, 2], for a batch-wide ratio."""
    _check_image_pair(pred, gt)
    vb = _mask_bytes(valid, pred)
    n, c, h, w = pred.shape
    lib = _lib.load_library()
    ws = torch.empty(lib.dbsr_mse_workspace_floats(n), dtype=torch.float32, device=pred.device)
    out = torch.empty((n, 2) if vb is not None else (n,), dtype=torch.float32, device=pred.device)
    _lib.check(lib.dbsr_mse_per_image(pred.data_ptr(), gt.data_ptr(), _ptr(vb), n, c, h, w, crop, ws.data_ptr(), out.data_ptr(),
                                      _stream()), 'dbsr_mse_per_image')
    if vb is None or raw:
        return out
    return out[:, 0] / (out[:, 1] + 1e-12)


# ---- synthetic burst generation: inverse camera pipeline (include/dbsr_b200.h) ------------------------------------------------
def unprocess_rgb(image: torch.Tensor, rgb2cam: torch.Tensor, gains3, smoothstep: bool = True, gamma: bool = True) -> torch.Tensor:
    """sRGB [3, h, w] / [b, 3, h, w] -> clamped linear camera RGB (inverse tone curve, gamma, CCM, safe gain inversion), one pass"""
    require_device(image)
    assert image.dtype == torch.float32 and image.dim() in (3, 4) and image.shape[-3] == 3
    x = image.contiguous()
    out = torch.empty_like(x)
    b = 1 if x.dim() == 3 else x.shape[0]
    ccm = (ctypes.c_float * 9)(*[float(v) for v in rgb2cam.detach().float().cpu().flatten().tolist()])
    g3 = (ctypes.c_float * 3)(*[float(v) for v in gains3])
    _lib.check(_lib.load_library().dbsr_unprocess_rgb(x.data_ptr(), out.data_ptr(), b, x.shape[-2], x.shape[-1], ccm, g3,
                                                      1 if smoothstep else 0, 1 if gamma else 0, _stream()), 'dbsr_unprocess_rgb')
    return out


def mosaic_noise(rgb: torch.Tensor, shot_noise: float = 0.0, read_noise: float = 0.0, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """RGB burst [n, 3, h, w] -> clamped noisy RGGB burst [n, 4, h/2, w/2]; `noise`: standard-normal tensor of that shape or None"""
    require_device(rgb)
    assert rgb.dtype == torch.float32 and rgb.dim() == 4 and rgb.shape[1] == 3
    x = rgb.contiguous()
    n, _, h, w = x.shape
    raw = torch.empty(n, 4, h // 2, w // 2, dtype=torch.float32, device=x.device)
    if noise is not None:
        require_device(noise)
        assert noise.dtype == torch.float32 and tuple(noise.shape) == tuple(raw.shape) and noise.is_contiguous()
    _lib.check(_lib.load_library().dbsr_mosaic_noise(x.data_ptr(), _ptr(noise), raw.data_ptr(), n, h, w, float(shot_noise),
                                                     float(read_noise), _stream()), 'dbsr_mosaic_noise')
    return raw


def single2lrburst(image: torch.Tensor, inverse_maps: torch.Tensor, position_maps: torch.Tensor, factor: int, border_crop: int = 0,
                   normalize: bool = True, want_flow: bool = True):
    """image [3, H, W] fp32 -> (burst [n, 3, h, w], flow [n, 2, h, w] or None): quantise, warp, crop, down-sample (OpenCV-exact)"""
    require_device(image)
    assert image.dtype == torch.float32 and image.dim() == 3 and image.shape[0] == 3
    x = image.contiguous()
    n = inverse_maps.shape[0]
    inv = inverse_maps.to(device=x.device, dtype=torch.float64).contiguous()
    pos = position_maps.to(device=x.device, dtype=torch.float32).contiguous()
    assert tuple(inv.shape) == (n, 6) and tuple(pos.shape) == (n, 6)
    _, H, W = x.shape
    hc, wc = H - 2 * border_crop, W - 2 * border_crop
    if hc <= 0 or wc <= 0 or hc % factor or wc % factor:
        raise ValueError(f'single2lrburst: cropped size {hc}x{wc} must be a positive multiple of the down-sampling factor {factor}')
    burst = torch.empty(n, 3, hc // factor, wc // factor, dtype=torch.float32, device=x.device)
    flow = torch.empty(n, 2, hc // factor, wc // factor, dtype=torch.float32, device=x.device) if want_flow else None
    _lib.check(_lib.load_library().dbsr_single2lrburst(x.data_ptr(), H, W, inv.data_ptr(), pos.data_ptr(), n, factor, border_crop,
                                                       1 if normalize else 0, burst.data_ptr(), _ptr(flow), _stream()),
               'dbsr_single2lrburst')
    return burst, flow


# ---- batched generator kernels: per-burst parameters live in device memory (one upload per batch) --------------------------------
def unprocess_rgb_batch(images: torch.Tensor, params12: torch.Tensor, smoothstep: bool = True, gamma: bool = True) -> torch.Tensor:
    """images [b, 3, h, w]; params12 [b, 12] device fp32 = rgb2cam (row-major) + [1 / red, 1, 1 / blue] / rgb_gain per image"""
    require_device(images)
    assert images.dtype == torch.float32 and images.dim() == 4 and images.shape[1] == 3
    x = images.contiguous()
    assert params12.is_cuda and params12.dtype == torch.float32 and tuple(params12.shape) == (x.shape[0], 12) and params12.is_contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.load_library().dbsr_unprocess_rgb_batch(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[2], x.shape[3],
                                                            params12.data_ptr(), 1 if smoothstep else 0, 1 if gamma else 0, _stream()),
               'dbsr_unprocess_rgb_batch')
    return out


def single2lrburst_batch(images: torch.Tensor, inverse_maps: torch.Tensor, position_maps: torch.Tensor, frames: int, factor: int,
                         border_crop: int = 0, normalize: bool = True, want_flow: bool = True):
    """images [b, 3, H, W] -> (burst [b * frames, 3, h, w], flow [b * frames, 2, h, w] or None); maps: DEVICE [b * frames, 6]"""
    require_device(images)
    assert images.dtype == torch.float32 and images.dim() == 4 and images.shape[1] == 3
    x = images.contiguous()
    b, _, H, W = x.shape
    assert inverse_maps.is_cuda and inverse_maps.dtype == torch.float64 and tuple(inverse_maps.shape) == (b * frames, 6)
    assert position_maps.is_cuda and position_maps.dtype == torch.float32 and tuple(position_maps.shape) == (b * frames, 6)
    hc, wc = H - 2 * border_crop, W - 2 * border_crop
    if hc <= 0 or wc <= 0 or hc % factor or wc % factor:
        raise ValueError(f'single2lrburst: cropped size {hc}x{wc} must be a positive multiple of the down-sampling factor {factor}')
    burst = torch.empty(b * frames, 3, hc // factor, wc // factor, dtype=torch.float32, device=x.device)
    flow = torch.empty(b * frames, 2, hc // factor, wc // factor, dtype=torch.float32, device=x.device) if want_flow else None
    _lib.check(_lib.load_library().dbsr_single2lrburst_batch(x.data_ptr(), b, H, W, inverse_maps.contiguous().data_ptr(),
                                                             position_maps.contiguous().data_ptr(), frames, factor, border_crop,
                                                             1 if normalize else 0, burst.data_ptr(), _ptr(flow), _stream()),
               'dbsr_single2lrburst_batch')
    return burst, flow


def mosaic_noise_batch(rgb: torch.Tensor, levels: torch.Tensor, frames_per_burst: int, noise: torch.Tensor) -> torch.Tensor:
    """rgb [n, 3, h, w] -> noisy RGGB [n, 4, h/2, w/2]; levels [n / frames_per_burst, 2] device fp32 = (shot, read) per burst"""
    require_device(rgb)
    x = rgb.contiguous()
    n, _, h, w = x.shape
    raw = torch.empty(n, 4, h // 2, w // 2, dtype=torch.float32, device=x.device)
    assert levels.is_cuda and levels.dtype == torch.float32 and tuple(levels.shape) == (n // frames_per_burst, 2) and levels.is_contiguous()
    assert noise.is_cuda and noise.dtype == torch.float32 and tuple(noise.shape) == tuple(raw.shape) and noise.is_contiguous()
    _lib.check(_lib.load_library().dbsr_mosaic_noise_batch(x.data_ptr(), noise.data_ptr(), raw.data_ptr(), n, h, w, levels.data_ptr(),
                                                           frames_per_burst, _stream()), 'dbsr_mosaic_noise_batch')
    return raw
