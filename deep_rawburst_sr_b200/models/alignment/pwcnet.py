"""PWC-Net optical flow with the reference's interface (models/alignment/pwcnet.py:41-281): `PWCNet(load_pretrained,
weights_path, rgb2bgr)`, attribute `.net` (a `Network` with the reference's parameter names), `forward(source_img,
target_img) -> flow [B, 2, H, W]` mapping target pixels to source pixels, (x, y) order.

Execution is the sm_100a kernel sequence of `DBSREngine` (cost volume with fused backwarp + LeakyReLU, in-place dense
concatenation, transposed convs, dilated refiner, flow head); `backwarp` and `correlation.FunctionCorrelation` are also
exposed with the reference's signatures.
"""
import math

import torch

from ... import ops
from ...engine import DBSREngine
from ...external.pwcnet.correlation import correlation  # noqa: F401  (same import the reference performs)
from ..engine_owner import EngineOwner


def backwarp(tenInput, tenFlow):
    """pwcnet.py:16-38: sample tenInput at x + flow_x * W/(W-1) (the reference's linspace grid is pixel centres and the
    flow is divided by (W-1)/2), zero where the sampled ones-channel is <= 0.999.  Runs `dbsr_corr81`'s fused
    staging path is internal; standalone this uses the same arithmetic through the warp kernel + mask."""
    ops.require_device(tenInput)
    n, c, h, w = tenInput.shape
    assert h > 1 and w > 1
    scale = torch.tensor([w / (w - 1.0), h / (h - 1.0)], device=tenFlow.device, dtype=torch.float32).view(1, 2, 1, 1)
    flow = (tenFlow.float() * scale).contiguous()
    cp = (c + 1 + 3) // 4 * 4
    src = ops.Act.empty(n, h, w, cp, torch.float32, tenInput.device, zero=True)
    src.slice(0, c).from_nchw(tenInput.contiguous().float())
    src.buf[..., c] = 1.0
    dst = ops.Act.empty(n, h, w, cp, torch.float32, tenInput.device)
    ops.warp(src, flow, dst, frames=0)
    out = dst.slice(0, c + 1).to_nchw()
    mask = (out[:, -1:] > 0.999).to(out.dtype)
    return out[:, :-1] * mask


class Network(EngineOwner, torch.nn.Module):
    """The reference's module tree (pwcnet.py:41-219) as a parameter container; `forward` runs the sm_100a kernel sequence."""

    def __init__(self):
        super(Network, self).__init__()
        self.precision = 'fp32'       # 'fp32': exact CUDA-core path; 'bf16': tcgen05 tensor cores (PWCNet.set_precision)
        L = torch.nn.LeakyReLU
        C = torch.nn.Conv2d

        class Extractor(torch.nn.Module):
            def __init__(self):
                super(Extractor, self).__init__()
                chans = [3, 16, 32, 64, 96, 128, 196]
                for i, name in enumerate(['One', 'Two', 'Thr', 'Fou', 'Fiv', 'Six']):
                    ci, co = chans[i], chans[i + 1]
                    setattr(self, 'net' + name, torch.nn.Sequential(
                        C(ci, co, 3, 2, 1), L(inplace=False, negative_slope=0.1),
                        C(co, co, 3, 1, 1), L(inplace=False, negative_slope=0.1),
                        C(co, co, 3, 1, 1), L(inplace=False, negative_slope=0.1)))

        class Decoder(torch.nn.Module):
            def __init__(self, intLevel):
                super(Decoder, self).__init__()
                tab = [None, None, 81 + 32 + 2 + 2, 81 + 64 + 2 + 2, 81 + 96 + 2 + 2, 81 + 128 + 2 + 2, 81, None]
                intPrevious, intCurrent = tab[intLevel + 1], tab[intLevel + 0]
                if intLevel < 6:
                    self.netUpflow = torch.nn.ConvTranspose2d(2, 2, kernel_size=4, stride=2, padding=1)
                    self.netUpfeat = torch.nn.ConvTranspose2d(intPrevious + 128 + 128 + 96 + 64 + 32, 2, kernel_size=4,
                                                              stride=2, padding=1)
                    self.fltBackwarp = [None, None, None, 5.0, 2.5, 1.25, 0.625, None][intLevel + 1]
                cin = intCurrent
                for name, co in zip(['One', 'Two', 'Thr', 'Fou', 'Fiv'], [128, 128, 96, 64, 32]):
                    setattr(self, 'net' + name, torch.nn.Sequential(C(cin, co, 3, 1, 1),
                                                                     L(inplace=False, negative_slope=0.1)))
                    cin += co
                self.netSix = torch.nn.Sequential(C(cin, 2, 3, 1, 1))

        class Refiner(torch.nn.Module):
            def __init__(self):
                super(Refiner, self).__init__()
                spec = [(565, 128, 1), (128, 128, 2), (128, 128, 4), (128, 96, 8), (96, 64, 16), (64, 32, 1), (32, 2, 1)]
                layers = []
                for i, (ci, co, d) in enumerate(spec):
                    layers.append(C(ci, co, 3, 1, d, d))
                    if i < len(spec) - 1:
                        layers.append(L(inplace=False, negative_slope=0.1))
                self.netMain = torch.nn.Sequential(*layers)

        self.netExtractor = Extractor()
        self.netTwo = Decoder(2)
        self.netThr = Decoder(3)
        self.netFou = Decoder(4)
        self.netFiv = Decoder(5)
        self.netSix = Decoder(6)
        self.netRefiner = Refiner()

    def engine(self, device):
        if not self._engine_is_current(device, precision=self.precision):
            self._set_engine(DBSREngine(self.state_dict(), device, precision=self.precision, pwc_prefix='', parts=('pwc',)))
        return self._engine

    @ops.tensor_device_guard
    @torch.no_grad()
    def forward(self, tenFirst, tenSecond):
        """pwcnet.py:220-231: extractor pyramids of both images, decoders 6..2, refiner -> flow [P, 2, H/4, W/4] (in units of
        1/20 px of the input grid, as the reference returns it).  H and W must be multiples of 64 (the reference's decoder
        concatenations only line up then; `PWCNet.forward` resizes to such a size first)."""
        ops.require_device(tenFirst)
        assert tenFirst.dim() == 4 and tenFirst.shape == tenSecond.shape and tenFirst.shape[1] == 3
        P, _, H, W = tenFirst.shape
        if H % 64 or W % 64:
            raise ValueError(f'Network.forward needs H, W multiples of 64 (got {H}x{W}); PWCNet.forward resizes for you')
        eng = self.engine(tenFirst.device)
        ws = eng.workspace(('net_pairs', P, H, W))
        pwc_in = eng._buf(ws, 'pwc_in', 2 * P, H, W, 4, torch.float32)
        pwc_in.slice(0, 3).from_nchw(torch.cat([tenFirst, tenSecond], 0).contiguous().float())
        feats = eng.pwc_extract(ws, pwc_in)
        flow4 = eng.pwc_decode(ws, [f.images(0, P) for f in feats], [f.images(P, P) for f in feats], P, 0, 0)
        return flow4.to_nchw()


class PWCNet(torch.nn.Module):
    def __init__(self, load_pretrained=True, weights_path=None, rgb2bgr=False):
        super(PWCNet, self).__init__()
        self.net = Network()
        self.rgb2bgr = rgb2bgr
        # 'fp32': exact CUDA-core path (flows within 1e-4 px of the reference); 'bf16': tcgen05 tensor cores (the precision
        # the burst forward uses for PWC-Net, ~1e-2 px) -- e.g. for the output-resolution alignment of the BurstSR metric
        self.precision = 'fp32'
        if load_pretrained:
            if weights_path is None:
                raise Exception
            weights_dict = torch.load(weights_path)
            self.net.load_state_dict({strKey.replace('module', 'net'): tenWeight for strKey, tenWeight
                                      in weights_dict.items()})

    def set_precision(self, precision: str):
        assert precision in ('fp32', 'bf16')
        self.precision = precision
        return self

    def engine(self, device):
        """the engine lives on `.net` (one set of packed weights whether `PWCNet.forward` or `Network.forward` is called)"""
        self.net.precision = self.precision
        return self.net.engine(device)

    @ops.tensor_device_guard
    @torch.no_grad()
    def forward(self, source_img, target_img):
        assert (source_img.shape[-1] == target_img.shape[-1])
        assert (source_img.shape[-2] == target_img.shape[-2])
        ops.require_device(source_img)
        W, H = source_img.shape[-1], source_img.shape[-2]
        source_img = source_img.reshape(-1, 3, H, W).float()
        target_img = target_img.reshape(-1, 3, H, W).float()
        if self.rgb2bgr:
            source_img = source_img[:, [2, 1, 0]]
            target_img = target_img[:, [2, 1, 0]]
        P = source_img.shape[0]
        Wp = int(math.floor(math.ceil(W / 64.0) * 64.0))
        Hp = int(math.floor(math.ceil(H / 64.0) * 64.0))
        eng = self.engine(source_img.device)
        ws = eng.workspace(('pwc_pairs', P, H, W))
        # images [0, P) = target (tenFirst), [P, 2P) = source (tenSecond)  (pwcnet.py:273)
        both = torch.cat([target_img, source_img], 0).contiguous()
        # packed 4-channel pseudo-RAW [R, G, G, B] so the burst-prep kernel (RGGB->RGB + resize) can be reused verbatim
        raw = torch.stack([both[:, 0], both[:, 1], both[:, 1], both[:, 2]], 1).unsqueeze(0).contiguous()
        enc_dummy = eng._buf(ws, 'enc_dummy', 2 * P, H, W, 4, torch.float32)
        pwc_in = eng._buf(ws, 'pwc_in', 2 * P, Hp, Wp, 4, torch.float32)
        ops.prep_burst(raw, enc_dummy, pwc_in)
        feats = eng.pwc_extract(ws, pwc_in)
        first = [f.images(0, P) for f in feats]
        second = [f.images(P, P) for f in feats]
        flow4 = eng.pwc_decode(ws, first, second, P, 0, 0)
        flow = torch.empty((P, 2, H, W), dtype=torch.float32, device=source_img.device)
        ops.flow_head(flow4, flow, H, W, Hp, Wp)
        return flow
