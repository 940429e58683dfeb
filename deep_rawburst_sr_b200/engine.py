"""Execution engine of the DBSR burst forward pass on B200.

Owns (i) the weights re-packed once for the kernels, (ii) per-shape channels-last workspaces in HBM, (iii) the
launch sequence that replaces `DBSRNet.forward` (reference models/dbsr/dbsrnet.py:33-38, call stack in
SURVEY.md 3.1).  The Python modules in `deep_rawburst_sr_b200.models` keep the reference's interfaces and
parameters and delegate here.  Every op is a hand-written sm_100a kernel behind the C ABI (`ops.py`); there is
no eager-PyTorch fallback for any of them.

Data layout in HBM: every activation is NHWC.  `precision='bf16'`: DBSR activations are bf16 and the 3x3 / 1x1
convolutions run on tcgen05 tensor cores with fp32 accumulation in TMEM; PWC-Net (flow) stays fp32.
`precision='fp32'`: everything fp32 on CUDA cores (exact path, <= 1e-4 on the output).
PWC-Net dense concatenations (pwcnet.py:171-177) are channel slices of one buffer per pyramid level, segments
aligned to 8 channels: [o5 32 | o4 64 | o3 96 | o2 128 | o1 128 | V 81(+7) | f1 C | upflow 2(+6) | upfeat 2(+6)].
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Optional

import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, Act

PWC_NAMES = ['One', 'Two', 'Thr', 'Fou', 'Fiv', 'Six']
PWC_EXT_CH = [3, 16, 32, 64, 96, 128, 196]
PWC_DEC_OUT = [128, 128, 96, 64, 32]           # netOne..netFiv outputs (netSix -> 2 flow channels)
PWC_BACKWARP = {5: 0.625, 4: 1.25, 3: 2.5, 2: 5.0}   # reference pwcnet.py:121
PWC_REFINER_DIL = [1, 2, 4, 8, 16, 1, 1]


def _align8(v: int) -> int:
    return (v + 7) // 8 * 8


class PwcLayout:
    """Channel layout of the level-l concat buffer (new channels in FRONT, reference pwcnet.py:173-177)."""

    def __init__(self, level: int):
        segs = [('o5', 32), ('o4', 64), ('o3', 96), ('o2', 128), ('o1', 128), ('V', 81)]
        if level < 6:
            segs += [('f1', PWC_EXT_CH[level]), ('upflow', 2), ('upfeat', 2)]
        self.level = level
        self.names = [s[0] for s in segs]
        self.sizes = {s[0]: s[1] for s in segs}
        self.off = {}
        o = 0
        for name, sz in segs:
            self.off[name] = o
            o += _align8(sz)
        self.total = o

    def chmap_from(self, first_seg: str):
        """buffer channel (relative to the slice start) of every original input channel of a conv that reads
        `cat[first_seg:]`; and the slice (start, length)."""
        i0 = self.names.index(first_seg)
        start = self.off[first_seg]
        cm = []
        for name in self.names[i0:]:
            cm += list(range(self.off[name] - start, self.off[name] - start + self.sizes[name]))
        return cm, start, self.total - start


# Weight packing runs on the HOST (CPU tensors in, one H2D copy per packed tensor out): building an engine issues no
# device kernels, so a profiler's launch list of a forward starts with the forward's own kernels.
def _host(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to('cpu', torch.float32)


def pack_direct(w: torch.Tensor, chmap=None, cin_buf: Optional[int] = None, device=None) -> torch.Tensor:
    """[Cout, Cin, k, k] -> fp32 [k*k, Cin_buf, Cout] (rows of padded channels are zero).  Packed on the host; the result
    lands on `device` (default: the device of `w`)."""
    device = w.device if device is None else device
    w = _host(w)
    cout, cin, kh, kw = w.shape
    cin_buf = cin if cin_buf is None else cin_buf
    out = torch.zeros((kh * kw, cin_buf, cout), dtype=torch.float32)
    src = w.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)
    if chmap is None:
        out[:, :cin] = src
    else:
        out[:, torch.as_tensor(chmap)] = src
    return out.contiguous().to(device)


def pack_tc(w: torch.Tensor, shuffle_r: int = 0, chmap=None, cin_buf: Optional[int] = None, device=None) -> torch.Tensor:
    """[Cout, Cin, k, k] -> bf16 [k*k, cout_pad, kpad] K-major for dbsr_conv2d_tc; (kpad, cout_pad) come from the
    kernel's own tiling rule (dbsr_conv2d_tc_geometry).  Padded rows / columns are zero.  `chmap`/`cin_buf`: input
    channel -> position inside a padded concat slice.  shuffle_r = 8: rows permuted to (i, j, c) order so that an N
    tile is contiguous in the pixel-shuffled output (nn.PixelShuffle: co = c*r*r + i*r + j)."""
    device = w.device if device is None else device
    w = _host(w)
    cout, cin, kh, kw = w.shape
    cin_buf = cin if cin_buf is None else cin_buf
    _ck, kpad, _nt, cout_pad = ops.conv2d_tc_geometry(cin_buf, cout)
    src = w
    if shuffle_r and shuffle_r > 1:
        r = shuffle_r
        c = cout // (r * r)
        src = src.view(c, r, r, cin, kh, kw).permute(1, 2, 0, 3, 4, 5).reshape(cout, cin, kh, kw)
    out = torch.zeros((kh * kw, cout_pad, kpad), dtype=torch.float32)
    src = src.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    if chmap is None:
        out[:, :cout, :cin] = src
    else:
        out[:, :cout, torch.as_tensor(chmap)] = src
    return out.to(torch.bfloat16).contiguous().to(device)


def permute_shuffle_rows(t: torch.Tensor, r: int) -> torch.Tensor:
    """reorder dim 0 from PixelShuffle order (c, i, j) to the tensor-core kernel's (i, j, c) order"""
    c = t.shape[0] // (r * r)
    return t.reshape(c, r, r, *t.shape[1:]).permute(1, 2, 0, *range(3, t.dim() + 2)).reshape(t.shape).contiguous()


def pack_s2d_weight(w: torch.Tensor) -> torch.Tensor:
    """3x3 / stride-2 / pad-1 conv weight [Cout, C, 3, 3] -> the equivalent 3x3 / stride-1 / pad-1 weight
    [Cout, 4C, 3, 3] over the space-to-depth input (channel (p*2 + q)*C + c = x[2Y + p, 2X + q, c]):
    input row 2Y + dy is s2d row Y - 1, parity 1 for dy = -1; row Y, parity dy for dy = 0, 1 (same along x)."""
    cout, c = w.shape[0], w.shape[1]
    out = torch.zeros((cout, 4 * c, 3, 3), dtype=w.dtype, device=w.device)
    pos = {-1: (0, 1), 0: (1, 0), 1: (1, 1)}      # d -> (tap index in the s2d kernel, parity)
    for dy in (-1, 0, 1):
        ky, p = pos[dy]
        for dx in (-1, 0, 1):
            kx, q = pos[dx]
            out[:, (p * 2 + q) * c:(p * 2 + q + 1) * c, ky, kx] = w[:, :, dy + 1, dx + 1]
    return out


def pack_deconv(w: torch.Tensor, chmap=None, cin_buf: Optional[int] = None, device=None) -> torch.Tensor:
    """ConvTranspose2d weight [Cin, 2, 4, 4] -> fp32 [4, 4, 2, Cin_buf]."""
    device = w.device if device is None else device
    w = _host(w)
    cin = w.shape[0]
    cin_buf = cin if cin_buf is None else cin_buf
    out = torch.zeros((4, 4, 2, cin_buf), dtype=torch.float32)
    src = w.permute(2, 3, 1, 0)
    if chmap is None:
        out[..., :cin] = src
    else:
        out[..., torch.as_tensor(chmap)] = src
    return out.contiguous().to(device)


class ConvW:
    __slots__ = ('direct', 'tc', 'bias', 'bias_tc', 'ksize', 'cout', 'cin', 'shuffle_r')

    def __init__(self, direct, tc, bias, ksize, cout, cin, shuffle_r=0):
        self.direct, self.tc, self.bias, self.ksize, self.cout, self.cin, self.shuffle_r = \
            direct, tc, bias, ksize, cout, cin, shuffle_r
        self.bias_tc = bias
        if bias is not None and shuffle_r and shuffle_r > 1:   # the tensor-core path stores in (i, j, c) row order
            self.bias_tc = permute_shuffle_rows(bias.cpu(), shuffle_r).to(bias.device)


def on_engine_device(fn):
    """run a DBSREngine method with the engine's device current: the ctypes layer launches on `torch.cuda.current_stream()` of
    the CURRENT device and the library keeps per-device caches, so a model on cuda:1 must not launch while cuda:0 is current"""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        if torch.cuda.current_device() == (self.device.index if self.device.index is not None else torch.cuda.current_device()):
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapper


class DBSREngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device, precision: str = 'bf16', offset_modulo: float = 1.0,
                 gauss_kernel=None, logits_fp32: bool = False, pwc_precision: Optional[str] = None,
                 pwc_prefix: str = 'encoder.alignment_net.net.', parts=('pwc', 'encoder', 'merging', 'decoder')):
        assert precision in ('bf16', 'fp32')
        self.device = torch.device(device)
        ops.require_device(torch.empty(1, device=self.device))
        self.precision = precision
        self.bf16 = precision == 'bf16'
        self.act_dtype = torch.bfloat16 if self.bf16 else torch.float32
        self.logits_dtype = torch.float32 if (logits_fp32 or not self.bf16) else torch.bfloat16
        # PWC-Net activations: bf16 + tensor cores with the bf16 path unless pwc_precision='fp32' (flows stay fp32)
        self.pwc_precision = pwc_precision or precision
        assert self.pwc_precision in ('bf16', 'fp32')
        self.pwc_dtype = torch.bfloat16 if self.pwc_precision == 'bf16' else torch.float32
        self.offset_modulo = float(offset_modulo) if offset_modulo is not None else 0.0
        self.pwc_prefix = pwc_prefix
        self.W: Dict[str, ConvW] = {}
        self.D: Dict[str, tuple] = {}
        self._ws: Dict[tuple, dict] = {}
        self._graphs: Dict[tuple, tuple] = {}
        self._side = None                 # second stream of the alignment / encoder overlap
        self.overlap_alignment = True
        # decoder post-res blocks (32 channels at 8H x 8W): one fused launch per block (csrc/resblock_tc.cu) instead of two
        # conv launches; bit-identical results.  False = A/B against the two-launch path.
        self.fuse_hr_resblocks = True
        # persistent-grid size of the encoder convs while the alignment stream is active: leaving ~1/6 of the SMs to the
        # short PWC-Net launches is worth another 1-1.5 % per step on B200 (148 SMs -> 124; measured 9.39 -> 9.25 ms)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.encoder_grid_limit = sms - 24 if sms >= 96 else 0
        # the packed weights are constants of this engine (uploaded at construction, never written by a kernel): the tensor-core
        # kernels fetch them before their programmatic-dependent-launch wait.  DBSR_NO_EARLY_WEIGHTS=1: A/B switch
        self.static_weights = os.environ.get('DBSR_NO_EARLY_WEIGHTS', '0') != '1'
        if os.environ.get('DBSR_ENC_GRID_LIMIT'):          # tuning aid: A/B other splits of the SMs between the two streams
            self.encoder_grid_limit = int(os.environ['DBSR_ENC_GRID_LIMIT'])
        self.launches = 0
        self.layer_events = None   # when a dict (and timers is on): conv layer key -> [(events, flops, family, shape)]
        self.timers = None   # when a dict: family -> list of (start, end) CUDA events on the launching stream
        self.flops = {}      # family -> algorithmic FLOPs (2*MAC, real channel counts) launched since reset
        self.hbm_bytes = {}  # family -> algorithmic HBM bytes (input + output + residual maps, once each) since reset
        sd = {k: _host(v) for k, v in state_dict.items()}      # packing happens on the host, see pack_*
        if 'pwc' in parts:
            self._pack_pwc(sd)
        if 'encoder' in parts or 'merging' in parts or 'decoder' in parts:
            self._pack_dbsr(sd, parts)
        if gauss_kernel is False:
            gauss_kernel = None
        elif gauss_kernel is None:
            k = torch.arange(-1.0, 2.0)
            g = torch.exp(-0.5 * k ** 2) / math.sqrt(2 * math.pi)
            K = g.view(1, -1) * g.view(-1, 1)
            gauss_kernel = K / K.sum()
        self.gauss = [float(v) for v in gauss_kernel.reshape(-1).tolist()] if gauss_kernel is not None else None

    # ------------------------------------------------------------------------------------------------
    # weight packing
    # ------------------------------------------------------------------------------------------------
    def _add(self, key, w, b, tc=False, chmap=None, cin_buf=None, shuffle_r=0, direct_too=True):
        direct = pack_direct(w, chmap, cin_buf, self.device) if direct_too else None
        tcw = pack_tc(w, shuffle_r, chmap, cin_buf, self.device) if tc else None
        self.W[key] = ConvW(direct, tcw, None if b is None else b.float().contiguous().to(self.device), w.shape[2], w.shape[0],
                             w.shape[1], shuffle_r)

    def _pack_pwc(self, sd):
        pre = self.pwc_prefix
        ptc = self.pwc_precision == 'bf16'
        for name in PWC_NAMES:
            for idx in (0, 2, 4):
                k = f'{pre}netExtractor.net{name}.{idx}'
                self._add(k, sd[k + '.weight'], sd[k + '.bias'], tc=ptc and idx != 0)   # idx 0 is the stride-2 conv
                if idx == 0 and ptc:     # tensor-core form of the stride-2 conv: stride 1 over the space-to-depth input
                    self._add(k + '.s2d', pack_s2d_weight(sd[k + '.weight']), sd[k + '.bias'], tc=True)
        self.pwc_layouts = {l: PwcLayout(l) for l in (2, 3, 4, 5, 6)}
        segs = ['V', 'o1', 'o2', 'o3', 'o4', 'o5']   # input of netOne..netSix starts at this segment
        for lvl in (6, 5, 4, 3, 2):
            lay = self.pwc_layouts[lvl]
            lname = PWC_NAMES[lvl - 1]
            for j, sub in enumerate(PWC_NAMES):
                k = f'{pre}net{lname}.net{sub}.0'
                cm, _start, length = lay.chmap_from(segs[j])
                self._add(k, sd[k + '.weight'], sd[k + '.bias'], tc=ptc, chmap=cm, cin_buf=length)
            if lvl < 6:
                prev = self.pwc_layouts[lvl + 1]
                cm, _s, length = prev.chmap_from('o5')
                k = f'{pre}net{lname}.netUpfeat'
                self.D[k] = (pack_deconv(sd[k + '.weight'], cm, length, self.device), sd[k + '.bias'].float().contiguous().to(self.device))
                # the same transposed conv as a 1x1 conv Cin -> 32 planes (ky, kx, oc) + a scatter (ops.deconv_col2im)
                w1 = sd[k + '.weight'].permute(2, 3, 1, 0).reshape(32, -1, 1, 1).contiguous()
                self._add(k + '.taps', w1, None, tc=ptc, chmap=cm, cin_buf=length)
                k = f'{pre}net{lname}.netUpflow'
                self.D[k] = (pack_deconv(sd[k + '.weight'], device=self.device), sd[k + '.bias'].float().contiguous().to(self.device))
        # Tensor-core path: the flow head of level L (netSix: 3x3, ~600 -> 2 channels) is nine N = 16 MMAs per K step that
        # each fetch the whole activation tile from shared memory.  Its channel contraction is run as a 1x1 convolution to 18
        # (tap, oc) planes instead -- one MMA per K step -- in the SAME launch as the 32 planes of the next level's netUpfeat
        # (a transposed conv over the same concat buffer); the taps are summed by deconv_col2im / flow_from_taps.
        #   rows [0, 32): netUpfeat of level L - 1 (levels 6..3), then 18 rows (ky*3 + kx)*2 + oc of netSix
        self.pwc_flow_taps = ptc and os.environ.get('DBSR_NO_FLOW_TAPS', '0') != '1'
        if self.pwc_flow_taps:
            for lvl in (6, 5, 4, 3, 2):
                lname = PWC_NAMES[lvl - 1]
                cm, _s, length = self.pwc_layouts[lvl].chmap_from('o5')
                w6 = _host(sd[f'{pre}net{lname}.netSix.0.weight'])                      # [2, K, 3, 3]
                rows = [w6.permute(2, 3, 0, 1).reshape(18, -1)]
                if lvl > 2:
                    wu = _host(sd[f'{pre}net{PWC_NAMES[lvl - 2]}.netUpfeat.weight'])     # [K, 2, 4, 4]
                    rows.insert(0, wu.permute(2, 3, 1, 0).reshape(32, -1))
                wm = torch.cat(rows, 0)
                self._add(f'{pre}net{lname}.netSix.taps', wm.reshape(wm.shape[0], -1, 1, 1).contiguous(), None, tc=True, chmap=cm,
                          cin_buf=length, direct_too=False)
        cm, _s, length = self.pwc_layouts[2].chmap_from('o5')
        for j in range(7):
            k = f'{pre}netRefiner.netMain.{2 * j}'
            if j == 0:
                self._add(k, sd[k + '.weight'], sd[k + '.bias'], tc=ptc, chmap=cm, cin_buf=length)
            else:
                self._add(k, sd[k + '.weight'], sd[k + '.bias'], tc=ptc)

    def _pack_dbsr(self, sd, parts):
        def add(k, bias=True, shuffle_r=0):
            self._add(k, sd[k + '.weight'], sd.get(k + '.bias') if bias else None, tc=self.bf16, shuffle_r=shuffle_r)

        if 'encoder' in parts:
            add('encoder.init_layer.0')
            self.enc_res = len([k for k in sd if k.startswith('encoder.res_layers.') and k.endswith('conv1.0.weight')])
            for i in range(self.enc_res):
                add(f'encoder.res_layers.{i}.conv1.0')
                add(f'encoder.res_layers.{i}.conv2.0')
            add('encoder.out_layer.0')
            self.enc_dim = sd['encoder.init_layer.0.weight'].shape[0]
            self.feat_dim = sd['encoder.out_layer.0.weight'].shape[0]
        if 'merging' in parts:
            add('merging.feat_project_layer.0')
            add('merging.offset_feat_extractor.0.0')
            self.off_res = len([k for k in sd if k.startswith('merging.offset_feat_extractor.') and k.endswith('conv1.0.weight')])
            for i in range(self.off_res):
                add(f'merging.offset_feat_extractor.{i + 1}.conv1.0')
                add(f'merging.offset_feat_extractor.{i + 1}.conv2.0')
            self.wp_res = len([k for k in sd if k.startswith('merging.weight_predictor.') and k.endswith('conv1.0.weight')])
            add('merging.weight_predictor.0.0')
            for i in range(self.wp_res):
                add(f'merging.weight_predictor.{i + 1}.conv1.0')
                add(f'merging.weight_predictor.{i + 1}.conv2.0')
            add(f'merging.weight_predictor.{self.wp_res + 1}.0')
            self.proj_dim = sd['merging.feat_project_layer.0.weight'].shape[0]
            # The first conv of the weight predictor sees [p_0 | p_n - p_0 | e_n] (merging.py:108-112).  It is linear, so
            #   W [p_0 | p_n - p_0 | e_n] = W_d p_n + W_o e_n + (W_b - W_d) p_0 :
            # the last term depends on the burst only.  On the tensor-core path it is computed ONCE per burst (a 64 -> 128 conv
            # over B maps instead of B * N) and added as a broadcast residual; the per-frame conv contracts 128 instead of 192
            # channels and the warp kernel writes p_n only (no replicated p_0, no difference).
            pd_ = self.proj_dim
            w0 = sd['merging.weight_predictor.0.0.weight']
            self.split_wp0 = bool(self.bf16 and w0.shape[1] == 2 * pd_ + sd['merging.offset_feat_extractor.0.0.weight'].shape[0]
                                  and os.environ.get('DBSR_NO_SPLIT_WP0') is None)
            if self.split_wp0:
                self._add('merging.weight_predictor.0.0.frame', torch.cat([w0[:, pd_:2 * pd_], w0[:, 2 * pd_:]], dim=1).contiguous(),
                          None, tc=True, direct_too=False)
                self._add('merging.weight_predictor.0.0.burst', (w0[:, :pd_] - w0[:, pd_:2 * pd_]).contiguous(),
                          sd['merging.weight_predictor.0.0.bias'], tc=True, direct_too=False)
            self.offf_dim = sd['merging.offset_feat_extractor.0.0.weight'].shape[0]
            self.wp_dim = sd['merging.weight_predictor.0.0.weight'].shape[0]
            self.feat_dim = sd['merging.feat_project_layer.0.weight'].shape[1]
        if 'decoder' in parts:
            add('decoder.init_layer.0')
            self.dec_pre = len([k for k in sd if k.startswith('decoder.pre_res_layers.') and k.endswith('conv1.0.weight')])
            for i in range(self.dec_pre):
                add(f'decoder.pre_res_layers.{i}.conv1.0')
                add(f'decoder.pre_res_layers.{i}.conv2.0')
            up_w = sd['decoder.upsample_layer.conv_layer.0.weight']
            self.dec_dim = sd['decoder.init_layer.0.weight'].shape[0]
            self.post_dim = sd['decoder.predictor.0.weight'].shape[1]
            self.up_r = int(round(math.sqrt(up_w.shape[0] // self.post_dim)))
            self._add('decoder.upsample_layer.conv_layer.0', up_w, sd.get('decoder.upsample_layer.conv_layer.0.bias'),
                      tc=self.bf16 and self.up_r == 8 and self.post_dim == 32, shuffle_r=self.up_r)
            self.dec_post = len([k for k in sd if k.startswith('decoder.post_res_layers.') and k.endswith('conv1.0.weight')])
            for i in range(self.dec_post):
                add(f'decoder.post_res_layers.{i}.conv1.0')
                add(f'decoder.post_res_layers.{i}.conv2.0')
            pw = sd['decoder.predictor.0.weight'].float().reshape(sd['decoder.predictor.0.weight'].shape[0], -1).contiguous()
            pb = sd['decoder.predictor.0.bias'].float().contiguous()
            # host copies for the fused predictor epilogue (its weights travel in the kernel parameters)
            self.pred_w_host = (ctypes.c_float * pw.numel())(*pw.reshape(-1).tolist())
            self.pred_b_host = (ctypes.c_float * pb.numel())(*pb.tolist())
            self.pred_w, self.pred_b = pw.to(self.device), pb.to(self.device)
            self.feat_dim = sd['decoder.init_layer.0.weight'].shape[1]

    # ------------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------------
    def _conv(self, key: str, x: Act, y: Act, act: int, stride: int = 1, dilation: int = 1,
              residual: Optional[Act] = None, force_direct: bool = False, no_bias: bool = False,
              real_cin: Optional[int] = None, pred: Optional[torch.Tensor] = None, grid_limit: int = 0,
              residual_group: int = 0) -> Act:
        cw = self.W[key]
        cin_alg = cw.cin if real_cin is None else real_cin       # algorithmic input channels for the FLOP count
        bias, bias_tc = (None, None) if no_bias else (cw.bias, cw.bias_tc)
        use_tc = (cw.tc is not None and not force_direct and x.dtype == torch.bfloat16 and stride == 1)
        if use_tc:
            use_tc = ops.conv2d_tc_supported(x, cw.tc, bias_tc, y, cw.ksize, stride, dilation, residual, cw.shuffle_r,
                                             residual_group)
        assert use_tc or residual_group <= 1, 'a broadcast residual exists on the tensor-core path only'
        self.launches += 1
        fam = 'conv_tc' if use_tc else 'conv_direct'
        ho, wo = (y.h, y.w) if cw.shuffle_r <= 1 else (y.h // cw.shuffle_r, y.w // cw.shuffle_r)
        self.flops[fam] = self.flops.get(fam, 0) + 2 * x.n * ho * wo * cw.cout * cin_alg * cw.ksize * cw.ksize
        es_in, es_out = x.buf.element_size(), y.buf.element_size()
        nbytes = x.n * x.h * x.w * cin_alg * es_in + y.n * y.h * y.w * y.c * es_out
        if residual is not None:     # (a broadcast residual is still read once per output tile: count it per output image)
            nbytes += y.n * residual.h * residual.w * residual.c * residual.buf.element_size()
        if pred is not None:     # the output map is replaced by the fp32 prediction
            nbytes += pred.numel() * 4 - y.n * y.h * y.w * y.c * es_out
        self.hbm_bytes[fam] = self.hbm_bytes.get(fam, 0) + nbytes
        ev = self._tic(fam)
        if ev is not None and self.layer_events is not None:
            fl = 2 * x.n * ho * wo * cw.cout * cin_alg * cw.ksize * cw.ksize
            self.layer_events.setdefault(key, []).append((self.timers[fam][-1], fl, fam, (x.n, x.h, x.w, cw.cin, cw.cout)))
        if pred is not None:
            assert use_tc, 'the fused predictor epilogue exists on the tensor-core path only'
            self.flops[fam] += 2 * x.n * ho * wo * self.pred_w.shape[0] * cw.cout
            ops.conv2d_tc_predictor(x, cw.tc, bias_tc, y, cw.ksize, act, residual, self.pred_w_host, self.pred_b_host, pred,
                                    grid_limit=grid_limit, static_weights=self.static_weights)
        elif use_tc:
            ops.conv2d(x, cw.tc, bias_tc, y, cw.ksize, stride, dilation, act, residual, cw.shuffle_r, tensor_core=True,
                       grid_limit=grid_limit, residual_group=residual_group, static_weights=self.static_weights)
        else:
            ops.conv2d(x, cw.direct, bias, y, cw.ksize, stride, dilation, act, residual, cw.shuffle_r)
        self._toc(ev)
        return y

    def _run(self, family: str, fn, *args, **kw):
        """launch one kernel through ops.*, counted and (optionally) timed under `family`"""
        ev = self._tic(family)
        r = fn(*args, **kw)
        self._toc(ev)
        self.launches += 1
        return r

    def _tic(self, family: str):
        if self.timers is None:
            return None
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        self.timers.setdefault(family, []).append((a, b))
        return b

    @staticmethod
    def _toc(ev):
        if ev is not None:
            ev.record()

    def layer_summary(self) -> list:
        """per conv layer: (key, family, shape, launches, total ms, TFLOP/s), sorted by time; call after a synchronize"""
        rows = []
        for key, evs in (self.layer_events or {}).items():
            ms = sum(a.elapsed_time(b) for (a, b), _f, _fam, _s in evs)
            fl = sum(f for _e, f, _fam, _s in evs)
            rows.append((key, evs[0][2], evs[0][3], len(evs), ms, fl / (ms / 1e3) / 1e12 if ms > 0 else 0.0))
        return sorted(rows, key=lambda r: -r[4])

    def timer_summary(self) -> dict:
        """family -> (total ms, launches); call after a synchronize."""
        out = {}
        for fam, evs in (self.timers or {}).items():
            out[fam] = (sum(a.elapsed_time(b) for a, b in evs), len(evs))
        return out

    def _resblock(self, key: str, x: Act, tmp: Act, y: Act, grid_limit: int = 0) -> Act:
        """reference models/layers/blocks.py:84-96: relu(x + conv2(relu(conv1(x))))"""
        self._conv(key + '.conv1.0', x, tmp, ACT_RELU, grid_limit=grid_limit)
        return self._conv(key + '.conv2.0', tmp, y, ACT_RELU, residual=x, grid_limit=grid_limit)

    def _buf(self, ws: dict, name: str, n, h, w, c, dtype, zero=False) -> Act:
        """workspace buffer; the pixel pitch is rounded up to 8 channels (16-byte rows for TMA / vector access)"""
        a = ws.get(name)
        if a is None:
            pitch = _align8(c)
            a = Act.empty(n, h, w, pitch, dtype, self.device, zero=zero or pitch != c).slice(0, c)
            ws[name] = a
        return a

    def workspace(self, key: tuple) -> dict:
        ws = self._ws.get(key)
        if ws is None:
            ws = {}
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------------
    # PWC-Net  (reference models/alignment/pwcnet.py)
    # ------------------------------------------------------------------------------------------------
    @on_engine_device
    def pwc_extract(self, ws: dict, pwc_in: Optional[Act], s2d0: Optional[Act] = None) -> list:
        """Extractor pyramid (pwcnet.py:45-111) on every image of `pwc_in` [n, Hp, Wp, >=3] -- or, on the bf16 path, of
        `s2d0` [n, Hp/2, Wp/2, 12]: the same images already in the space-to-depth layout of the first stride-2 conv
        (written by ops.prep_burst_s2d)."""
        pre = self.pwc_prefix
        if s2d0 is not None:
            n, h, w = s2d0.n, 2 * s2d0.h, 2 * s2d0.w
            x = None
        else:
            n = pwc_in.n
            x = pwc_in.slice(0, 3)
            h, w = pwc_in.h, pwc_in.w
        feats = []
        for l, name in enumerate(PWC_NAMES):
            c = PWC_EXT_CH[l + 1]
            h, w = (h + 1) // 2, (w + 1) // 2
            t1 = self._buf(ws, f'ext{l}_a', n, h, w, c, self.pwc_dtype)
            t2 = self._buf(ws, f'ext{l}_b', n, h, w, c, self.pwc_dtype)
            f = self._buf(ws, f'ext{l}_f', n, h, w, c, self.pwc_dtype)
            k0 = f'{pre}netExtractor.net{name}.0'
            if l == 0 and s2d0 is not None:
                self._conv(k0 + '.s2d', s2d0, t1, ACT_LRELU, real_cin=3)
            elif (k0 + '.s2d') in self.W:
                xs = self._buf(ws, f'ext{l}_s2d', n, h, w, 4 * x.c, torch.bfloat16)
                self._run('copy', ops.space_to_depth2, x, xs)
                self._conv(k0 + '.s2d', xs, t1, ACT_LRELU, real_cin=x.c)
            else:
                self._conv(k0, x, t1, ACT_LRELU, stride=2)
            self._conv(f'{pre}netExtractor.net{name}.2', t1, t2, ACT_LRELU)
            self._conv(f'{pre}netExtractor.net{name}.4', t2, f, ACT_LRELU)
            feats.append(f)
            x = f
        return feats

    @on_engine_device
    def pwc_decode(self, ws: dict, first: list, second: list, pairs: int, group: int, src_group: int) -> Act:
        """Decoders 6..2 + refiner (pwcnet.py:113-231).  `first`/`second`: per-level feature Acts.  group > 0: burst
        mode (both lists are the same per-frame pyramid, frame 0 of each burst is the reference)."""
        pre = self.pwc_prefix
        segs = ['V', 'o1', 'o2', 'o3', 'o4', 'o5']
        prev_cat = None
        prev_flow = None
        prev_taps = None          # tensor-core path: [pairs, h, w, 64] fp32, [0, 32) netUpfeat planes, [32, 50) flow-head planes
        prev_b6 = None
        ftaps = getattr(self, 'pwc_flow_taps', False)
        for lvl in (6, 5, 4, 3, 2):
            lay = self.pwc_layouts[lvl]
            lname = PWC_NAMES[lvl - 1]
            f1, f2 = first[lvl - 1], second[lvl - 1]
            h, w = f1.h, f1.w
            cat = self._buf(ws, f'cat{lvl}', pairs, h, w, lay.total, self.pwc_dtype, zero=True)
            flow = None if (ftaps and lvl > 2) else self._buf(ws, f'flow{lvl}', pairs, h, w, 2, torch.float32)
            vol = cat.slice(lay.off['V'], 81)
            if prev_cat is None:
                self._run('corr81', ops.corr81, f1, f2, vol, pairs, group, act=ACT_LRELU)
            else:
                upflow = self._buf(ws, f'upflow{lvl}', pairs, h, w, 2, torch.float32)
                wf, bf = self.D[f'{pre}net{lname}.netUpflow']
                _wt, bt = self.D[f'{pre}net{lname}.netUpfeat']
                if ftaps:
                    self._run('deconv', ops.deconv_col2im, prev_taps.slice(0, 32), bt, cat.slice(lay.off['upfeat'], 2), None, wf, bf,
                              cat.slice(lay.off['upflow'], 2), upflow, flow_taps=prev_taps.slice(32, 18), flow_bias=prev_b6)
                else:
                    taps = self._buf(ws, f'taps{lvl}', pairs, prev_cat.h, prev_cat.w, 32, torch.float32)
                    self._conv(f'{pre}net{lname}.netUpfeat.taps', prev_cat, taps, ACT_NONE)
                    self._run('deconv', ops.deconv_col2im, taps, bt, cat.slice(lay.off['upfeat'], 2), prev_flow, wf, bf,
                              cat.slice(lay.off['upflow'], 2), upflow)
                # the first map's slice of the concat buffer is written by the cost-volume launch (it stages that tile anyway)
                assert group == 0 or src_group == group + 1
                self._run('corr81', ops.corr81, f1, f2, vol, pairs, group, flow=upflow, flow_scale=PWC_BACKWARP[lvl],
                          act=ACT_LRELU, f1_copy=cat.slice(lay.off['f1'], lay.sizes['f1']))
            for j, sub in enumerate(PWC_NAMES[:5]):
                _cm, start, length = lay.chmap_from(segs[j])
                out_name = 'o%d' % (j + 1)
                self._conv(f'{pre}net{lname}.net{sub}.0', cat.slice(start, length),
                           cat.slice(lay.off[out_name], lay.sizes[out_name]), ACT_LRELU)
            if ftaps:
                mt = self._buf(ws, f'mtaps{lvl}', pairs, h, w, 64, torch.float32)
                b6 = self.W[f'{pre}net{lname}.netSix.0'].bias
                if lvl > 2:
                    self._conv(f'{pre}net{lname}.netSix.taps', cat, mt.slice(0, 50), ACT_NONE)
                    prev_taps, prev_b6 = mt, b6
                else:       # the finest level's flow is a map: the refiner adds its output to it (pwcnet.py:231)
                    self._conv(f'{pre}net{lname}.netSix.taps', cat, mt.slice(0, 18), ACT_NONE)
                    self._run('deconv', ops.flow_from_taps, mt.slice(0, 18), b6, flow)
            else:
                self._conv(f'{pre}net{lname}.netSix.0', cat, flow, ACT_NONE)
            prev_cat, prev_flow = cat, flow
        # refiner (dilated convs), result added to the level-2 flow (pwcnet.py:231)
        h, w = prev_cat.h, prev_cat.w
        chans = [128, 128, 128, 96, 64, 32]
        x = prev_cat
        for j in range(6):
            y = self._buf(ws, f'ref{j}', pairs, h, w, chans[j], self.pwc_dtype)
            self._conv(f'{pre}netRefiner.netMain.{2 * j}', x, y, ACT_LRELU, dilation=PWC_REFINER_DIL[j])
            x = y
        flow4 = self._buf(ws, 'flow_quarter', pairs, h, w, 2, torch.float32)
        self._conv(f'{pre}netRefiner.netMain.12', x, flow4, ACT_NONE, dilation=1, residual=prev_flow)
        return flow4

    @on_engine_device
    def pwc_burst(self, ws: dict, pwc_in: Optional[Act], B: int, N: int, H: int, W: int, offsets: torch.Tensor,
                  s2d0: Optional[Act] = None) -> torch.Tensor:
        """PWCNet.forward (pwcnet.py:248-281) for a burst batch: frame 0 of every burst is the target; the pyramid of
        each frame is computed once (the reference recomputes the reference frame's pyramid N-1 times)."""
        feats = self.pwc_extract(ws, pwc_in, s2d0)
        flow4 = self.pwc_decode(ws, feats, feats, B * (N - 1), N - 1, N)
        Hp, Wp = (2 * s2d0.h, 2 * s2d0.w) if s2d0 is not None else (pwc_in.h, pwc_in.w)
        self._run('flow_head', ops.flow_head, flow4, offsets, H, W, Hp, Wp)
        return offsets

    @on_engine_device
    def prep_and_align(self, ws: dict, burst: torch.Tensor, enc_in: Act, offsets: torch.Tensor) -> torch.Tensor:
        """burst [B, N, 4, H, W] fp32 -> `enc_in` (channels-last packed RAW for the encoder) and the flows `offsets`
        [B*(N-1), 2, H, W] of every frame towards frame 0 (encoders.py:52-61 + PWCNet.forward)."""
        B, N, _, H, W = burst.shape
        s2d0, pwc_in = self.prep(ws, burst, enc_in)
        return self.pwc_burst(ws, pwc_in, B, N, H, W, offsets, s2d0)

    @on_engine_device
    def prep(self, ws: dict, burst: torch.Tensor, enc_in: Act):
        """burst -> `enc_in` + the PWC-Net input: (s2d0, None) on the bf16 path (RGGB->RGB + resize written straight into
        the space-to-depth layout of the extractor's first stride-2 conv), (None, pwc_in) on the fp32 path."""
        B, N, _, H, W = burst.shape
        Hp, Wp = int(math.ceil(H / 64.0) * 64), int(math.ceil(W / 64.0) * 64)
        F_ = B * N
        if (self.pwc_prefix + 'netExtractor.netOne.0.s2d') in self.W:
            s2d0 = self._buf(ws, 'ext0_s2d', F_, Hp // 2, Wp // 2, 12, torch.bfloat16)
            self._run('prep_burst', ops.prep_burst_s2d, burst, enc_in, s2d0, Hp, Wp)
            return s2d0, None
        pwc_in = self._buf(ws, 'pwc_in', F_, Hp, Wp, 4, torch.float32)
        self._run('prep_burst', ops.prep_burst, burst, enc_in, pwc_in)
        return None, pwc_in

    # ------------------------------------------------------------------------------------------------
    # DBSR stages
    # ------------------------------------------------------------------------------------------------
    @on_engine_device
    def encode(self, ws: dict, enc_in: Act, grid_limit: int = 0) -> Act:
        """conv stack of ResEncoderWarpAlignnet (reference models/dbsr/encoders.py:66-72) on all B*N frames.
        grid_limit: persistent-grid cap of these launches (per call; > 0 while PWC-Net shares the GPU on a second stream)."""
        F_, H, W = enc_in.n, enc_in.h, enc_in.w
        dt = self.act_dtype
        xa = self._buf(ws, 'enc_a', F_, H, W, self.enc_dim, dt)
        xb = self._buf(ws, 'enc_b', F_, H, W, self.enc_dim, dt)
        xt = self._buf(ws, 'enc_t', F_, H, W, self.enc_dim, dt)
        feat = self._buf(ws, 'feat', F_, H, W, self.feat_dim, dt)
        self._conv('encoder.init_layer.0', enc_in.slice(0, 4), xa, ACT_RELU, grid_limit=grid_limit)
        cur, nxt = xa, xb
        for i in range(self.enc_res):
            self._resblock(f'encoder.res_layers.{i}', cur, xt, nxt, grid_limit=grid_limit)
            cur, nxt = nxt, cur
        self._conv('encoder.out_layer.0', cur, feat, ACT_RELU, grid_limit=grid_limit)
        return feat

    @on_engine_device
    def project(self, ws: dict, feat: Act, grid_limit: int = 0) -> Act:
        """q = W_p . feat (merging.py:75 without bias / ReLU, which follow the warp in `warp_proj`): depends on the
        embeddings only, so the forward runs it before joining the alignment stream"""
        q = self._buf(ws, 'proj_q', feat.n, feat.h, feat.w, self.proj_dim, self.act_dtype)
        return self._conv('merging.feat_project_layer.0', feat, q, ACT_NONE, no_bias=True, grid_limit=grid_limit)

    @on_engine_device
    def merge(self, ws: dict, feat: Act, offsets: torch.Tensor, B: int, N: int,
              weights_out: Optional[torch.Tensor] = None, aligned: bool = False, projected: bool = False) -> Act:
        """WeightedSum.forward (reference models/dbsr/merging.py:61-127) with the warp of encoders.py:80 folded in.
        `feat` [B*N, H, W, C]: the frame embeddings -- unwarped (aligned=False: the kernels gather through `offsets`
        on the fly, the warped 512-channel tensor is never materialised) or already aligned (aligned=True, the
        WeightedSum module seam).  offsets [B*(N-1), 2, H, W]."""
        F_, H, W = feat.n, feat.h, feat.w
        dt = self.act_dtype
        pd, od, wd = self.proj_dim, self.offf_dim, self.wp_dim
        gather = None if aligned else offsets
        # 1x1 projection commutes with the bilinear warp: project first (tensor cores), warp 64 channels instead of 512
        q = self._buf(ws, 'proj_q', F_, H, W, pd, dt)
        if not projected:
            self.project(ws, feat)
        split = getattr(self, 'split_wp0', False)
        if split:
            wp_in = self._buf(ws, 'wp_in_split', F_, H, W, pd + od, dt)        # [p_n | e_n]
            p0 = self._buf(ws, 'proj_p0', B, H, W, pd, dt)
            self._run('warp_proj', ops.warp_proj_split, q, self.W['merging.feat_project_layer.0'].bias, wp_in, p0, N, gather)
            e_off = pd
        else:
            wp_in = self._buf(ws, 'wp_in', F_, H, W, 2 * pd + od, dt)          # [p_0 | p_n - p_0 | e_n]
            self._run('warp_proj', ops.warp_proj, q, self.W['merging.feat_project_layer.0'].bias, wp_in, N, gather)
            e_off = 2 * pd
        offm = self._buf(ws, 'offm', F_, H, W, 8, dt)
        self._run('offsets_mod', ops.offsets_mod, offsets, offm, B, N, self.offset_modulo)
        oa = self._buf(ws, 'off_a', F_, H, W, od, dt)
        ob = self._buf(ws, 'off_b', F_, H, W, od, dt)
        ot = self._buf(ws, 'off_t', F_, H, W, od, dt)
        self._conv('merging.offset_feat_extractor.0.0', offm.slice(0, 2), oa, ACT_RELU)
        cur, nxt = oa, ob
        for i in range(self.off_res):
            last = i == self.off_res - 1
            dst = wp_in.slice(e_off, od) if last else nxt
            self._resblock(f'merging.offset_feat_extractor.{i + 1}', cur, ot, dst)
            cur, nxt = dst, cur
        if self.off_res == 0:
            self._run('copy', ops.copy_channels, oa, wp_in.slice(e_off, od))
        wa = self._buf(ws, 'wp_a', F_, H, W, wd, dt)
        wb = self._buf(ws, 'wp_b', F_, H, W, wd, dt)
        wt = self._buf(ws, 'wp_t', F_, H, W, wd, dt)
        if split:
            base = self._buf(ws, 'wp_base', B, H, W, wd, dt)                   # (W_b - W_d) p_0 + bias, one map per burst
            self._conv('merging.weight_predictor.0.0.burst', p0, base, ACT_NONE)
            self._conv('merging.weight_predictor.0.0.frame', wp_in, wa, ACT_RELU, residual=base, residual_group=N)
        else:
            self._conv('merging.weight_predictor.0.0', wp_in, wa, ACT_RELU)
        cur, nxt = wa, wb
        for i in range(self.wp_res):
            self._resblock(f'merging.weight_predictor.{i + 1}', cur, wt, nxt)
            cur, nxt = nxt, cur
        logits = self._buf(ws, 'logits', F_, H, W, self.feat_dim, self.logits_dtype)
        self._conv(f'merging.weight_predictor.{self.wp_res + 1}.0', cur, logits, ACT_NONE)
        fused = self._buf(ws, 'fused', B, H, W, self.feat_dim, dt)
        self._run('softmax_wsum', ops.softmax_wsum, feat, logits, fused, N, offsets=gather, weights_out=weights_out)
        self.launches += (1 if weights_out is not None else 0)
        return fused

    @on_engine_device
    def decode(self, ws: dict, fused: Act, pred: torch.Tensor) -> torch.Tensor:
        """ResPixShuffleConv.forward (reference models/dbsr/decoders.py:54-62, models/layers/upsampling.py:51-66)."""
        B, H, W = fused.n, fused.h, fused.w
        dt = self.act_dtype
        r = self.up_r
        da = self._buf(ws, 'dec_a', B, H, W, self.dec_dim, dt)
        db = self._buf(ws, 'dec_b', B, H, W, self.dec_dim, dt)
        dtmp = self._buf(ws, 'dec_t', B, H, W, self.dec_dim, dt)
        self._conv('decoder.init_layer.0', fused, da, ACT_RELU)
        cur, nxt = da, db
        for i in range(self.dec_pre):
            self._resblock(f'decoder.pre_res_layers.{i}', cur, dtmp, nxt)
            cur, nxt = nxt, cur
        ha = self._buf(ws, 'hr_a', B, H * r, W * r, self.post_dim, dt)
        hb = self._buf(ws, 'hr_b', B, H * r, W * r, self.post_dim, dt)
        ht = self._buf(ws, 'hr_t', B, H * r, W * r, self.post_dim, dt)
        self._conv('decoder.upsample_layer.conv_layer.0', cur, ha, ACT_RELU)      # 1x1 conv + ReLU + PixelShuffle
        if self.gauss is not None:
            self._run('blur3x3', ops.blur3x3, ha, hb, self.gauss)
            cur, nxt = hb, ha
        else:
            cur, nxt = ha, hb
        for i in range(self.dec_post):
            key = f'decoder.post_res_layers.{i}'
            last = i == self.dec_post - 1
            if self._fused_resblock_ok(key, cur, nxt):
                # relu(x + conv2(relu(conv1(x)))) in ONE launch, the intermediate map stays on the SM; the last block also
                # applies the 1x1 predictor + ReLU (decoders.py:52,61) in its epilogue and writes `pred` directly
                fuse_pred = last and self.pred_w.shape[0] <= 4
                self._resblock_fused(key, cur, nxt, pred if fuse_pred else None)
                if fuse_pred:
                    return pred
                cur, nxt = nxt, cur
                continue
            if last and self._fuse_predictor(key + '.conv2.0', ht, nxt, cur):
                # last block: relu(x + conv2(relu(conv1(x)))) never goes to HBM -- the tcgen05 epilogue applies the 1x1
                # predictor + ReLU (decoders.py:52,61) to the 32 channels each thread holds and writes `pred` directly
                self._conv(key + '.conv1.0', cur, ht, ACT_RELU)
                self._conv(key + '.conv2.0', ht, nxt, ACT_RELU, residual=cur, pred=pred)
                return pred
            self._resblock(key, cur, ht, nxt)
            cur, nxt = nxt, cur
        if pred.dtype == torch.int16:       # unfused path (fp32 precision): float prediction, then the 14-bit quantisation
            pf = ws.get('pred_f32')
            if pf is None or pf.shape != pred.shape:
                pf = ws['pred_f32'] = torch.empty(pred.shape, dtype=torch.float32, device=self.device)
            self._run('predictor', ops.predictor, cur, self.pred_w, self.pred_b, pf)
            self._run('predictor', ops.quantize_q14, pf, pred)
            return pred
        self._run('predictor', ops.predictor, cur, self.pred_w, self.pred_b, pred)
        return pred

    def _fused_resblock_ok(self, key: str, x: Act, y: Act) -> bool:
        c1, c2 = self.W.get(key + '.conv1.0'), self.W.get(key + '.conv2.0')
        if not (self.fuse_hr_resblocks and c1 is not None and c2 is not None and c1.tc is not None and c2.tc is not None):
            return False
        if not (x.dtype == torch.bfloat16 and c1.cin == 32 and c1.cout == 32 and c2.cin == 32 and c2.cout == 32 and
                c1.ksize == 3 and c2.ksize == 3 and c1.bias is not None and c2.bias is not None):
            return False
        return ops.resblock32_tc_supported(x, y, c1.tc, c1.bias, c2.tc, c2.bias)

    def _resblock_fused(self, key: str, x: Act, y: Act, pred: Optional[torch.Tensor]) -> None:
        c1, c2 = self.W[key + '.conv1.0'], self.W[key + '.conv2.0']
        fam = 'resblock_tc'
        fl = 2 * 2 * x.n * x.h * x.w * 32 * 32 * 9
        nbytes = 2 * x.n * x.h * x.w * 32 * 2
        if pred is not None:
            fl += 2 * x.n * x.h * x.w * self.pred_w.shape[0] * 32
            nbytes += pred.numel() * pred.element_size() - x.n * x.h * x.w * 32 * 2
        self.flops[fam] = self.flops.get(fam, 0) + fl
        self.hbm_bytes[fam] = self.hbm_bytes.get(fam, 0) + nbytes
        self.launches += 1
        ev = self._tic(fam)
        if ev is not None and self.layer_events is not None:
            self.layer_events.setdefault(key, []).append((self.timers[fam][-1], fl, fam, (x.n, x.h, x.w, 32, 32)))
        if pred is not None:
            ops.resblock32_tc(x, None, c1.tc, c1.bias, c2.tc, c2.bias, self.pred_w_host, self.pred_b_host, pred, static_weights=self.static_weights)
        else:
            ops.resblock32_tc(x, y, c1.tc, c1.bias, c2.tc, c2.bias, static_weights=self.static_weights)
        self._toc(ev)

    def _fuse_predictor(self, key: str, x: Act, y: Act, residual: Act) -> bool:
        cw = self.W[key]
        return (cw.tc is not None and x.dtype == torch.bfloat16 and cw.cout == 32 and self.pred_w.shape[0] <= 4 and
                x.w > 8 and ops.conv2d_tc_supported(x, cw.tc, cw.bias_tc, y, cw.ksize, 1, 1, residual, 0))

    # ------------------------------------------------------------------------------------------------
    # whole forward
    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    @on_engine_device
    def forward(self, burst: torch.Tensor, return_weights: bool = False, out: Optional[dict] = None, quantize: bool = False):
        """DBSRNet.forward: burst [B, N, 4, H, W] fp32 CUDA -> pred [B, 3, 8H, 8W], offsets [B, N-1, 2, H, W],
        fusion_weights [B, N, C, H, W] (only when return_weights).  `out` (optional): preallocated 'pred' / 'offsets' to
        write into, and 'offsets_in': flows to use INSTEAD of running PWC-Net."""
        assert burst.dim() == 5 and burst.shape[2] == 4, 'burst must be [B, N, 4, H, W]'
        ops.require_device(burst)
        burst = burst.contiguous().float()
        B, N, _, H, W = burst.shape
        assert N >= 2, 'a burst needs at least 2 frames'
        Hp, Wp = int(math.ceil(H / 64.0) * 64), int(math.ceil(W / 64.0) * 64)
        ws = self.workspace((B, N, H, W))
        F_ = B * N
        enc_in = self._buf(ws, 'enc_in', F_, H, W, 8, self.act_dtype)
        if out is None:
            out = {}
        offsets = out.get('offsets')
        if offsets is None:
            offsets = torch.empty((B * (N - 1), 2, H, W), dtype=torch.float32, device=self.device)
        given = out.get('offsets_in')      # precomputed flows [B, N-1, 2, H, W]: the alignment network is skipped
        if given is not None:
            ops.require_device(given)
            assert given.numel() == offsets.numel(), 'offsets_in must be [B, N-1, 2, H, W]'
            offsets.copy_(given.reshape(offsets.shape).float())
            self.prep(ws, burst, enc_in)
            feat = self.encode(ws, enc_in)
            projected = False
        elif self.overlap_alignment and self.timers is None:
            # PWC-Net (many short launches, most of them on 50-75 of the 148 SMs) and the encoder conv stack (persistent
            # full-grid launches) are independent until the fusion: run them on two streams so that encoder CTAs fill
            # the SMs the small alignment kernels leave idle.  Inside a CUDA-graph capture this becomes a fork / join.
            s2d0, pwc_in = self.prep(ws, burst, enc_in)
            cur = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.pwc_burst(ws, pwc_in, B, N, H, W, offsets, s2d0)
            feat = self.encode(ws, enc_in, grid_limit=self.encoder_grid_limit)
            self.project(ws, feat, grid_limit=self.encoder_grid_limit)           # needs the embeddings only: before the join
            cur.wait_stream(self._side)
            projected = True
        else:
            self.prep_and_align(ws, burst, enc_in, offsets)
            feat = self.encode(ws, enc_in)
            projected = False
        weights = None
        if return_weights:
            weights = torch.empty((B, N, self.feat_dim, H, W), dtype=torch.float32, device=self.device)
        fused = self.merge(ws, feat, offsets, B, N, weights, aligned=False, projected=projected)
        pred = out.get('pred')
        if pred is None:
            # quantize: int16 = (pred.clamp(0, 1) * 2^14).short(), what the reference's evaluation / result writers store
            pred = torch.empty((B, 3, H * self.up_r, W * self.up_r), dtype=torch.int16 if quantize else torch.float32,
                               device=self.device)
        self.decode(ws, fused, pred)
        return pred, offsets.view(B, N - 1, 2, H, W), weights

    # ------------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole forward (launch-bound at small batch: ~400 launches per forward)
    # ------------------------------------------------------------------------------------------------
    @on_engine_device
    def graph_entry(self, shape, return_weights: bool = False, quantize: bool = False, slot: int = 0):
        """(graph, static_in, outs, launches) of the CUDA graph that runs forward() on `static_in` [B, N, 4, H, W] and leaves
        (pred, offsets, weights) in `outs`; captured on first use.  `slot` selects one of several independent graphs of the
        same shape, each with its own static input / output buffers (workspaces are shared: graphs of one engine must be
        replayed on one stream) -- HostPipeline double-buffers with it so that host copies go straight into / out of the
        graph's buffers."""
        key = (tuple(shape), bool(return_weights), bool(quantize), int(slot))
        entry = self._graphs.get(key)
        if entry is None:
            assert self.timers is None, 'per-kernel timers cannot be recorded inside a graph capture'
            static_in = torch.zeros(tuple(shape), dtype=torch.float32, device=self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up outside capture: workspaces, smem attributes, identity tiles
                self.forward(static_in, return_weights, quantize=quantize)
            torch.cuda.current_stream().wait_stream(side)
            launches0 = self.launches
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = self.forward(static_in, return_weights, quantize=quantize)
            entry = (graph, static_in, outs, self.launches - launches0)
            self._graphs[key] = entry
        return entry

    @torch.no_grad()
    @on_engine_device
    def forward_graphed(self, burst: torch.Tensor, return_weights: bool = False, quantize: bool = False, slot: int = 0):
        """Same as forward(), but the launch sequence is captured once per input shape into a CUDA graph and replayed.
        The returned tensors are the graph's static outputs: they are overwritten by the next call of the same shape (and
        slot).  A `burst` that already IS the slot's static input (HostPipeline copies host data straight into it) is not
        copied again."""
        ops.require_device(burst)
        burst = burst.contiguous().float()
        graph, static_in, outs, n_launch = self.graph_entry(burst.shape, return_weights, quantize, slot)
        if burst.data_ptr() != static_in.data_ptr():
            static_in.copy_(burst, non_blocking=True)
        graph.replay()
        self.launches += n_launch
        return outs
