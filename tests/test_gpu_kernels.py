"""GPU: per-kernel parity of the C-ABI entry points against the CPU oracle (oracle/dbsr_oracle.py) on seeded
inputs.  fp32 kernels: tolerance is fp32 round-off of a re-associated sum (stated per test)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import dbsr_oracle as O  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from deep_rawburst_sr_b200 import ops
    ops.require_device(torch.empty(1, device='cuda:0'))
    return torch.device('cuda:0')


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _act_from(x_nchw, dev, dtype=torch.float32, pitch=None, c_off=0):
    from deep_rawburst_sr_b200.ops import Act
    n, c, h, w = x_nchw.shape
    pitch = pitch or c
    buf = Act(torch.zeros((n, h, w, pitch), dtype=dtype, device=dev))
    v = buf.slice(c_off, c)
    v.from_nchw(x_nchw.to(dev).contiguous())
    return v


def test_copy_channels_vector_and_generic(dev):
    """copy of a feature map into a channel slice of a wider buffer with the burst pair->image mapping (reference frame of
    burst b replicated to its N-1 pairs): 16-byte vector path (bf16, 8-aligned) and the generic scalar path"""
    from deep_rawburst_sr_b200 import ops
    B, N, C, H, W = 3, 4, 32, 6, 5
    src = torch.randn(B * N, C, H, W, generator=_gen(21)).bfloat16().float()
    want = src.view(B, N, C, H, W)[:, :1].expand(-1, N - 1, -1, -1, -1).reshape(-1, C, H, W)
    for c_off, c in ((16, C), (13, C - 3)):          # aligned -> vector kernel; odd offset / count -> generic kernel
        dst = ops.Act(torch.full((B * (N - 1), H, W, 64), 2.0, dtype=torch.bfloat16, device=dev))
        ops.copy_channels(_act_from(src[:, :c].contiguous(), dev, dtype=torch.bfloat16, pitch=40), dst.slice(c_off, c), N - 1, N, 0)
        assert torch.equal(dst.slice(c_off, c).to_nchw().cpu(), want[:, :c])
        assert (dst.buf[..., :c_off] == 2.0).all() and (dst.buf[..., c_off + c:] == 2.0).all()


def test_layout_roundtrip(dev):
    x = torch.randn(3, 37, 9, 11, generator=_gen(0))
    for dtype, tol in ((torch.float32, 0.0), (torch.bfloat16, 2e-2)):
        v = _act_from(x, dev, dtype=dtype, pitch=48, c_off=8)
        back = v.to_nchw().cpu()
        assert (back - x).abs().max() <= tol
        assert float(v.buf[..., :8].float().abs().max()) == 0.0 and float(v.buf[..., 45:].float().abs().max()) == 0.0


CONV_CASES = [
    # cin, cout, k, stride, dil, n, h, w, act, residual
    (3, 16, 3, 2, 1, 2, 64, 64, 2, False),
    (16, 16, 3, 1, 1, 2, 32, 32, 2, False),
    (81, 128, 3, 1, 1, 5, 2, 2, 2, False),
    (117, 128, 3, 1, 1, 2, 16, 16, 2, False),
    (565, 2, 3, 1, 1, 2, 16, 16, 0, True),
    (128, 128, 3, 1, 2, 2, 16, 16, 2, False),
    (96, 64, 3, 1, 16, 1, 16, 16, 2, False),
    (64, 64, 3, 1, 1, 2, 24, 40, 1, True),
    (512, 64, 1, 1, 1, 2, 12, 12, 1, False),
    (4, 64, 3, 1, 1, 3, 16, 16, 1, False),
    (196, 196, 3, 1, 1, 3, 1, 1, 2, False),
    (32, 3, 1, 1, 1, 1, 16, 16, 1, False),
]


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_conv2d_direct(dev, case, dtype):
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_direct
    cin, cout, k, stride, dil, n, h, w, act, use_res = case
    g = _gen(hash(case) % 1000)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    pad = dil * (k - 1) // 2
    ref = F.conv2d(x, wt, b, stride=stride, padding=pad, dilation=dil)
    res = torch.randn(ref.shape, generator=g) if use_res else None
    if res is not None:
        if dtype == torch.bfloat16:
            res = res.bfloat16().float()
        ref = ref + res
    ref = torch.relu(ref) if act == 1 else (O.lrelu(ref) if act == 2 else ref)
    xa = _act_from(x, dev, dtype=dtype, pitch=cin + 5, c_off=3)      # unaligned slice on purpose
    ya = ops.Act(torch.zeros((n, ref.shape[2], ref.shape[3], cout + 8), dtype=torch.float32, device=dev)).slice(8, cout)
    ra = _act_from(res, dev, dtype=dtype) if res is not None else None
    ops.conv2d(xa, pack_direct(wt.to(dev)), b.to(dev), ya, k, stride, dil, act, ra)
    got = ya.to_nchw().cpu()
    assert (got - ref).abs().max() < 2e-4 * max(1.0, float(ref.abs().max()))
    assert float(ya.buf[..., :8].abs().max()) == 0.0   # neighbours of the slice untouched


def test_conv2d_direct_pixel_shuffle(dev):
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_direct
    g = _gen(5)
    x = torch.randn(2, 64, 6, 5, generator=g)
    wt = torch.randn(2048, 64, 1, 1, generator=g) / 8
    ref = O.pixel_shuffle(torch.relu(F.conv2d(x, wt)), 8)
    xa = _act_from(x, dev)
    ya = ops.Act.empty(2, 48, 40, 32, torch.float32, dev)
    ops.conv2d(xa, pack_direct(wt.to(dev)), None, ya, 1, 1, 1, ops.ACT_RELU, None, shuffle_r=8)
    assert (ya.to_nchw().cpu() - ref).abs().max() < 1e-5


@pytest.mark.parametrize('cin,h,w', [(2, 4, 4), (529, 1, 1), (661, 2, 3), (597, 8, 8)])
def test_deconv4x4s2(dev, cin, h, w):
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_deconv
    g = _gen(cin)
    x = torch.randn(3, cin, h, w, generator=g)
    wt = torch.randn(cin, 2, 4, 4, generator=g) / (cin * 4) ** 0.5
    b = torch.randn(2, generator=g)
    ref = F.conv_transpose2d(x, wt, b, stride=2, padding=1)
    assert torch.allclose(O.deconv4x4s2(x, wt, b), ref, atol=1e-5)
    xa = _act_from(x, dev)
    ya = ops.Act.empty(3, 2 * h, 2 * w, 16, torch.float32, dev, zero=True).slice(8, 2)
    y2 = ops.Act.empty(3, 2 * h, 2 * w, 2, torch.float32, dev)
    ops.deconv4x4s2(xa, pack_deconv(wt.to(dev)), b.to(dev), ya, y2)
    assert (ya.to_nchw().cpu() - ref).abs().max() < 1e-4
    assert (y2.to_nchw().cpu() - ref).abs().max() < 1e-4


@pytest.mark.parametrize('cin,h,w', [(529, 1, 1), (661, 2, 3), (597, 8, 8), (565, 16, 16)])
def test_deconv_as_conv1x1_plus_col2im(dev, cin, h, w):
    """netUpfeat as a 1x1 conv to 32 (ky, kx, oc) planes + dbsr_deconv_col2im, with netUpflow fused into the scatter"""
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_deconv, pack_direct
    g = _gen(cin + 7)
    x = torch.randn(3, cin, h, w, generator=g)
    wt = torch.randn(cin, 2, 4, 4, generator=g) / (cin * 4) ** 0.5
    bt = torch.randn(2, generator=g)
    fl = torch.randn(3, 2, h, w, generator=g) * 3
    wf = torch.randn(2, 2, 4, 4, generator=g) / 3
    bf = torch.randn(2, generator=g)
    ref_t = O.deconv4x4s2(x, wt, bt)
    ref_f = O.deconv4x4s2(fl, wf, bf)
    w1 = wt.permute(2, 3, 1, 0).reshape(32, cin, 1, 1).contiguous()
    taps = ops.Act.empty(3, h, w, 32, torch.float32, dev)
    ops.conv2d(_act_from(x, dev), pack_direct(w1.to(dev)), None, taps, 1, 1, 1, ops.ACT_NONE)
    y_t = ops.Act.empty(3, 2 * h, 2 * w, 16, torch.float32, dev, zero=True).slice(8, 2)
    y_f = ops.Act.empty(3, 2 * h, 2 * w, 8, torch.float32, dev, zero=True).slice(2, 2)
    y_f2 = ops.Act.empty(3, 2 * h, 2 * w, 2, torch.float32, dev)
    ops.deconv_col2im(taps, bt.to(dev), y_t, _act_from(fl, dev), pack_deconv(wf.to(dev)), bf.to(dev), y_f, y_f2)
    assert (y_t.to_nchw().cpu() - ref_t).abs().max() < 1e-4
    assert (y_f.to_nchw().cpu() - ref_f).abs().max() < 1e-4
    assert (y_f2.to_nchw().cpu() - ref_f).abs().max() < 1e-4
    # without the flow half
    y_t2 = ops.Act.empty(3, 2 * h, 2 * w, 2, torch.float32, dev)
    ops.deconv_col2im(taps, bt.to(dev), y_t2)
    assert (y_t2.to_nchw().cpu() - ref_t).abs().max() < 1e-4


@pytest.mark.parametrize('c,co,h,w', [(3, 16, 64, 64), (16, 32, 32, 32), (64, 96, 8, 8), (128, 196, 2, 2), (8, 16, 7, 9)])
def test_stride2_conv_as_space_to_depth_plus_tc(dev, c, co, h, w):
    """PWC extractor stride-2 layers: dbsr_space_to_depth2 (fp32 -> bf16) + the tcgen05 conv with the repacked weight"""
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_s2d_weight, pack_tc
    g = _gen(c + co)
    x = torch.randn(3, c, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(co, c, 3, 3, generator=g) / (9 * c) ** 0.5).bfloat16().float()
    b = torch.randn(co, generator=g)
    ref = O.lrelu(F.conv2d(x, wt, b, stride=2, padding=1))
    ho, wo = (h + 1) // 2, (w + 1) // 2
    xs = ops.Act.empty(3, ho, wo, (4 * c + 7) // 8 * 8, torch.bfloat16, dev, zero=True).slice(0, 4 * c)
    ops.space_to_depth2(_act_from(x, dev), xs)
    ref_s = F.pad(x, (0, wo * 2 - w, 0, ho * 2 - h)).view(3, c, ho, 2, wo, 2).permute(0, 3, 5, 1, 2, 4).reshape(3, 4 * c, ho, wo)
    assert (xs.to_nchw().cpu() - ref_s).abs().max() == 0
    if c % 8 == 0:      # bf16 -> bf16, 16-byte vector path
        xb = ops.Act(x.permute(0, 2, 3, 1).contiguous().to(dev).bfloat16())
        xs2 = ops.Act.empty(3, ho, wo, 4 * c + 8, torch.bfloat16, dev, zero=True).slice(8, 4 * c)
        ops.space_to_depth2(xb, xs2)
        assert (xs2.to_nchw().cpu() - ref_s).abs().max() == 0
    y = ops.Act.empty(3, ho, wo, (co + 7) // 8 * 8, torch.bfloat16, dev, zero=True).slice(0, co)
    ops.conv2d(xs, pack_tc(pack_s2d_weight(wt).to(dev)), b.to(dev), y, 3, 1, 1, ops.ACT_LRELU, tensor_core=True)
    got = y.to_nchw().cpu()
    assert (got - ref).abs().max() < 2e-2 * max(1.0, float(ref.abs().max()))


CORR_SHAPES = [(196, 1, 1), (128, 2, 2), (96, 4, 4), (64, 8, 8), (32, 16, 16), (32, 48, 48), (6, 5, 7), (64, 20, 33),
               (196, 3, 3), (128, 6, 6)]


@pytest.mark.parametrize('c,h,w', CORR_SHAPES)
def test_corr81_plain(dev, c, h, w):
    """every PWC level shape of the 48^2 / 80^2 / 96^2 / 160^2 / 256^2 configs incl. C=196, h=1"""
    from deep_rawburst_sr_b200 import ops
    g = _gen(c * 100 + h)
    f1 = torch.randn(3, c, h, w, generator=g)
    f2 = torch.randn(3, c, h, w, generator=g)
    ref = O.correlation81(f1, f2)
    out = ops.Act.empty(3, h, w, 96, torch.float32, dev, zero=True).slice(8, 81)
    ops.corr81(_act_from(f1, dev), _act_from(f2, dev), out, pairs=3, group=0)
    assert (out.to_nchw().cpu() - ref).abs().max() < 2e-5
    ref_l = O.lrelu(ref)
    ops.corr81(_act_from(f1, dev), _act_from(f2, dev), out, pairs=3, group=0, act=ops.ACT_LRELU)
    assert (out.to_nchw().cpu() - ref_l).abs().max() < 2e-5


@pytest.mark.parametrize('c,h,w,mag', [(32, 16, 16, 3.0), (196, 1, 1, 0.0), (128, 2, 2, 1.0), (64, 8, 8, 2.0), (100, 5, 9, 2.0)])
def test_corr81_bf16_vectorised_staging(dev, c, h, w, mag):
    """bf16 feature maps in 8-channel-aligned buffers (the layout the engine uses): vectorised halo staging, channel
    tails (C % 8 != 0) masked, with and without the fused backwarp"""
    from deep_rawburst_sr_b200 import ops
    g = _gen(c * 3 + h)
    pitch = (c + 7) // 8 * 8
    f1 = torch.randn(3, c, h, w, generator=g).bfloat16().float()
    f2 = torch.randn(3, c, h, w, generator=g).bfloat16().float()

    def act_bf16(t):
        buf = torch.full((3, h, w, pitch), 7.0, dtype=torch.bfloat16, device=dev)     # poison in the pad channels
        buf[..., :c] = t.permute(0, 2, 3, 1).to(dev).bfloat16()
        return ops.Act(buf).slice(0, c)

    out = ops.Act.empty(3, h, w, 88, torch.float32, dev, zero=True).slice(0, 81)
    ops.corr81(act_bf16(f1), act_bf16(f2), out, pairs=3, group=0)
    assert (out.to_nchw().cpu() - O.correlation81(f1, f2)).abs().max() < 2e-5
    # bf16 volume written into its 8-aligned concat segment (16-byte stores of 8 displacements; the 7 pad channels after
    # the 81st are zeroed, the neighbouring segments must stay untouched)
    cat = torch.full((3, h, w, 104), 3.0, dtype=torch.bfloat16, device=dev)
    ops.corr81(act_bf16(f1), act_bf16(f2), ops.Act(cat).slice(8, 81), pairs=3, group=0)
    ref81 = O.correlation81(f1, f2)
    got81 = cat[..., 8:89].float().permute(0, 3, 1, 2).cpu()
    assert (got81 - ref81).abs().max() <= 2e-5 + ref81.abs().max() * 2.0 ** -8
    # neighbouring segments untouched; the segment's own pad channels are either untouched (small-map kernel) or zeroed
    assert (cat[..., :8] == 3.0).all() and (cat[..., 96:] == 3.0).all()
    assert ((cat[..., 89:96] == 0) | (cat[..., 89:96] == 3.0)).all()
    if mag > 0:
        flow = (torch.rand(3, 2, h, w, generator=g) * 2 - 1) * mag
        ref = O.lrelu(O.correlation81(f1, O.backwarp(f2, flow * 1.25)))
        ops.corr81(act_bf16(f1), act_bf16(f2), out, pairs=3, group=0, flow=_act_from(flow, dev), flow_scale=1.25,
                   act=ops.ACT_LRELU)
        err = (out.to_nchw().cpu() - ref).abs()
        # tensor-core path (C in {32, 64, 96, 128}, map > 8x8): the warped map is rounded to bf16 (relative 2^-9 per value)
        f2w = O.backwarp(f2, flow * 1.25)
        bound = 5e-5 + 2.0 ** -9 * O.correlation81(f1.abs(), f2w.abs())
        assert (err > bound).any(dim=1).float().mean() < 0.01, float(err.max())
        # CUDA-core kernels (per-call switch): fp32 arithmetic on the bf16 inputs
        ops.corr81(act_bf16(f1), act_bf16(f2), out, pairs=3, group=0, flow=_act_from(flow, dev), flow_scale=1.25,
                   act=ops.ACT_LRELU, tensor_core=False)
        err = (out.to_nchw().cpu() - ref).abs()
        assert (err > 5e-5).any(dim=1).float().mean() < 0.01, float(err.max())


@pytest.mark.parametrize('c,h,w,mag', [(32, 32, 24, 0.0), (32, 32, 24, 5.0), (64, 16, 16, 3.0), (96, 12, 12, 2.0), (128, 9, 9, 1.5),
                                       (32, 13, 21, 8.0), (64, 48, 48, 0.0), (32, 10, 7, 0.0)])
def test_corr81_tensor_core_banded_product(dev, c, h, w, mag):
    """bf16 cost volume on the tensor cores (corr81_mma_kernel: mma.sync m16n8k16, diagonal scatter of the accumulator
    fragments): full / ragged tiles, every supported channel count, burst pair -> image mapping, fused backwarp, bf16 volume
    written into its concat slice.  Without a flow the products of bf16 values are exact and only the fp32 summation order
    differs from the oracle (<= 2e-5); with a flow the warped map is rounded to bf16: |err| <= 2^-9 * mean_c |f1||f2w|."""
    from deep_rawburst_sr_b200 import ops
    g = _gen(c * 7 + h + w)
    B, N = 2, 3
    feats = torch.randn(B * N, c, h, w, generator=g).bfloat16().float()
    f1 = feats.view(B, N, c, h, w)[:, :1].expand(-1, N - 1, -1, -1, -1).reshape(-1, c, h, w)
    f2 = feats.view(B, N, c, h, w)[:, 1:].reshape(-1, c, h, w)
    P = B * (N - 1)
    buf = torch.full((B * N, h, w, c + 8), 7.0, dtype=torch.bfloat16, device=dev)          # poison behind the channels
    buf[..., :c] = feats.permute(0, 2, 3, 1).to(dev).bfloat16()
    fa = ops.Act(buf).slice(0, c)
    out = ops.Act.empty(P, h, w, 88, torch.float32, dev, zero=True).slice(0, 81)
    if mag > 0:
        flow = (torch.rand(P, 2, h, w, generator=g) * 2 - 1) * mag
        f2w = O.backwarp(f2, flow * 2.5)
        ref = O.lrelu(O.correlation81(f1, f2w))
        bound = 5e-5 + 2.0 ** -9 * O.correlation81(f1.abs(), f2w.abs())
        kw = dict(flow=_act_from(flow, dev), flow_scale=2.5)
    else:
        ref = O.lrelu(O.correlation81(f1, f2))
        bound = torch.full_like(ref, 2e-5)
        kw = {}
    ops.corr81(fa, fa, out, pairs=P, group=N - 1, act=ops.ACT_LRELU, **kw)
    got = out.to_nchw().cpu()
    err = (got - ref).abs()
    assert (err > bound).any(dim=1).float().mean() < (0.01 if mag > 0 else 1e-9), float(err.max())
    # A/B against the CUDA-core kernel on the same inputs
    out2 = ops.Act.empty(P, h, w, 88, torch.float32, dev, zero=True).slice(0, 81)
    ops.corr81(fa, fa, out2, pairs=P, group=N - 1, act=ops.ACT_LRELU, tensor_core=False, **kw)
    d = (out2.to_nchw().cpu() - got).abs()
    assert (d > bound).any(dim=1).float().mean() < (0.01 if mag > 0 else 1e-9), float(d.max())
    # bf16 volume into an 8-aligned concat slice: neighbours untouched, pad channels zeroed, deterministic
    cat = torch.full((P, h, w, 104), 3.0, dtype=torch.bfloat16, device=dev)
    ops.corr81(fa, fa, ops.Act(cat).slice(8, 81), pairs=P, group=N - 1, act=ops.ACT_LRELU, **kw)
    assert torch.equal(cat[..., 8:89].float().permute(0, 3, 1, 2).cpu(), got.bfloat16().float())
    assert (cat[..., :8] == 3.0).all() and (cat[..., 96:] == 3.0).all() and (cat[..., 89:96] == 0).all()


def test_corr81_golden_reference_vector(dev, golden_dir):
    """the vector produced by the reference module's own code path (tests/golden/ops.npz)"""
    import os
    from deep_rawburst_sr_b200.external.pwcnet.correlation import correlation
    gold = np.load(os.path.join(golden_dir, 'ops.npz'))
    f1 = torch.from_numpy(gold['corr_f1']).to(dev)
    f2 = torch.from_numpy(gold['corr_f2']).to(dev)
    got = correlation.FunctionCorrelation(tenFirst=f1, tenSecond=f2).cpu().numpy()
    assert got.shape == gold['corr'].shape and np.abs(got - gold['corr']).max() < 2e-6
    with pytest.raises(NotImplementedError):
        correlation.FunctionCorrelation(tenFirst=f1.cpu(), tenSecond=f2.cpu())


@pytest.mark.parametrize('c,h,w,mag', [(32, 16, 16, 3.0), (64, 8, 8, 6.0), (128, 2, 2, 1.0), (96, 4, 5, 30.0), (32, 32, 24, 6.0)])
def test_corr81_backwarp_burst_mapping(dev, c, h, w, mag):
    """fused backwarp (flows up to +-mag px, incl. all taps out of bounds) + burst pair->image mapping"""
    from deep_rawburst_sr_b200 import ops
    g = _gen(c + h)
    B, N = 2, 4
    feats = torch.randn(B * N, c, h, w, generator=g)
    flow = (torch.rand(B * (N - 1), 2, h, w, generator=g) * 2 - 1) * mag
    scale = 1.25
    f1 = feats.view(B, N, c, h, w)[:, :1].expand(-1, N - 1, -1, -1, -1).reshape(-1, c, h, w)
    f2 = feats.view(B, N, c, h, w)[:, 1:].reshape(-1, c, h, w)
    ref = O.lrelu(O.correlation81(f1, O.backwarp(f2, flow * scale)))
    fa = _act_from(feats, dev)
    out = ops.Act.empty(B * (N - 1), h, w, 81, torch.float32, dev)
    ops.corr81(fa, fa, out, pairs=B * (N - 1), group=N - 1, flow=_act_from(flow, dev), flow_scale=scale,
               act=ops.ACT_LRELU)
    got = out.to_nchw().cpu()
    err = (got - ref).abs()
    # the 0.999 mask is a discontinuity: allow a handful of pixels to flip, everything else to fp32 round-off
    bad = (err > 5e-5).any(dim=1).float().mean()
    assert bad < 0.01, float(err.max())


@pytest.mark.parametrize('c,h,w,dtype,coff', [(32, 16, 16, torch.bfloat16, 88), (64, 8, 8, torch.bfloat16, 88), (128, 2, 2, torch.bfloat16, 88),
                                              (96, 4, 4, torch.bfloat16, 88), (32, 20, 24, torch.bfloat16, 88), (64, 8, 8, torch.float32, 88),
                                              (96, 4, 4, torch.float32, 88), (32, 16, 16, torch.float32, 88), (32, 16, 16, torch.bfloat16, 92)])
def test_corr81_writes_first_map_into_concat_slice(dev, c, h, w, dtype, coff):
    """dbsr_corr81_copy: the launch that builds the cost volume also writes the first map of every pair into the `tenFirst` slice
    of the decoder's concat buffer (pwcnet.py:173), with the burst pair->image mapping -- bit for bit what dbsr_copy_channels writes
    (tensor-core kernel, small-map kernel, and the paths that run the separate copy: fp32 tiled kernel, unaligned slice)"""
    from deep_rawburst_sr_b200 import ops
    g = _gen(c * h + w)
    B, N = 2, 4
    P = B * (N - 1)
    feats = torch.randn(B * N, c, h, w, generator=g)
    flow = (torch.rand(P, 2, h, w, generator=g) * 2 - 1) * 2.0
    fa = _act_from(feats, dev, dtype=dtype)
    fl = _act_from(flow, dev)
    pitch = 88 + c + 16
    cat_a = ops.Act(torch.zeros((P, h, w, pitch), dtype=dtype, device=dev))
    cat_b = ops.Act(torch.zeros((P, h, w, pitch), dtype=dtype, device=dev))
    ops.corr81(fa, fa, cat_a.slice(0, 81), pairs=P, group=N - 1, flow=fl, flow_scale=1.25, act=ops.ACT_LRELU,
               f1_copy=cat_a.slice(coff, c))
    ops.corr81(fa, fa, cat_b.slice(0, 81), pairs=P, group=N - 1, flow=fl, flow_scale=1.25, act=ops.ACT_LRELU)
    ops.copy_channels(fa, cat_b.slice(coff, c), N - 1, N, 0)
    torch.cuda.synchronize()
    assert torch.equal(cat_a.buf, cat_b.buf)
    want = feats.view(B, N, c, h, w)[:, :1].expand(-1, N - 1, -1, -1, -1).reshape(P, c, h, w).to(dtype).float()
    assert torch.equal(cat_a.slice(coff, c).to_nchw().cpu().float(), want)
    # pair mode (group 0): image p -> pair p
    cat_c = ops.Act(torch.zeros((P, h, w, pitch), dtype=dtype, device=dev))
    ops.corr81(fa.images(0, P), fa.images(1, P), cat_c.slice(0, 81), pairs=P, group=0, flow=fl, flow_scale=1.25, act=ops.ACT_LRELU,
               f1_copy=cat_c.slice(coff, c))
    assert torch.equal(cat_c.slice(coff, c).to_nchw().cpu().float(), feats[:P].to(dtype).float())


def _guarded(n, h, w, pitch, dtype, dev, fill):
    """an Act over the MIDDLE images of a larger allocation whose first and last image are a sentinel: an out-of-bounds store of
    a kernel (before or after the view, or into channels outside its slice) changes a sentinel value"""
    from deep_rawburst_sr_b200 import ops
    big = torch.full((n + 2, h, w, pitch), fill, dtype=dtype, device=dev)
    return big, ops.Act(big[1:n + 1])


@pytest.mark.parametrize('n,h,w', [(1, 5, 7), (2, 37, 9), (3, 70, 24), (1, 384, 48)])
def test_store_paths_stay_inside_their_views(dev, n, h, w):
    """guard bands around the outputs of the kernels whose store loops changed in round 2: blur3x3 with every rows-per-thread
    choice (1..3 images, ragged last row block), the cost volume with the folded first-map copy (tensor-core and small-map
    kernels), a tensor-core conv with the halved N tile.  Sentinel images before / after the view and the channels outside the
    written slices must keep their value; the written values are checked by the parity tests."""
    from deep_rawburst_sr_b200 import ops
    g = _gen(n * h + w)
    S = 7.0
    # blur: bf16 [n, h, w, 32] inside a 48-channel pitch at offset 8
    x = _act_from(torch.rand(n, 32, h, w, generator=g), dev, dtype=torch.bfloat16)
    big, ya = _guarded(n, h, w, 48, torch.bfloat16, dev, S)
    ops.blur3x3(x, ya.slice(8, 32), O.gauss_kernel3().reshape(-1).tolist())
    torch.cuda.synchronize()
    assert bool((big[0] == S).all()) and bool((big[-1] == S).all())
    assert bool((big[1:-1, :, :, :8] == S).all()) and bool((big[1:-1, :, :, 40:] == S).all())
    assert not bool((big[1:-1, :, :, 8:40] == S).any())
    if h <= 70:
        # cost volume + first-map copy into a concat buffer [V 81 (+7) | f1 C | rest]
        for c in (32, 64):
            feats = _act_from(torch.randn(2 * n, c, h, w, generator=g), dev, dtype=torch.bfloat16)
            flow = _act_from((torch.rand(n, 2, h, w, generator=g) * 2 - 1) * 3, dev)
            pitch = 88 + c + 8
            big, cat = _guarded(n, h, w, pitch, torch.bfloat16, dev, S)
            if h > 1 and w > 1:
                ops.corr81(feats.images(0, n), feats.images(n, n), cat.slice(0, 81), pairs=n, flow=flow, flow_scale=1.0,
                           act=ops.ACT_LRELU, f1_copy=cat.slice(88, c))
                torch.cuda.synchronize()
                assert bool((big[0] == S).all()) and bool((big[-1] == S).all())
                assert bool((big[1:-1, :, :, 88 + c:] == S).all())
                assert torch.equal(cat.slice(88, c).to_nchw(), feats.images(0, n).to_nchw())


@pytest.mark.parametrize('coff', [32, 33])          # 8-byte aligned (oc pair in one load) / odd offset (scalar loads)
@pytest.mark.parametrize('n,h,w', [(3, 1, 1), (2, 2, 2), (2, 4, 5), (1, 8, 8), (2, 16, 16)])
def test_flow_head_as_tap_planes(dev, n, h, w, coff):
    """pwcnet.py:150 (netSix: Conv2d(Cin, 2, 3, 1, 1)) split into a 1x1 channel contraction to 18 (tap, oc) planes and a tap sum:
    (a) `flow_from_taps` of the planes of a reference conv equals that conv; (b) `deconv_col2im` fed with the planes (+ bias)
    gives bit for bit what it gives when fed with the flow map those planes sum to."""
    from deep_rawburst_sr_b200 import ops
    g = _gen(n * 31 + h)
    cin = 24
    x = torch.randn(n, cin, h, w, generator=g)
    w6 = torch.randn(2, cin, 3, 3, generator=g) * 0.2
    b6 = torch.randn(2, generator=g)
    ref = F.conv2d(x, w6, b6, padding=1)
    planes = F.conv2d(x, w6.permute(2, 3, 0, 1).reshape(18, cin, 1, 1))              # [(ky*3+kx)*2+oc] = w6[oc, :, ky, kx] . x
    # the planes live in a wider buffer next to the 32 netUpfeat planes, as in the engine
    taps32 = torch.randn(n, 32, h, w, generator=g)
    buf = ops.Act(torch.zeros((n, h, w, 64), dtype=torch.float32, device=dev))
    buf.slice(0, 32).from_nchw(taps32.to(dev).contiguous())
    buf.slice(coff, 18).from_nchw(planes.to(dev).contiguous())
    flow = ops.Act.empty(n, h, w, 2, torch.float32, dev)
    ops.flow_from_taps(buf.slice(coff, 18), b6.to(dev), flow)
    assert (flow.to_nchw().cpu() - ref).abs().max() < 1e-5
    wf = torch.randn(4, 4, 2, 2, generator=g).to(dev)
    bf, bt = torch.randn(2, generator=g).to(dev), torch.randn(2, generator=g).to(dev)
    outs = []
    for from_planes in (False, True):
        y_t = ops.Act.empty(n, 2 * h, 2 * w, 2, torch.float32, dev)
        y_f = ops.Act.empty(n, 2 * h, 2 * w, 2, torch.float32, dev)
        y_f2 = ops.Act.empty(n, 2 * h, 2 * w, 2, torch.float32, dev)
        if from_planes:
            ops.deconv_col2im(buf.slice(0, 32), bt, y_t, None, wf, bf, y_f, y_f2, flow_taps=buf.slice(coff, 18), flow_bias=b6.to(dev))
        else:
            ops.deconv_col2im(buf.slice(0, 32), bt, y_t, flow, wf, bf, y_f, y_f2)
        outs.append((y_t.buf.clone(), y_f.buf.clone(), y_f2.buf.clone()))
    for a_, b_ in zip(*outs):
        assert torch.equal(a_, b_)
    with pytest.raises(Exception):
        ops.deconv_col2im(buf.slice(0, 32), bt, y_t, flow, wf, bf, y_f, None, flow_taps=buf.slice(coff, 18), flow_bias=b6.to(dev))


def test_prep_burst_and_flow_head(dev):
    from deep_rawburst_sr_b200 import ops
    g = _gen(11)
    burst = torch.rand(2, 3, 4, 24, 40, generator=g)
    enc_in = ops.Act.empty(6, 24, 40, 8, torch.float32, dev)
    pwc_in = ops.Act.empty(6, 64, 64, 4, torch.float32, dev)
    ops.prep_burst(burst.to(dev), enc_in, pwc_in)
    assert (enc_in.slice(0, 4).to_nchw().cpu() - burst.view(6, 4, 24, 40)).abs().max() == 0
    assert float(enc_in.buf[..., 4:].abs().max()) == 0
    ref = O.resize_bilinear(O.rggb_to_rgb(burst).view(6, 3, 24, 40), 64, 64)
    assert (pwc_in.slice(0, 3).to_nchw().cpu() - ref).abs().max() < 1e-6
    flow4 = torch.randn(5, 2, 16, 16, generator=g)
    offsets = torch.empty(5, 2, 24, 40, device=dev)
    ops.flow_head(_act_from(flow4, dev), offsets, 24, 40, 64, 64)
    ref = 20.0 * O.resize_bilinear(flow4, 24, 40)
    ref = torch.stack((ref[:, 0] * (40 / 64.0), ref[:, 1] * (24 / 64.0)), 1)
    assert (offsets.cpu() - ref).abs().max() < 2e-6 * max(1.0, float(ref.abs().max()))   # fp32 round-off, |flow| ~ 20


@pytest.mark.parametrize('H,W', [(24, 40), (48, 48), (80, 80), (64, 128)])
def test_prep_burst_s2d_equals_prep_plus_space_to_depth(dev, H, W):
    """the fused preparation of the bf16 PWC-Net path (RGGB->RGB + resize + space-to-depth + bf16) is bit-identical to
    prep_burst followed by space_to_depth2, and its encoder input equals the packed RAW frame (bf16-rounded)"""
    from deep_rawburst_sr_b200 import ops
    burst = torch.rand(2, 3, 4, H, W, generator=_gen(H + W)).to(dev)
    Hp, Wp = (H + 63) // 64 * 64, (W + 63) // 64 * 64
    enc_a = ops.Act.empty(6, H, W, 8, torch.bfloat16, dev)
    pwc_in = ops.Act.empty(6, Hp, Wp, 4, torch.float32, dev)
    ops.prep_burst(burst, enc_a, pwc_in)
    ref = ops.Act.empty(6, Hp // 2, Wp // 2, 16, torch.bfloat16, dev, zero=True).slice(0, 12)
    ops.space_to_depth2(pwc_in.slice(0, 3), ref)
    enc_b = ops.Act(torch.full((6, H, W, 8), 9.0, dtype=torch.bfloat16, device=dev))
    got = ops.Act(torch.full((6, Hp // 2, Wp // 2, 16), 9.0, dtype=torch.bfloat16, device=dev)).slice(0, 12)
    ops.prep_burst_s2d(burst, enc_b, got, Hp, Wp)
    assert torch.equal(got.buf, ref.buf)          # pad channels included (zero)
    assert torch.equal(enc_b.buf, enc_a.buf)
    assert torch.equal(enc_b.buf[..., :4].float(), burst.view(6, 4, H, W).permute(0, 2, 3, 1).bfloat16().float())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_warp(dev, dtype, golden_dir):
    import os
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.models.layers.warp import warp as warp_b200
    gold = np.load(os.path.join(golden_dir, 'ops.npz'))
    if dtype == torch.float32:
        got = warp_b200(torch.from_numpy(gold['feat']).to(dev), torch.from_numpy(gold['flow']).to(dev)).cpu().numpy()
        assert np.abs(got - gold['warp']).max() < 2e-5
    g = _gen(3)
    B, N, C, H, W = 2, 3, 64, 12, 20
    feat = torch.randn(B * N, C, H, W, generator=g)
    if dtype == torch.bfloat16:
        feat = feat.bfloat16().float()
    offs = (torch.rand(B * (N - 1), 2, H, W, generator=g) * 2 - 1) * 6.0
    offs[0, :, 0, 0] = torch.tensor([-100.0, 50.0])
    fa = _act_from(feat, dev, dtype=dtype)
    out = ops.Act.empty(B * N, H, W, C, dtype, dev)
    ops.warp(fa, offs.to(dev), out, frames=N)
    f5 = feat.view(B, N, C, H, W)
    ref = torch.cat([f5[:, :1], O.warp(f5[:, 1:].reshape(-1, C, H, W), offs).view(B, N - 1, C, H, W)], 1)
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    assert (out.to_nchw().cpu() - ref.reshape(-1, C, H, W)).abs().max() < tol


def test_offsets_mod_and_wp_input(dev):
    from deep_rawburst_sr_b200 import ops
    g = _gen(4)
    B, N, H, W = 2, 3, 5, 7
    offs = (torch.rand(B * (N - 1), 2, H, W, generator=g) * 2 - 1) * 3
    offs[0, 0, 0, :5] = torch.tensor([-0.25, 1.75, -1e-9, 0.0, 3.0])
    out = ops.Act.empty(B * N, H, W, 8, torch.float32, dev)
    ops.offsets_mod(offs.to(dev), out, B, N, 1.0)
    ref = torch.cat([torch.zeros(B, 1, 2, H, W), torch.remainder(offs, 1.0).view(B, N - 1, 2, H, W)], 1).view(-1, 2, H, W)
    got = out.slice(0, 2).to_nchw().cpu()
    assert torch.equal(got, ref), (got - ref).abs().max()
    assert float(out.buf[..., 2:].abs().max()) == 0
    # the engine's layout: dense 8-channel bf16 rows (one 16-byte store per pixel)
    out16 = ops.Act(torch.full((B * N, H, W, 8), 5.0, dtype=torch.bfloat16, device=dev))
    ops.offsets_mod(offs.to(dev), out16, B, N, 1.0)
    assert torch.equal(out16.slice(0, 2).to_nchw().cpu(), ref.bfloat16().float())
    assert float(out16.buf[..., 2:].float().abs().max()) == 0
    proj = torch.randn(B * N, 16, H, W, generator=g)
    wp = ops.Act.empty(B * N, H, W, 48, torch.float32, dev, zero=True)
    ops.build_wp_input(_act_from(proj, dev), wp, N)
    p5 = proj.view(B, N, 16, H, W)
    refw = torch.cat([p5[:, :1].expand(-1, N, -1, -1, -1), p5 - p5[:, :1]], 2).reshape(B * N, 32, H, W)
    assert torch.equal(wp.slice(0, 32).to_nchw().cpu(), refw)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('with_offsets', [False, True])
def test_softmax_wsum(dev, with_offsets, dtype):
    from deep_rawburst_sr_b200 import ops
    g = _gen(9)
    B, N, C, H, W = 2, 5, 64, 6, 9
    feat = torch.rand(B * N, C, H, W, generator=g)
    logits = torch.randn(B * N, C, H, W, generator=g) * 3
    if dtype == torch.bfloat16:      # bf16 path (two-pass register-resident kernel): same rounded operands on both sides
        feat, logits = feat.bfloat16().float(), logits.bfloat16().float()
    logits[0, :, 0, 0] = 30.0
    logits[1, :, 0, 0] = -30.0     # overflow guard
    offs = (torch.rand(B * (N - 1), 2, H, W, generator=g) * 2 - 1) * 4
    f5 = feat.view(B, N, C, H, W)
    if with_offsets:
        a5 = torch.cat([f5[:, :1], O.warp(f5[:, 1:].reshape(-1, C, H, W), offs).view(B, N - 1, C, H, W)], 1)
    else:
        a5 = f5
    w = torch.softmax(logits.view(B, N, C, H, W), dim=1)
    ref = (a5 * w).sum(1)
    fused = ops.Act.empty(B, H, W, C, dtype, dev)
    wout = torch.empty(B, N, C, H, W, device=dev)
    ops.softmax_wsum(_act_from(feat, dev, dtype=dtype), _act_from(logits, dev, dtype=dtype), fused, N,
                     offsets=offs.to(dev) if with_offsets else None, weights_out=wout)
    assert (fused.to_nchw().cpu() - ref).abs().max() < (1e-5 if dtype == torch.float32 else 6e-3)   # bf16 output rounding
    assert (wout.cpu() - w).abs().max() < 1e-6


@pytest.mark.parametrize('N', [2, 3, 4, 14, 17])
@pytest.mark.parametrize('C,H,W', [(64, 6, 9), (512, 9, 8), (72, 5, 5)])
def test_softmax_wsum_bf16_gather_frame_counts(dev, N, C, H, W):
    """the bf16 warp-on-the-fly kernel pipelines the burst in PAIRS of frames: odd / even numbers of other frames, a
    single other frame, the maximum burst length, channel counts with a partial 64-channel slice, flows that leave the
    image (zero padding) and logits far apart (overflow guard)"""
    from deep_rawburst_sr_b200 import ops
    g = _gen(N * 7 + C)
    B = 2
    feat = torch.rand(B * N, C, H, W, generator=g).bfloat16().float()
    logits = (torch.randn(B * N, C, H, W, generator=g) * 3).bfloat16().float()
    logits[0, :, 0, 0] = 40.0
    logits[N - 1, :, 1, 1] = -40.0
    logits[N - 1, :, 2, 2] = 60.0
    offs = (torch.rand(B * (N - 1), 2, H, W, generator=g) * 2 - 1) * 5
    offs[0, :, 0, 0] = torch.tensor([-100.0, 3.0])       # far outside
    offs[0, :, 0, 1] = torch.tensor([0.0, 0.0])          # exactly on the grid
    offs[0, :, H - 1, W - 1] = torch.tensor([0.5, 0.5])  # half of the taps outside
    f5 = feat.view(B, N, C, H, W)
    a5 = torch.cat([f5[:, :1], O.warp(f5[:, 1:].reshape(-1, C, H, W), offs).view(B, N - 1, C, H, W)], 1)
    w = torch.softmax(logits.view(B, N, C, H, W), dim=1)
    ref = (a5 * w).sum(1)
    fused = ops.Act.empty(B, H, W, C, torch.bfloat16, dev)
    ops.softmax_wsum(_act_from(feat, dev, dtype=torch.bfloat16), _act_from(logits, dev, dtype=torch.bfloat16), fused, N,
                     offsets=offs.to(dev))
    err = (fused.to_nchw().cpu() - ref).abs().max().item()
    assert err < 6e-3, err          # bf16 output rounding (values in [0, 1))


def test_blur_and_predictor(dev):
    from deep_rawburst_sr_b200 import ops
    g = _gen(2)
    x = torch.rand(2, 32, 16, 24, generator=g)
    K = O.gauss_kernel3()
    ref = F.conv2d(x.view(-1, 1, 16, 24), K.view(1, 1, 3, 3), padding=1).view(2, 32, 16, 24)
    y = ops.Act.empty(2, 16, 24, 32, torch.float32, dev)
    ops.blur3x3(_act_from(x, dev), y, K.reshape(-1).tolist())          # Gaussian = rank 1: separable kernel
    assert (y.to_nchw().cpu() - ref).abs().max() < 1e-6
    # a kernel that is NOT separable takes the 9-tap path; tall maps cross the 32-row blocks of the separable one
    K9 = torch.rand(3, 3, generator=g)
    for kk in (K9, torch.outer(torch.tensor([0.2, 0.5, 0.3]), torch.tensor([0.1, 0.7, 0.2]))):
        xt = torch.rand(1, 8, 70, 9, generator=g)
        reft = F.conv2d(xt.view(-1, 1, 70, 9), kk.view(1, 1, 3, 3), padding=1).view(1, 8, 70, 9)
        yt = ops.Act.empty(1, 70, 9, 8, torch.float32, dev)
        ops.blur3x3(_act_from(xt, dev), yt, kk.reshape(-1).tolist())
        assert (yt.to_nchw().cpu() - reft).abs().max() < 1e-6
    # bf16 maps take a packed-fp32 (two values per instruction) form of the separable kernel: same operations, same order --
    # its output must be the bf16 rounding of what the fp32 kernel computes from the same bf16-rounded input, bit for bit
    for (nn, hh, ww) in ((2, 70, 9), (1, 33, 40)):
        xb = torch.rand(nn, 32, hh, ww, generator=g).bfloat16().float()
        y32 = ops.Act.empty(nn, hh, ww, 32, torch.float32, dev)
        ops.blur3x3(_act_from(xb, dev), y32, K.reshape(-1).tolist())
        y16 = ops.Act.empty(nn, hh, ww, 32, torch.bfloat16, dev)
        ops.blur3x3(_act_from(xb, dev, dtype=torch.bfloat16), y16, K.reshape(-1).tolist())
        assert torch.equal(y16.buf, y32.buf.bfloat16())
    wt = torch.randn(3, 32, generator=g)
    b = torch.randn(3, generator=g)
    pred = torch.empty(2, 3, 16, 24, device=dev)
    ops.predictor(_act_from(x, dev), wt.to(dev), b.to(dev), pred)
    refp = torch.relu(F.conv2d(x, wt.view(3, 32, 1, 1), b))
    assert (pred.cpu() - refp).abs().max() < 1e-5


def test_backwarp_seam(dev, golden_dir):
    import os
    from deep_rawburst_sr_b200.models.alignment.pwcnet import backwarp
    gold = np.load(os.path.join(golden_dir, 'ops.npz'))
    got = backwarp(torch.from_numpy(gold['feat']).to(dev), torch.from_numpy(gold['flow']).to(dev)).cpu().numpy()
    err = np.abs(got - gold['backwarp'])
    assert (err > 5e-5).mean() < 0.01


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('with_offsets', [False, True])
def test_warp_proj(dev, dtype, with_offsets):
    """p_n = relu(warp(q_n) + b); wp_in = [p_0 | p_n - p_0]  (projection commuted with the warp)"""
    from deep_rawburst_sr_b200 import ops
    g = _gen(21)
    B, N, C, H, W = 2, 4, 64, 9, 12
    q = torch.randn(B * N, C, H, W, generator=g)
    if dtype == torch.bfloat16:
        q = q.bfloat16().float()
    bias = torch.randn(C, generator=g)
    offs = (torch.rand(B * (N - 1), 2, H, W, generator=g) * 2 - 1) * 5
    q5 = q.view(B, N, C, H, W)
    if with_offsets:
        a5 = torch.cat([q5[:, :1], O.warp(q5[:, 1:].reshape(-1, C, H, W), offs).view(B, N - 1, C, H, W)], 1)
    else:
        a5 = q5
    p5 = torch.relu(a5 + bias.view(1, 1, C, 1, 1))
    ref = torch.cat([p5[:, :1].expand(-1, N, -1, -1, -1), p5 - p5[:, :1]], 2).reshape(B * N, 2 * C, H, W)
    wp = ops.Act.empty(B * N, H, W, 3 * C, dtype, dev, zero=True)
    ops.warp_proj(_act_from(q, dev, dtype=dtype), bias.to(dev), wp, N, offs.to(dev) if with_offsets else None)
    got = wp.slice(0, 2 * C).to_nchw().cpu()
    tol = 1e-5 if dtype == torch.float32 else 4e-2
    assert (got - ref).abs().max() < tol
    assert float(wp.buf[..., 2 * C:].float().abs().max()) == 0.0
