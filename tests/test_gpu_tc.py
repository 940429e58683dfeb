"""GPU: the tcgen05 / TMEM implicit-GEMM convolution (dbsr_conv2d_tc) against a CPU fp32 convolution of the SAME
bf16-rounded operands (so the only differences are fp32 accumulation order and the bf16 rounding of the output).
Tolerance: 2^-8 relative to the output scale for bf16 outputs, 1e-4 for fp32 outputs."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

from oracle import dbsr_oracle as O  # noqa: E402

# name: cin, cout, k, dil, n, h, w, act, residual, out_fp32, shuffle
TC_CASES = {
    'enc_res_64x64': (64, 64, 3, 1, 3, 48, 48, 1, True, False, 0),
    'enc_init_4x64': (4, 64, 3, 1, 2, 16, 16, 1, False, False, 0),
    'enc_out_64x512': (64, 512, 3, 1, 2, 16, 24, 1, False, False, 0),
    'proj_1x1_512x64': (512, 64, 1, 1, 2, 16, 16, 1, False, False, 0),
    'wp0_192x128': (192, 128, 3, 1, 2, 16, 16, 1, False, False, 0),
    'wp_res_128x128': (128, 128, 3, 1, 2, 24, 40, 1, True, False, 0),
    'wp_out_128x512_f32': (128, 512, 3, 1, 1, 16, 16, 0, False, True, 0),
    'dec_init_512x64': (512, 64, 3, 1, 1, 16, 16, 1, False, False, 0),
    'post_res_32x32': (32, 32, 3, 1, 1, 64, 64, 1, True, False, 0),
    'ragged_20x13': (64, 64, 3, 1, 2, 20, 13, 2, False, False, 0),
    'offset_2x64': (2, 64, 3, 1, 2, 16, 16, 1, False, False, 0),
    'upsample_shuffle': (64, 2048, 1, 1, 2, 6, 5, 1, False, False, 8),
    'dilated_d2': (128, 128, 3, 2, 1, 32, 32, 2, False, False, 0),
    # PWC-Net long tail: odd channel counts (masked / padded N tiles), tiny maps, fp32 flow outputs, big dilations
    'pwc_flow_565x2_f32res': (565, 2, 3, 1, 3, 16, 16, 0, True, True, 0),
    'pwc_ext_196x196_1x1map': (196, 196, 3, 1, 5, 1, 1, 2, False, False, 0),
    'pwc_dec_136x128': (136, 128, 3, 1, 3, 8, 8, 2, False, False, 0),
    'pwc_dec_440x96': (440, 96, 3, 1, 3, 4, 4, 2, False, False, 0),
    'pwc_ext_16x16': (16, 16, 3, 1, 2, 32, 32, 2, False, False, 0),
    'pwc_ref_d8_128x96': (128, 96, 3, 8, 1, 16, 16, 2, False, False, 0),
    'pwc_ref_d16_96x64': (96, 64, 3, 16, 1, 32, 32, 2, False, False, 0),
    # small maps -> "flat" mode (several whole images per M tile); image counts not multiples of the packing factor
    'flat_1x1_196x196_n37': (196, 196, 3, 1, 37, 1, 1, 2, False, False, 0),
    'flat_2x2_565x64_n21': (565, 64, 3, 1, 21, 2, 2, 2, False, False, 0),
    'flat_4x4_437x96_n11': (437, 96, 3, 1, 11, 4, 4, 2, False, False, 0),
    'flat_4x4_629x2_f32': (629, 2, 3, 1, 7, 4, 4, 0, True, True, 0),
    'flat_2x3_128x128': (128, 128, 3, 1, 9, 2, 3, 1, True, False, 0),
    # 1x1 kernels on small maps: flat mode without a border (128 / (h*w) images per tile): the tap GEMMs of the flow heads
    'flat_k1_1x1_529x50_n326_f32': (529, 50, 1, 1, 326, 1, 1, 0, False, True, 0),
    'flat_k1_1x1_196x64_n300': (196, 64, 1, 1, 300, 1, 1, 2, False, False, 0),
    'flat_k1_2x2_661x50_n337_f32': (661, 50, 1, 1, 337, 2, 2, 0, False, True, 0),
    'flat_k1_4x4_629x50_n416_f32': (629, 50, 1, 1, 416, 4, 4, 0, False, True, 0),
    'flat_k1_8x8_597x50_n297_f32': (597, 50, 1, 1, 297, 8, 8, 0, False, True, 0),
    'flat_k1_3x5_64x32_res': (64, 32, 1, 1, 319, 3, 5, 1, True, False, 0),
    # many small images: an item holds TWO M tiles that share every weight tile -- flat mode (2 x ni images, odd image count
    # so the last item is half empty) and narrow regular maps (two consecutive 8x8 / 8x5 images per item)
    'flat_4x4_pair_209x128_n331': (209, 128, 3, 1, 331, 4, 4, 2, False, False, 0),
    'flat_2x2_pair_96x48_n1001_f32': (96, 48, 3, 1, 1001, 2, 2, 0, False, True, 0),
    'narrow_8x8_pair_136x128_n301': (136, 128, 3, 1, 301, 8, 8, 2, False, False, 0),
    'narrow_8x5_pair_64x64_n297_k1': (64, 64, 1, 1, 297, 8, 5, 1, False, False, 0),
    'narrow_20x8_pair_32x32_n300': (32, 32, 3, 1, 300, 20, 8, 1, False, False, 0),
    # CTA pairs (tcgen05 cta_group::2, M = 256 over two CTAs that share every weight tile): large N = 64 / 128 layers with at
    # least 592 two-tile items.  Odd tile counts (the last pair has one idle half), ragged tiles, resident and streamed weights,
    # tensor-core and epilogue-side residuals, several N tiles, 1x1 kernels, fp32 outputs.
    'cta_pair_64x64_res_n67': (64, 64, 3, 1, 67, 48, 48, 1, True, False, 0),
    'cta_pair_128x128_res_n70_ragged': (128, 128, 3, 1, 70, 40, 44, 1, True, False, 0),
    'cta_pair_64x512_n99': (64, 512, 3, 1, 99, 32, 48, 1, False, False, 0),
    'cta_pair_192x128_n67': (192, 128, 3, 1, 67, 48, 48, 1, False, False, 0),
    'cta_pair_1x1_512x64_n67': (512, 64, 1, 1, 67, 48, 48, 0, False, False, 0),
    'cta_pair_128x128_f32_n67': (128, 128, 3, 1, 67, 48, 48, 2, False, True, 0),
    # 1x1 maps, image count a multiple of 8: centre tap only, the n maps viewed as one (n/8) x 8 image
    'centre_1x1_196x196_n448': (196, 196, 3, 1, 448, 1, 1, 2, False, False, 0),
    'centre_1x1_512x196_n448_res': (512, 192, 3, 1, 448, 1, 1, 1, True, False, 0),
    'centre_1x1_529x2_f32_n416': (529, 2, 3, 1, 416, 1, 1, 0, False, True, 0),
    'k1_1x1_529x32_f32_n416': (529, 32, 1, 1, 416, 1, 1, 0, False, True, 0),
}


def run_case(name, dev):
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_tc, permute_shuffle_rows
    cin, cout, k, dil, n, h, w, act, use_res, out_f32, shuffle = TC_CASES[name]
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, wt, b, padding=dil * (k - 1) // 2, dilation=dil)
    res = None
    if use_res:
        res = torch.randn(ref.shape, generator=g).bfloat16().float()
        ref = ref + res
    ref = torch.relu(ref) if act == 1 else (O.lrelu(ref) if act == 2 else ref)
    if shuffle:
        ref = O.pixel_shuffle(ref, shuffle)
    pitch = max(8, (cin + 7) // 8 * 8)
    c_off = 0
    if name.startswith('pwc_dec'):       # a channel slice of a wider concat buffer
        c_off, pitch = 448, 448 + pitch + 16
    xa = ops.Act(torch.full((n, h, w, pitch), 7.0, dtype=torch.bfloat16, device=dev)).slice(c_off, cin)
    xa.from_nchw(x.to(dev))
    ydt = torch.float32 if out_f32 else torch.bfloat16
    if shuffle:
        ya = ops.Act.empty(n, h * shuffle, w * shuffle, cout // shuffle ** 2, ydt, dev)
    else:
        ya = ops.Act(torch.zeros((n, h, w, cout + 16), dtype=ydt, device=dev)).slice(16, cout)
    ra = None
    if res is not None:
        ra = ops.Act.empty(n, h, w, cout, ydt, dev).from_nchw(res.to(dev))
    wp = pack_tc(wt.to(dev), shuffle)
    bd = b.to(dev) if not shuffle else permute_shuffle_rows(b.to(dev), shuffle)   # bias follows the packed row order
    assert ops.conv2d_tc_supported(xa, wp, bd, ya, k, 1, dil, ra, shuffle)
    ops.conv2d(xa, wp, bd, ya, k, 1, dil, act, ra, shuffle, tensor_core=True)
    torch.cuda.synchronize()
    got = ya.to_nchw().cpu()
    err = (got - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    tol = 1e-4 * scale if out_f32 else scale * 2.0 ** -8
    return err, tol


@pytest.mark.parametrize('name', list(TC_CASES))
def test_conv2d_tc(name):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    err, tol = run_case(name, torch.device('cuda:0'))
    assert err <= tol, f'{name}: max abs err {err} > {tol}'


@pytest.mark.parametrize('n,h,w,use_res,k', [(2, 48, 40, True, 3), (1, 21, 19, True, 3), (3, 16, 16, False, 3), (301, 20, 8, True, 3),
                                             (2, 32, 32, False, 1)])
def test_conv2d_tc_fused_predictor(n, h, w, use_res, k):
    """dbsr_conv2d_tc_predictor: 32 -> 32 conv (+ tensor-core residual) + ReLU with the 1x1 predictor 32 -> 3 + ReLU folded
    into the epilogue (decoders.py:52,61), fp32 NCHW output, the conv's own output map untouched.  Reference: fp32 conv of
    the same bf16-rounded operands, predictor on the UNROUNDED activations (what the fused kernel computes).  Covers ragged
    tiles (21x19), the narrow-map image-pair item layout (20x8) and the 1x1 form."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_tc
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(n * 1000 + h)
    x = torch.randn(n, 32, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(32, 32, k, k, generator=g) / (32 * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(32, generator=g)
    pw = torch.randn(3, 32, generator=g) / 32 ** 0.5
    pb = torch.randn(3, generator=g) * 0.1
    mid = F.conv2d(x, wt, b, padding=(k - 1) // 2)
    res = None
    if use_res:
        res = torch.randn(mid.shape, generator=g).bfloat16().float()
        mid = mid + res
    ref = torch.relu(F.conv2d(torch.relu(mid), pw.view(3, 32, 1, 1), pb))
    xa = ops.Act.empty(n, h, w, 32, torch.bfloat16, dev).from_nchw(x.to(dev))
    ya = ops.Act(torch.full((n, h, w, 32), 5.0, dtype=torch.bfloat16, device=dev))
    ra = ops.Act.empty(n, h, w, 32, torch.bfloat16, dev).from_nchw(res.to(dev)) if use_res else None
    pred = torch.full((n, 3, h, w), -1.0, device=dev)
    ops.conv2d_tc_predictor(xa, pack_tc(wt.to(dev)), b.to(dev), ya, k, ops.ACT_RELU, ra, pw, pb, pred)
    torch.cuda.synchronize()
    err = (pred.cpu() - ref).abs().max().item()
    assert err <= 1e-4 * max(1.0, ref.abs().max().item()), err
    assert (ya.buf == 5.0).all()        # the conv's output map is not written in this mode


def _fuzz_cases():
    """seeded random conv shapes around the decision boundaries of the tcgen05 planner: CTA pairs on / off (>= 592 two-tile
    items, K >= 128, N tile 64 / 128), ragged tiles in both directions, odd image counts (idle half of the last pair), 1x1
    and 3x3 kernels, with / without bias and residual, channel counts that need padded K chunks / masked N tiles"""
    import random
    rnd = random.Random(20260)
    cases = []
    for i in range(28):
        cin = rnd.choice([64, 128, 128, 192, 136, 256])
        cout = rnd.choice([64, 64, 128, 128, 96, 256])
        k = rnd.choice([3, 3, 3, 1])
        h, w = rnd.choice([(16, 16), (20, 24), (32, 48), (17, 33), (48, 40), (9, 31)])
        tiles2 = ((w + 15) // 16) * ((h + 15) // 16)
        n = rnd.choice([3, max(2, 592 // tiles2 + rnd.choice([-1, 0, 1, 2])), max(2, 700 // tiles2 + 1)])
        cases.append((i, cin, cout, k, n, h, w, rnd.choice([0, 1, 2]), rnd.random() < 0.4, rnd.random() < 0.8))
    return cases


@pytest.mark.parametrize('case', _fuzz_cases(), ids=lambda c: 'fuzz%d_%dx%d_k%d_n%d_%dx%d' % c[:7])
def test_conv2d_tc_fuzz(case):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_tc
    i, cin, cout, k, n, h, w, act, use_res, use_bias = case
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(1000 + i)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g) if use_bias else None
    res = torch.randn(n, cout, h, w, generator=g).bfloat16().float() if use_res else None
    pitch = (cin + 7) // 8 * 8
    xa = ops.Act(torch.full((n, h, w, pitch + 8), 7.0, dtype=torch.bfloat16, device=dev)).slice(8, cin).from_nchw(x.to(dev))
    ya = ops.Act(torch.full((n, h, w, cout + 8), 3.0, dtype=torch.bfloat16, device=dev)).slice(0, cout)
    ra = ops.Act.empty(n, h, w, cout, torch.bfloat16, dev).from_nchw(res.to(dev)) if use_res else None
    wp = pack_tc(wt.to(dev))
    bd = b.to(dev) if use_bias else None
    assert ops.conv2d_tc_supported(xa, wp, bd, ya, k, 1, 1, ra, 0)
    ops.conv2d(xa, wp, bd, ya, k, 1, 1, act, ra, 0, tensor_core=True)
    torch.cuda.synchronize()
    assert (ya.buf[..., cout:] == 3.0).all(), 'pad channels behind the output view were written'
    got = ya.to_nchw().cpu()
    # CPU reference of the first, a middle and the last images (the last ones sit in the half-empty pair / ragged items)
    for lo in sorted({0, max(0, n // 2 - 1), max(0, n - 3)}):
        sl = slice(lo, min(n, lo + 3))
        ref = F.conv2d(x[sl], wt, b, padding=(k - 1) // 2)
        if use_res:
            ref = ref + res[sl]
        ref = torch.relu(ref) if act == 1 else (O.lrelu(ref) if act == 2 else ref)
        err = (got[sl] - ref).abs().max().item()
        assert err <= max(1.0, ref.abs().max().item()) * 2.0 ** -8, (case, lo, err)


@pytest.mark.parametrize('bursts,frames,h,w', [(3, 5, 24, 40), (48, 14, 48, 48)])
def test_conv2d_tc_broadcast_residual(bursts, frames, h, w):
    """dbsr_conv_t.residual_group: output image i adds residual image i // group -- one map per burst broadcast over its
    frames (the per-burst term of the split weight-predictor conv), accumulated on the tensor core; single CTAs (small case)
    and CTA pairs (672 images)"""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_tc
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(bursts * 100 + frames)
    n = bursts * frames
    x = torch.randn(n, 128, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(128, 128, 3, 3, generator=g) / (128 * 9) ** 0.5).bfloat16().float()
    res = torch.randn(bursts, 128, h, w, generator=g).bfloat16().float()
    xa = ops.Act.empty(n, h, w, 128, torch.bfloat16, dev).from_nchw(x.to(dev))
    ra = ops.Act.empty(bursts, h, w, 128, torch.bfloat16, dev).from_nchw(res.to(dev))
    ya = ops.Act.empty(n, h, w, 128, torch.bfloat16, dev)
    wp = pack_tc(wt.to(dev))
    assert ops.conv2d_tc_supported(xa, wp, None, ya, 3, 1, 1, ra, 0, residual_group=frames)
    ops.conv2d(xa, wp, None, ya, 3, 1, 1, ops.ACT_RELU, ra, 0, tensor_core=True, residual_group=frames)
    torch.cuda.synchronize()
    got = ya.to_nchw().cpu()
    for b in (0, bursts // 2, bursts - 1):          # the CPU reference of a few bursts is enough
        sl = slice(b * frames, (b + 1) * frames)
        ref = torch.relu(F.conv2d(x[sl], wt, padding=1) + res[b:b + 1])
        assert (got[sl] - ref).abs().max().item() <= max(1.0, ref.abs().max().item()) * 2.0 ** -8, b


@pytest.mark.parametrize('n,h,w,with_pred', [(2, 48, 48, False), (1, 16, 24, False), (3, 37, 53, False), (1, 80, 20, False),
                                             (2, 5, 7, False), (2, 48, 48, True), (1, 33, 50, True), (5, 64, 96, False)])
def test_resblock32_fused(n, h, w, with_pred):
    """dbsr_resblock32_tc: relu(x + conv2(relu(conv1(x) + b1)) + b2) in one launch (the intermediate map lives in shared
    memory) -- (a) BIT-IDENTICAL to the two-launch tcgen05 path (dbsr_conv2d_tc x 2, residual on the tensor core), incl.
    ragged tiles (37x53: tiles of 16x24), maps smaller than one tile, several items per CTA pipeline stage; (b) within
    bf16 rounding of the fp32 reference of the same bf16-rounded operands with the intermediate map rounded to bf16;
    (c) with the fused predictor: equal to dbsr_conv2d_tc_predictor on the same intermediate (fp32, 1e-6) and y untouched;
    poison around the input view must not leak (x is a channel slice of a wider buffer)."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from deep_rawburst_sr_b200 import ops
    from deep_rawburst_sr_b200.engine import pack_tc
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(n * 7919 + h * 31 + w)
    x = torch.randn(n, 32, h, w, generator=g).bfloat16().float()
    w1 = (torch.randn(32, 32, 3, 3, generator=g) / 288 ** 0.5).bfloat16().float()
    w2 = (torch.randn(32, 32, 3, 3, generator=g) / 288 ** 0.5).bfloat16().float()
    b1, b2 = torch.randn(32, generator=g) * 0.2, torch.randn(32, generator=g) * 0.2
    pw = torch.randn(3, 32, generator=g) / 32 ** 0.5
    pb = torch.randn(3, generator=g) * 0.1
    t_ref = torch.relu(F.conv2d(x, w1, b1, padding=1)).bfloat16().float()
    mid = torch.relu(x + F.conv2d(t_ref, w2, b2, padding=1))
    xbuf = torch.full((n, h, w, 48), 9.0, dtype=torch.bfloat16, device=dev)       # poison on both sides of the view
    xa = ops.Act(xbuf).slice(8, 32).from_nchw(x.to(dev))
    w1p, w2p, b1d, b2d = pack_tc(w1.to(dev)), pack_tc(w2.to(dev)), b1.to(dev), b2.to(dev)
    # two-launch path
    ta = ops.Act.empty(n, h, w, 32, torch.bfloat16, dev)
    y2 = ops.Act.empty(n, h, w, 32, torch.bfloat16, dev)
    ops.conv2d(xa, w1p, b1d, ta, 3, 1, 1, ops.ACT_RELU, None, tensor_core=True)
    if with_pred:
        pred2 = torch.full((n, 3, h, w), -1.0, device=dev)
        ops.conv2d_tc_predictor(ta, w2p, b2d, y2, 3, ops.ACT_RELU, xa, pw, pb, pred2)
        pred = torch.full((n, 3, h, w), -2.0, device=dev)
        assert ops.resblock32_tc_supported(xa, None, w1p, b1d, w2p, b2d, with_pred=True)
        ops.resblock32_tc(xa, None, w1p, b1d, w2p, b2d, pw, pb, pred)
        torch.cuda.synchronize()
        assert (pred - pred2).abs().max().item() <= 1e-6, (pred - pred2).abs().max().item()
        ref = torch.relu(F.conv2d(mid, pw.view(3, 32, 1, 1), pb))
        assert (pred.cpu() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
        q = torch.zeros((n, 3, h, w), dtype=torch.int16, device=dev)
        ops.resblock32_tc(xa, None, w1p, b1d, w2p, b2d, pw, pb, q)
        assert torch.equal(q, (pred.clamp(0.0, 1.0) * 2 ** 14).short())
        return
    ops.conv2d(ta, w2p, b2d, y2, 3, 1, 1, ops.ACT_RELU, xa, tensor_core=True)
    ybuf = torch.full((n, h, w, 40), 5.0, dtype=torch.bfloat16, device=dev)
    ya = ops.Act(ybuf).slice(8, 32)
    assert ops.resblock32_tc_supported(xa, ya, w1p, b1d, w2p, b2d)
    for _ in range(2):
        ops.resblock32_tc(xa, ya, w1p, b1d, w2p, b2d)
    torch.cuda.synchronize()
    # bit-identical whenever the two-launch path also accumulates the residual on the tensor core; launches of a few items
    # split their N tile (32 -> 16 columns) and then add the residual in the epilogue: one fp32 rounding order apart
    small = n * ((h + 15) // 16) * ((w + 15) // 16) * 2 <= 148
    d = (ybuf[..., 8:].float() - y2.buf.float()).abs().max().item()
    assert torch.equal(ybuf[..., 8:], y2.buf) or (small and d <= max(1.0, y2.buf.float().abs().max().item()) * 2.0 ** -7), d
    assert (ybuf[..., :8] == 5.0).all() and (xbuf[..., :8] == 9.0).all() and (xbuf[..., 40:] == 9.0).all()
    got = ya.to_nchw().cpu()
    scale = max(1.0, mid.abs().max().item())
    assert (got - mid).abs().max().item() <= scale * 2.0 ** -7


if __name__ == '__main__':
    # probe mode: `python tests/test_gpu_tc.py <case>` prints the error of one case (one process per case so a
    # device fault in one variant does not hide the others)
    import sys
    sys.path.insert(0, '.')
    nm = sys.argv[1]
    e, t = run_case(nm, torch.device('cuda:0'))
    print(f'TC_CASE {nm} err={e:.3e} tol={t:.3e} {"OK" if e <= t else "FAIL"}')
