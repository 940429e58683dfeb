"""CPU: host-side logic of the package -- interface / state_dict contract of the drop-in modules, weight packing
(checked numerically against torch convs of the original weights), unsupported-flag and CPU-input behaviour."""
import os
import pytest
import torch
import torch.nn.functional as F

from oracle import dbsr_oracle as O
from deep_rawburst_sr_b200.engine import PwcLayout, pack_deconv, pack_direct, pack_tc
from deep_rawburst_sr_b200.models.dbsr import dbsrnet as D
from deep_rawburst_sr_b200.models.dbsr.merging import WeightedSum
from deep_rawburst_sr_b200.models.alignment.pwcnet import PWCNet


def test_state_dict_contract_and_constructor():
    net = D.dbsrnet_default_synthetic()
    spec = O.state_dict_spec()
    sd = net.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in spec)
    assert len(list(net.buffers())) == 0
    net.load_state_dict(O.make_state_dict(0), strict=True)
    assert net.constructor.fun_name == 'dbsrnet_cvpr2021' and net.constructor.fun_module.endswith('models.dbsr.dbsrnet')
    assert hasattr(net, 'encoder') and hasattr(net, 'merging') and hasattr(net, 'decoder')
    assert isinstance(net.encoder.alignment_net, PWCNet) and hasattr(net.encoder.alignment_net, 'net')
    # ICNR structure of the freshly initialised upsampler (reference initializations.py:21-38)
    w = D.dbsrnet_default_synthetic().state_dict()['decoder.upsample_layer.conv_layer.0.weight']
    assert torch.equal(w[0], w[63]) and not torch.equal(w[0], w[64])


def test_unsupported_flags_raise():
    for kw in (dict(softmax=False), dict(use_base_frame=False), dict(ref_offset_noise=0.1), dict(use_bn=True),
               dict(activation='lrelu'), dict(use_offset=False)):
        base = dict(softmax=True, use_base_frame=True)
        base.update(kw)
        with pytest.raises(NotImplementedError):
            WeightedSum(512, 64, 64, **base)
    with pytest.raises(Exception):
        PWCNet(load_pretrained=True, weights_path=None)


def test_cpu_tensors_raise_not_implemented():
    from deep_rawburst_sr_b200.external.pwcnet.correlation import correlation
    with pytest.raises(NotImplementedError):
        correlation.FunctionCorrelation(tenFirst=torch.zeros(1, 4, 3, 3), tenSecond=torch.zeros(1, 4, 3, 3))
    net = D.dbsrnet_default_synthetic()
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 2, 4, 16, 16))


def test_pwc_layout_and_packing_equivalence():
    """conv over the padded in-place concat layout with packed weights == conv over the reference's torch.cat"""
    g = torch.Generator().manual_seed(0)
    lay = PwcLayout(3)
    assert lay.total % 8 == 0 and all(v % 8 == 0 for v in lay.off.values())
    assert lay.off['V'] == 448 and lay.sizes['f1'] == 64
    cm, start, length = lay.chmap_from('o2')
    cin = 128 + 128 + 81 + 64 + 2 + 2
    assert len(cm) == cin and start == lay.off['o2'] and length == lay.total - start
    w = torch.randn(96, cin, 3, 3, generator=g)
    x_orig = torch.randn(2, cin, 5, 6, generator=g)
    buf = torch.zeros(2, length, 5, 6)
    buf[:, cm] = x_orig
    packed = pack_direct(w, cm, length)                       # [9, length, 96]
    w_buf = packed.view(3, 3, length, 96).permute(3, 2, 0, 1)  # back to [Cout, Cin_buf, k, k]
    assert torch.allclose(F.conv2d(buf, w_buf, padding=1), F.conv2d(x_orig, w, padding=1), atol=1e-4)
    # deconv packing
    wd = torch.randn(cin, 2, 4, 4, generator=g)
    pd = pack_deconv(wd, cm, length)                          # [4,4,2,length]
    wd_buf = pd.permute(3, 2, 0, 1)
    assert torch.allclose(F.conv_transpose2d(buf, wd_buf, stride=2, padding=1),
                          F.conv_transpose2d(x_orig, wd, stride=2, padding=1), atol=1e-4)


def test_pack_tc_layout_and_shuffle_permutation():
    g = torch.Generator().manual_seed(1)
    w = torch.randn(64, 4, 3, 3, generator=g)
    p = pack_tc(w)
    assert p.dtype == torch.bfloat16 and tuple(p.shape) == (9, 64, 32)
    assert torch.equal(p[4, :, :4].float(), w[:, :, 1, 1].bfloat16().float()) and float(p[:, :, 4:].float().abs().max()) == 0
    assert tuple(pack_tc(torch.randn(128, 192, 3, 3, generator=g)).shape) == (9, 128, 192)
    wu = torch.randn(2048, 64, 1, 1, generator=g)
    pu = pack_tc(wu, shuffle_r=8)
    # packed row i*256 + j*32 + c holds original output channel c*64 + i*8 + j
    for (c, i, j) in ((0, 0, 0), (5, 3, 7), (31, 7, 1)):
        assert torch.equal(pu[0, i * 256 + j * 32 + c].float(), wu[c * 64 + i * 8 + j, :, 0, 0].bfloat16().float())


def test_icnr_matches_reference_structure():
    from deep_rawburst_sr_b200.models.layers.initializations import ICNR
    torch.manual_seed(0)
    k = ICNR(torch.zeros(2048, 64, 1, 1), 8)
    assert tuple(k.shape) == (2048, 64, 1, 1)
    assert torch.equal(k[0], k[63]) and torch.equal(k[64], k[127]) and not torch.equal(k[0], k[64])


def test_pack_s2d_weight_is_the_stride2_conv():
    """3x3 / stride 2 / pad 1 == 3x3 / stride 1 / pad 1 over the space-to-depth input with the repacked weight
    (how the PWC extractor's stride-2 layers, pwcnet.py:49-97, reach the tensor-core kernel); odd sizes included"""
    import torch
    import torch.nn.functional as F
    from deep_rawburst_sr_b200.engine import pack_s2d_weight
    g = torch.Generator().manual_seed(5)
    for (c, co, h, w) in [(3, 16, 8, 8), (16, 32, 6, 10), (5, 7, 7, 9), (8, 4, 2, 2), (4, 4, 1, 3)]:
        x = torch.randn(2, c, h, w, generator=g)
        wt = torch.randn(co, c, 3, 3, generator=g)
        ref = F.conv2d(x, wt, stride=2, padding=1)
        hp, wp = (h + 1) // 2 * 2, (w + 1) // 2 * 2
        xp = F.pad(x, (0, wp - w, 0, hp - h))
        xs = xp.view(2, c, hp // 2, 2, wp // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(2, 4 * c, hp // 2, wp // 2)
        got = F.conv2d(xs, pack_s2d_weight(wt), stride=1, padding=1)
        assert got.shape == ref.shape and (got - ref).abs().max() < 1e-4


def test_gaussian_filter_glue_matches_oracle():
    """get_gaussian_kernel / apply_kernel (reference models/layers/filtering.py:43-62) are device-agnostic torch glue:
    on CPU they must equal the oracle's explicit-index restatement (reflect padding included)"""
    import torch
    from deep_rawburst_sr_b200.models.layers.filtering import apply_kernel, get_gaussian_kernel
    from oracle import sca_oracle as S
    K, ksz = get_gaussian_kernel(sd=1.5)
    Ko, ksz_o = S.gaussian_kernel(1.5)
    assert ksz == ksz_o == 7 and tuple(K.shape) == (1, 1, 7, 7)      # conv2d weight shape, as the reference
    assert (K[0, 0] - Ko).abs().max() < 1e-7 and abs(float(K.sum()) - 1.0) < 1e-6
    x = torch.rand(2, 3, 11, 9, generator=torch.Generator().manual_seed(3))
    assert (apply_kernel(x, ksz, K) - S.apply_kernel(x, ksz, Ko)).abs().max() < 1e-6


def test_generator_transform_sampling_matches_reference_draws(golden_dir):
    """host logic of the burst generator: under the same `random` seed, `sample_transforms` / `get_tmat` (with
    cv2.getRotationMatrix2D written out) reproduce the matrices the reference's `single2lrburst` drew (recorded by
    oracle/make_golden_lrburst.py) to the last bits, frame 0 being the pure half-pixel shift."""
    import os
    import random
    import numpy as np
    from deep_rawburst_sr_b200.data import synthetic_burst_generation as G
    from oracle.make_golden_lrburst import CASES
    for name, seed, H, W, n, f, crop, params in CASES:
        g = np.load(os.path.join(golden_dir, name + '.npz'))
        random.seed(seed)
        t_mats = np.stack(G.sample_transforms((H, W), n, f, dict(params)))
        assert t_mats.shape == g['t_mats'].shape and np.abs(t_mats - g['t_mats']).max() <= 1e-12
        shift = f / 2.0 - 0.5
        assert np.array_equal(t_mats[0], np.array([[1.0, 0.0, shift], [0.0, 1.0, shift]]))
    # the inverse map handed to the kernel is the exact inverse of the forward matrix
    M = G._inverse_map(t_mats[1])
    fwd = np.vstack([t_mats[1], [0, 0, 1]])
    inv = np.vstack([M.reshape(2, 3), [0, 0, 1]])
    assert np.abs(fwd @ inv - np.eye(3)).max() <= 1e-9


def test_scoring_loops_refuse_unknown_metrics_and_format_like_the_reference():
    from deep_rawburst_sr_b200.evaluation.burstsr import compute_score as burstsr
    from deep_rawburst_sr_b200.evaluation.synburst import compute_score as synburst
    import pytest
    with pytest.raises(NotImplementedError):
        synburst.score_dataset(None, [], metrics=('psnr', 'lpips'))
    with pytest.raises(NotImplementedError):
        burstsr.score_dataset(None, [], None, metrics=('lpips',))
    text = synburst.generate_formatted_report({'DBSR': {'psnr': 39.0912, 'ssim': 0.9512, 'count': 300}})
    assert text == '\n          | psnr       | ssim       |\nDBSR      | 39.091     | 0.951      |\n'
    data = synburst.TensorBurstSet(torch.zeros(3, 2, 4, 8, 8), torch.zeros(3, 3, 64, 64))
    burst, gt, meta = data[1]
    assert len(data) == 3 and burst.shape == (2, 4, 8, 8) and gt.shape == (3, 64, 64) and meta['burst_name'] == '0001'
    b, g = data.batch(1, 3)
    assert b.shape[0] == 2 and g.shape[0] == 2


def test_png16_codec_matches_the_reference_writer(tmp_path, golden_dir):
    """save_results.py:65-68 / compute_score.py:100-104: 16-bit BGR-ordered PNGs.  (a) a file written by the reference's own
    call (`cv2.imwrite`, committed by oracle/make_golden_png.py) decodes to the array that was written; (b) write -> read is
    the identity; (c) every PNG filter type decodes (OpenCV picks them adaptively); (d) the saved-results criterion and the
    float form `value / 2 ** 14` of the `using_saved_results` branch."""
    import struct
    import zlib
    import numpy as np
    from oracle.make_golden_png import golden_image
    from deep_rawburst_sr_b200.evaluation.synburst import save_results as SR
    img = golden_image()
    assert np.array_equal(SR.read_png16(os.path.join(golden_dir, 'pred_u16_cv2.png')), img)
    p = str(tmp_path / '0000.png')
    SR.write_png16(p, img)
    assert np.array_equal(SR.read_png16(p), img)
    with pytest.raises(ValueError):
        SR.write_png16(p, img.astype(np.uint8))
    # hand-encode the same image once per filter type (encoder side of the PNG spec, section 9)
    h, w, _ = img.shape
    rgb = np.ascontiguousarray(img[:, :, ::-1]).astype('>u2').view(np.uint8).reshape(h, w * 6).astype(np.int32)
    left = np.zeros_like(rgb); left[:, 6:] = rgb[:, :-6]
    up = np.zeros_like(rgb); up[1:] = rgb[:-1]
    ul = np.zeros_like(rgb); ul[1:, 6:] = rgb[:-1, :-6]
    pa, pb, pc = np.abs(up - ul), np.abs(left - ul), np.abs(left + up - 2 * ul)
    paeth = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
    for ft, pred in ((0, 0), (1, left), (2, up), (3, (left + up) >> 1), (4, paeth)):
        rows = np.concatenate([np.full((h, 1), ft, np.uint8), ((rgb - pred) & 255).astype(np.uint8)], axis=1)
        blob = SR._PNG_SIG + SR._chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 16, 2, 0, 0, 0)) + \
            SR._chunk(b'IDAT', zlib.compress(rows.tobytes())) + SR._chunk(b'IEND', b'')
        q = str(tmp_path / f'f{ft}.png')
        open(q, 'wb').write(blob)
        assert np.array_equal(SR.read_png16(q), img), ft
    t = SR.load_prediction(p)
    assert tuple(t.shape) == (1, 3, h, w) and t.dtype == torch.float32
    assert torch.equal(t[0], torch.from_numpy(img.astype(np.float32) / 2 ** 14).permute(2, 0, 1))
    q14 = torch.from_numpy(img.astype(np.int16)).permute(2, 0, 1).contiguous()
    assert np.array_equal(SR.prediction_to_array(q14), img)
    only = tmp_path / 'only'
    only.mkdir()
    SR.write_png16(str(only / '0000.png'), img)
    assert SR.saved_results_complete(str(only), [0]) and not SR.saved_results_complete(str(only), [0, 1])


def test_png16_four_channel_frames_and_val_set_layout(tmp_path, golden_dir):
    """dataset/synthetic_burst_val_set.py:42-55: burst frames are 16-bit FOUR-channel PNGs written / read through OpenCV (array
    channels B, G, R, A <-> file planes R, G, B, A).  (a) a 4-channel file written by `cv2.imwrite` (oracle/make_golden_png.py)
    decodes to the array that was written, and write -> read is the identity; (b) a miniature validation set written in the
    reference's directory layout reads back through `SyntheticBurstVal` with the reference's item contract."""
    import numpy as np
    from oracle.make_golden_png import golden_image4
    from deep_rawburst_sr_b200.evaluation.synburst import save_results as SR
    from deep_rawburst_sr_b200.dataset.synthetic_burst_val_set import SyntheticBurstVal, write_burst
    img = golden_image4()
    assert np.array_equal(SR.read_png16(os.path.join(golden_dir, 'burst_u16x4_cv2.png')), img)
    p = str(tmp_path / 'f.png')
    SR.write_png16(p, img)
    assert np.array_equal(SR.read_png16(p), img)
    g = torch.Generator().manual_seed(5)
    root = str(tmp_path / 'val')
    items = []
    for i in range(2):
        burst = (torch.rand(3, 4, 12, 16, generator=g) * 2 ** 14).round() / 2 ** 14       # 14-bit values: stored exactly
        gt = (torch.rand(3, 96, 128, generator=g) * 2 ** 14).round() / 2 ** 14
        meta = {'rgb2cam': torch.eye(3), 'gamma': True, 'smoothstep': True}
        write_burst(root, i, burst, gt, meta)
        items.append((burst, gt))
    ds = SyntheticBurstVal(root=root, num_bursts=2, burst_size=3)
    assert len(ds) == 2 and len(SyntheticBurstVal(root=root)) == 300 and SyntheticBurstVal(root=root).burst_size == 14
    for i, (burst, gt) in enumerate(items):
        b, t, meta = ds[i]
        assert torch.equal(b, burst) and torch.equal(t, gt) and b.dtype == torch.float32
        assert meta['burst_name'] == '%04d' % i and meta['gamma'] is True and torch.equal(meta['rgb2cam'], torch.eye(3))
    assert sorted(os.listdir(os.path.join(root, 'bursts', '0001'))) == ['im_raw_00.png', 'im_raw_01.png', 'im_raw_02.png']
    assert sorted(os.listdir(os.path.join(root, 'gt', '0001'))) == ['im_rgb.png', 'meta_info.pkl']


def test_network_param_and_experiment_registry():
    """evaluation/common_utils/network_param.py:64-111 and evaluation/synburst/experiments/dbsr_default.py: constructor
    constraints, generated names, the shipped experiment"""
    from deep_rawburst_sr_b200.evaluation.common_utils.network_param import NetworkParam
    from deep_rawburst_sr_b200.evaluation.synburst.compute_score import load_experiment
    n = NetworkParam(module='dbsr', parameter='default_synthetic')
    assert n.get_unique_name() == 'dbsr_default_synthetic' and n.get_display_name() == 'dbsr_default_synthetic'
    assert NetworkParam(module='dbsr', parameter='x', epoch=7).get_unique_name() == 'dbsr_x_ep0007'
    assert NetworkParam(module='dbsr', parameter='x', epoch=7, burst_sz=4).get_unique_name() == 'dbsr_x_ep0007_bsz04'
    assert NetworkParam(module='dbsr', parameter='x', burst_sz=12, display_name='short').get_display_name() == 'short'
    assert NetworkParam(unique_name='DBSR_results').get_unique_name() == 'DBSR_results'
    with pytest.raises(AssertionError):
        NetworkParam(network_path='a.pth')                                  # a downloaded checkpoint needs a unique_name
    with pytest.raises(AssertionError):
        NetworkParam(network_path='a.pth', unique_name='A', module='dbsr')   # ... and excludes module / parameter / epoch
    nets = load_experiment('dbsr_default')
    assert len(nets) == 1 and nets[0].network_path == 'dbsr_synthetic_default.pth' and nets[0].get_unique_name() == 'DBSR_syn'


def test_checkpoint_loading_incl_reference_format(tmp_path, monkeypatch, capsys):
    """admin/loading.py:24-100, utils/loading.py:6-19: checkpoint selection (file / latest of a directory / epoch), constructor
    keyword overrides, and a checkpoint whose pickle names the REFERENCE's modules (`admin.model_constructor.NetConstructor`,
    factory module `models.dbsr.dbsrnet`) -- what the published `dbsr_synthetic_default.pth` contains -- rebuilt with this
    package's classes"""
    import sys
    import types
    from deep_rawburst_sr_b200.admin import loading
    from deep_rawburst_sr_b200.utils.loading import load_network
    from deep_rawburst_sr_b200.evaluation.common_utils.network_param import NetworkParam
    from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic, DBSRNet
    torch.manual_seed(3)
    net = dbsrnet_default_synthetic()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    ws = tmp_path / 'ws'
    d = ws / 'checkpoints' / 'dbsr' / 'default_synthetic'
    d.mkdir(parents=True)
    for ep in (3, 5):
        torch.save({'epoch': ep, 'net': {k: v + ep for k, v in sd.items()}, 'constructor': net.constructor, 'net_info': {'e': ep}},
                   str(d / ('DBSRNet_ep%04d.pth.tar' % ep)))
    monkeypatch.setenv('DBSR_WORKSPACE_DIR', str(ws))
    key = next(iter(sd))
    latest = NetworkParam(module='dbsr', parameter='default_synthetic').load_net()
    assert isinstance(latest, DBSRNet) and torch.equal(latest.state_dict()[key], sd[key] + 5) and latest.info == {'e': 5}
    ep3 = NetworkParam(module='dbsr', parameter='default_synthetic', epoch=3).load_net()
    assert torch.equal(ep3.state_dict()[key], sd[key] + 3)
    with pytest.raises(Exception, match='No matching checkpoint'):
        NetworkParam(module='dbsr', parameter='default_synthetic', epoch=4).load_net()
    with pytest.raises(Exception, match='No matching checkpoint'):
        load_network('dbsr/nothing_here')
    # keyword overrides of constructor arguments: known ones replace the saved value, unknown ones are reported
    n2, ck = load_network(str(d / 'DBSRNet_ep0003.pth.tar'), return_dict=True, not_an_argument=1)
    assert ck['epoch'] == 3 and 'not_an_argument' in capsys.readouterr().out
    # a checkpoint pickled inside the reference's module layout
    ref_admin, ref_mc = types.ModuleType('admin'), types.ModuleType('admin.model_constructor')

    class NetConstructor:
        def __init__(self, fun_name, fun_module, args, kwds):
            self.fun_name, self.fun_module, self.args, self.kwds = fun_name, fun_module, args, kwds
    NetConstructor.__module__ = 'admin.model_constructor'
    NetConstructor.__qualname__ = 'NetConstructor'
    ref_mc.NetConstructor = NetConstructor
    ref_admin.model_constructor = ref_mc
    monkeypatch.setitem(sys.modules, 'admin', ref_admin)
    monkeypatch.setitem(sys.modules, 'admin.model_constructor', ref_mc)
    c = net.constructor
    ref_ck = tmp_path / 'dbsr_synthetic_default.pth'
    torch.save({'net': sd, 'constructor': NetConstructor(c.fun_name, 'models.dbsr.dbsrnet', c.args, dict(c.kwds)), 'net_info': None},
               str(ref_ck))
    monkeypatch.delitem(sys.modules, 'admin')
    monkeypatch.delitem(sys.modules, 'admin.model_constructor')
    monkeypatch.setenv('DBSR_PRETRAINED_NETS_DIR', str(tmp_path))
    got = NetworkParam(network_path='dbsr_synthetic_default.pth', unique_name='DBSR_syn').load_net()
    assert isinstance(got, DBSRNet) and type(got.constructor).__module__ == 'deep_rawburst_sr_b200.admin.model_constructor'
    assert got.constructor.fun_module == 'deep_rawburst_sr_b200.models.dbsr.dbsrnet'
    assert all(torch.equal(v, sd[k]) for k, v in got.state_dict().items())
    assert loading.package_module('numpy.core') == 'numpy.core' and loading.package_module('models.x') == 'deep_rawburst_sr_b200.models.x'


def test_flow_head_and_upfeat_as_one_tap_gemm_algebra():
    """The algebra behind `engine._pack_pwc`'s merged weight (DESIGN.md 4.1 "flow heads as tap planes"): one 1x1 contraction of
    the decoder's concat features with the rows [netUpfeat planes (ky*4+kx)*2+oc | netSix planes (ky*3+kx)*2+oc], followed by
    (a) the 9-tap sum `flow_from_taps` implements = Conv2d(K, 2, 3, 1, 1) (pwcnet.py:150), and (b) the <= 4-tap scatter
    `deconv_col2im` implements = ConvTranspose2d(K, 2, 4, 2, 1) (pwcnet.py:120).  Plain torch on the CPU, fp64."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(9)
    n, K, h, w = 2, 37, 5, 6
    x = torch.randn(n, K, h, w, generator=g, dtype=torch.float64)
    w6 = torch.randn(2, K, 3, 3, generator=g, dtype=torch.float64)
    b6 = torch.randn(2, generator=g, dtype=torch.float64)
    wu = torch.randn(K, 2, 4, 4, generator=g, dtype=torch.float64)           # ConvTranspose2d weight: [in, out, kh, kw]
    bu = torch.randn(2, generator=g, dtype=torch.float64)
    rows = torch.cat([wu.permute(2, 3, 1, 0).reshape(32, K), w6.permute(2, 3, 0, 1).reshape(18, K)], 0)   # as the engine packs them
    planes = torch.einsum('nkhw,rk->nrhw', x, rows)                           # the one 1x1 GEMM, 50 output planes
    up_planes, flow_planes = planes[:, :32], planes[:, 32:]
    # (a) flow head
    flow = b6.view(1, 2, 1, 1).expand(n, 2, h, w).clone()
    pad = F.pad(flow_planes, (1, 1, 1, 1))
    for ky in range(3):
        for kx in range(3):
            t = (ky * 3 + kx) * 2
            flow += pad[:, t:t + 2, ky:ky + h, kx:kx + w]
    assert torch.allclose(flow, F.conv2d(x, w6, b6, padding=1), atol=1e-10)
    # (b) transposed conv: out[oy, ox] = b + sum over (ky, kx) with iy = (oy + 1 - ky) / 2, ix = (ox + 1 - kx) / 2 integral and inside
    up = bu.view(1, 2, 1, 1).expand(n, 2, 2 * h, 2 * w).clone()
    for oy in range(2 * h):
        for ox in range(2 * w):
            for ky in range(4):
                for kx in range(4):
                    iy2, ix2 = oy + 1 - ky, ox + 1 - kx
                    if iy2 % 2 or ix2 % 2 or not (0 <= iy2 // 2 < h) or not (0 <= ix2 // 2 < w):
                        continue
                    t = (ky * 4 + kx) * 2
                    up[:, :, oy, ox] += up_planes[:, t:t + 2, iy2 // 2, ix2 // 2]
    assert torch.allclose(up, F.conv_transpose2d(x, wu, bu, stride=2, padding=1), atol=1e-10)


def test_batched_transform_sampling_is_bit_identical_to_the_per_burst_path():
    """rgb2rawburst_batch's host front end: the same `random` stream gives the same transform parameters, and the stacked
    matrix products / inversions equal the per-frame functions bit for bit"""
    import random
    import numpy as np
    from deep_rawburst_sr_b200.data import synthetic_burst_generation as G
    tp = {'max_translation': 24.0, 'max_rotation': 1.0, 'max_shear': 0.02, 'max_scale': 0.05, 'max_ar_factor': 0.03, 'border_crop': 24}
    random.seed(123)
    per = [G.sample_transforms((432, 400), 14, 4, tp) for _ in range(3)]
    random.seed(123)
    params = []
    for _ in range(3):
        params += G.sample_transform_params(14, 4, tp)
    bat = G.get_tmat_batch((432, 400), params)
    flat = np.stack([m for burst in per for m in burst])
    assert bat.shape == (42, 2, 3) and np.array_equal(bat, flat)
    assert np.array_equal(G._inverse_map_batch(bat), np.stack([G._inverse_map(m) for m in flat]))
    # identity case (max_translation <= 0.01 -> fixed shift) draws the same number of values
    random.seed(5)
    a = G.sample_transforms((64, 64), 3, 2, {'max_rotation': 2.0})
    random.seed(5)
    b = G.get_tmat_batch((64, 64), G.sample_transform_params(3, 2, {'max_rotation': 2.0}))
    assert np.array_equal(np.stack(a), b)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the product arm) prints ONE JSON line with the contract
    keys, the same metric / unit / workload string as the product arm, `impl: reference`, a `cpu_baseline` describing itself and an
    `e2e` equal to its own value with zero copy bytes; it runs the asked number of steps (no silent cap)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1',
                          '--size', '16'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split('\n') if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'impl'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'bursts/s' and d['higher_is_better'] is True and d['steps'] == 2
    assert d['metric'] == 'bursts/sec (14-frame, 4x SR)' and 'workload' in d['config'] and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'bursts/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert abs(d['value'] - 1e3 / d['ms_per_step']) < 1e-6 * d['value']
