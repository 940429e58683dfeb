"""GPU: the batched scoring loop of the SyntheticBurst protocol (evaluation/synburst/compute_score.py)
against the same protocol walked image by image on the CPU oracle: forward -> 14-bit quantisation -> PSNR / SSIM with
boundary_ignore -> mean over the set.  fp32 path: PSNR within 1e-3 dB, SSIM within 1e-5 (a prediction within 1.5e-7 of the
oracle's can flip single 14-bit codes); bf16 tensor-core path: PSNR within 0.02 dB (north_star), SSIM within 1e-3."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dbsr_oracle as O  # noqa: E402
from oracle import metrics_oracle as M  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def _oracle_scores(bursts, gts, sd, bi):
    ps, ss = [], []
    for b, g in zip(bursts, gts):
        pred, _ = O.dbsr_forward(b.unsqueeze(0), sd)
        pred = (pred.clamp(0.0, 1.0) * 2 ** 14).short().float() / 2 ** 14          # compute_score.py:110-111
        ps.append(float(M.psnr(pred, g.unsqueeze(0), bi)))
        ss.append(float(M.ssim_metric(pred, g.unsqueeze(0), bi, use_for_loss=False)))
    return sum(ps) / len(ps), sum(ss) / len(ss)


def test_score_dataset_matches_per_image_protocol(dev):
    from deep_rawburst_sr_b200.evaluation.synburst.compute_score import TensorBurstSet, generate_formatted_report, score_dataset
    from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic
    sd = O.make_state_dict(0)
    n, bi = 5, 40
    bursts = O.make_burst(3, n, 3, 16, 16)
    gts = torch.rand(n, 3, 128, 128, generator=torch.Generator().manual_seed(77))
    ref_psnr, ref_ssim = _oracle_scores(bursts, gts, sd, bi)
    net = dbsrnet_default_synthetic()
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    data = TensorBurstSet(bursts, gts)
    for precision, tp, ts in (('fp32', 1e-3, 1e-5), ('bf16', 0.02, 1e-3)):
        net.set_precision(precision)
        for batch in (2, 8):                       # ragged last batch / one batch
            got = score_dataset(net, data, batch_size=batch, boundary_ignore=bi, device=dev)
            assert got['count'] == n and not net.output_int16
            assert abs(got['psnr'] - ref_psnr) <= tp, (precision, batch, got, ref_psnr)
            assert abs(got['ssim'] - ref_ssim) <= ts, (precision, batch, got, ref_ssim)
    # a dataset without the contiguous-batch shortcut goes through the pinned double-buffered stager: same scores
    items = [data[i] for i in range(n)]
    assert score_dataset(net, items, batch_size=2, boundary_ignore=bi, device=dev) == score_dataset(net, data, batch_size=2, boundary_ignore=bi, device=dev)
    assert 'psnr' in generate_formatted_report({'dbsr': got})
    with pytest.raises(NotImplementedError):
        score_dataset(net, data, metrics=('lpips',), device=dev)


class _LookupNet:
    """Stand-in for the SR network of the BurstSR loop test: returns a prepared prediction per burst (matched by the burst's
    content), honouring `output_int16` like DBSRNet.  The real network inside a scoring loop is covered by the SyntheticBurst
    test above; here the prediction must RESEMBLE the ground truth, or the colour-error mask invalidates nearly every pixel."""

    def __init__(self, bursts, preds):
        self.keys, self.preds, self.output_int16 = bursts.flatten(1)[:, :64].clone(), preds, False

    def __call__(self, burst):
        idx = torch.cdist(burst.flatten(1)[:, :64].cpu(), self.keys).argmin(dim=1)
        pred = self.preds[idx].to(burst.device)
        if self.output_int16:
            pred = (pred.clamp(0.0, 1.0) * 2 ** 14).short()
        return pred, {}


def test_burstsr_score_dataset_matches_batch1_loop(dev):
    """BurstSR protocol (evaluation/burstsr/compute_score.py:97-128): the batched loop (per-image normalisation inside the
    alignment, masked PSNR / SSIM from the fused kernels, one reduction) against the reference's schedule -- one burst at a
    time through the drop-in modules: forward, quantise, SpatialColorAlignment, PSNR / SSIM with `valid`, python mean."""
    from deep_rawburst_sr_b200.evaluation.burstsr.compute_score import score_dataset
    from deep_rawburst_sr_b200.models.alignment.pwcnet import PWCNet
    from deep_rawburst_sr_b200.models.loss.image_quality_v2 import PSNR, SSIM
    from deep_rawburst_sr_b200.models.loss.spatial_color_alignment import SpatialColorAlignment
    from oracle import sca_oracle as S
    n, bi = 5, 8
    preds, gts, bursts = S.make_sca_inputs(0, n, 192)
    scale = torch.linspace(0.5, 1.0, n).view(n, 1, 1, 1)           # different maxima: per-image normalisation matters
    preds, gts, bursts = preds * scale, gts * scale, bursts * scale.view(n, 1, 1, 1, 1)
    net = _LookupNet(bursts, preds)
    pwc = PWCNet(load_pretrained=False)
    pwc.load_state_dict(S.pwc_state_dict(0, 1.0), strict=True)
    pwc = pwc.to(dev).eval()
    data = [{'burst': b, 'frame_gt': g, 'burst_name': str(i)} for i, (b, g) in enumerate(zip(bursts, gts))]
    sca = SpatialColorAlignment(pwc, sr_factor=4)
    sca.to(dev)
    psnr_fn, ssim_fn = PSNR(boundary_ignore=bi), SSIM(boundary_ignore=bi, use_for_loss=False)
    ps, ss, vf = [], [], []
    for it in data:
        burst, gt = it['burst'].unsqueeze(0).to(dev), it['frame_gt'].unsqueeze(0).to(dev)
        pred, _ = net(burst)
        pred = (pred.clamp(0.0, 1.0) * 2 ** 14).short().float() / 2 ** 14
        pm, valid = sca(pred, gt, burst)
        vf.append(float(valid.float().mean()))
        ps.append(psnr_fn(pm, gt, valid=valid).cpu().item())
        ss.append(ssim_fn(pm, gt, valid=valid).cpu().item())
    assert 0.05 < sum(vf) / n < 1.0, vf                           # a real mask: some, not all, pixels valid (cf. the goldens: 6-21 %)
    for batch in (2, 8):
        got = score_dataset(net, data, pwc, boundary_ignore=bi, batch_size=batch, device=dev)
        assert got['count'] == n and not net.output_int16
        assert abs(got['psnr'] - sum(ps) / n) <= 1e-3, (got, ps)
        assert abs(got['ssim'] - sum(ss) / n) <= 1e-5, (got, ss)


def test_save_results_and_score_from_saved_pngs(dev, tmp_path):
    """save_results.py:52-68 + compute_score.py:78-104: the batched writer stores, per burst, the PNG the reference's batch-1
    loop would store (pixel for pixel: the int16 epilogue IS (pred.clamp(0, 1) * 2 ** 14)), and scoring from the saved files
    (`using_saved_results`) gives the same report as scoring the network directly."""
    import numpy as np
    from deep_rawburst_sr_b200.evaluation.synburst.compute_score import TensorBurstSet, score_dataset
    from deep_rawburst_sr_b200.evaluation.synburst.save_results import read_png16, save_results
    from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic
    from oracle import dbsr_oracle as O
    sd = O.make_state_dict(0, dbsr_gain=1.5)
    net = dbsrnet_default_synthetic()
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval().set_precision('bf16')
    g = torch.Generator().manual_seed(11)
    data = TensorBurstSet(torch.rand(5, 4, 4, 24, 24, generator=g), torch.rand(5, 3, 192, 192, generator=g))
    out_dir = str(tmp_path / 'synburst' / 'net0')
    for graph in (False, True):
        net.use_cuda_graph = graph
        assert save_results(net, data, out_dir, batch_size=2, device=dev) == 5
        for i in range(5):
            pred, _ = net(data.bursts[i:i + 1].to(dev))
            want = (pred.squeeze(0).permute(1, 2, 0).clamp(0.0, 1.0) * 2 ** 14).cpu().numpy().astype(np.uint16)
            assert np.array_equal(read_png16(os.path.join(out_dir, '{:04d}.png'.format(i))), want), (graph, i)
    net.use_cuda_graph = False
    live = score_dataset(net, data, batch_size=2, device=dev, boundary_ignore=8)
    saved = score_dataset(None, data, batch_size=2, device=dev, boundary_ignore=8, saved_dir=out_dir)
    assert saved['using_saved_results'] and not live['using_saved_results']
    assert abs(saved['psnr'] - live['psnr']) < 1e-5 and abs(saved['ssim'] - live['ssim']) < 1e-6


def test_experiment_drivers_on_a_miniature_validation_set(dev, tmp_path, monkeypatch, capsys):
    """The reference's two command-line drivers end to end (evaluation/synburst/save_results.py:33-69, compute_score.py:36-122):
    experiment file -> NetworkParam -> checkpoint (admin/loading.py) -> SyntheticBurstVal files on disk -> predictions under
    <save_data_path>/synburst/<unique_name> -> report; `load_saved` scores the saved files instead of running the network and
    must print the same numbers; `burst_sz` truncates the bursts; a NetworkParam with only a unique_name scores downloaded
    predictions."""
    import sys
    from deep_rawburst_sr_b200.dataset.synthetic_burst_val_set import SyntheticBurstVal, write_burst
    from deep_rawburst_sr_b200.evaluation.synburst.compute_score import compute_score, score_dataset
    from deep_rawburst_sr_b200.evaluation.synburst.save_results import save_results
    from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic
    sd = O.make_state_dict(0, dbsr_gain=1.5)
    net = dbsrnet_default_synthetic()
    net.load_state_dict(sd, strict=True)
    ck = tmp_path / 'nets' / 'tiny_net.pth'
    ck.parent.mkdir()
    torch.save({'net': net.state_dict(), 'constructor': net.constructor, 'net_info': None}, str(ck))
    g = torch.Generator().manual_seed(21)
    root = str(tmp_path / 'val')
    for i in range(3):
        write_burst(root, i, torch.rand(6, 4, 48, 48, generator=g), torch.rand(3, 384, 384, generator=g), {'gamma': True})
    data = SyntheticBurstVal(root=root, num_bursts=3, burst_size=6)
    exp = tmp_path / 'exp_pkg'
    exp.mkdir()
    (exp / '__init__.py').write_text('')
    (exp / 'tiny.py').write_text(
        'from deep_rawburst_sr_b200.evaluation.common_utils.network_param import NetworkParam\n'
        'def main():\n'
        "    return [NetworkParam(network_path='tiny_net.pth', unique_name='TINY'),\n"
        "            NetworkParam(network_path='tiny_net.pth', unique_name='TINY4', burst_sz=4, display_name='four frames')]\n")
    (exp / 'downloaded.py').write_text(
        'from deep_rawburst_sr_b200.evaluation.common_utils.network_param import NetworkParam\n'
        'def main():\n'
        "    return [NetworkParam(unique_name='TINY')]\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.setenv('DBSR_PRETRAINED_NETS_DIR', str(ck.parent))
    monkeypatch.setenv('DBSR_SAVE_DATA_PATH', str(tmp_path / 'results'))
    live = compute_score('exp_pkg.tiny', load_saved=True, dataset=data, batch_size=2, device=dev)   # nothing saved yet: runs the nets
    assert set(live) == {'TINY', 'four frames'} and 'four frames' in capsys.readouterr().out
    assert save_results('exp_pkg.tiny', data, batch_size=2, device=dev) == {'TINY': 3, 'TINY4': 3}
    assert sorted(os.listdir(str(tmp_path / 'results' / 'synburst' / 'TINY4'))) == ['0000.png', '0001.png', '0002.png']
    saved = compute_score('exp_pkg.tiny', load_saved=True, dataset=data, batch_size=2, device=dev)
    for name in live:
        for m in ('psnr', 'ssim'):
            assert abs(live[name][m] - saved[name][m]) < 1e-5, (name, m)
    assert abs(live['TINY']['psnr'] - live['four frames']['psnr']) > 1e-4                 # burst_sz = 4 is a different run
    down = compute_score('exp_pkg.downloaded', load_saved=True, dataset=data, batch_size=2, device=dev, verbose=False)
    assert abs(down['TINY']['psnr'] - live['TINY']['psnr']) < 1e-5
    direct = score_dataset(net.to(dev).eval(), data, batch_size=2, device=dev)
    assert abs(direct['psnr'] - live['TINY']['psnr']) < 1e-5 and abs(direct['ssim'] - live['TINY']['ssim']) < 1e-6


def test_burstsr_drivers_save_and_score_by_experiment_name(dev, tmp_path, monkeypatch):
    """evaluation/burstsr/{save_results,compute_score}.py by experiment name: predictions go to <save_data_path>/burstsr/<name>,
    `load_saved` scores the files (criterion: at least one PNG per burst) and agrees with the live run"""
    from deep_rawburst_sr_b200.evaluation.burstsr.compute_score import compute_score
    from deep_rawburst_sr_b200.evaluation.burstsr.save_results import save_results
    from deep_rawburst_sr_b200.evaluation.synburst.compute_score import load_experiment
    from deep_rawburst_sr_b200.models.alignment.pwcnet import PWCNet
    from deep_rawburst_sr_b200.models.dbsr.dbsrnet import dbsrnet_default_synthetic
    from oracle import sca_oracle as S
    assert load_experiment('dbsr_default', 'burstsr')[0].get_unique_name() == 'DBSR_burstsr'
    net = dbsrnet_default_synthetic()
    net.load_state_dict(O.make_state_dict(0, dbsr_gain=1.5), strict=True)
    ck = tmp_path / 'nets' / 'real_net.pth'
    ck.parent.mkdir()
    torch.save({'net': net.state_dict(), 'constructor': net.constructor, 'net_info': None}, str(ck))
    pwc = PWCNet(load_pretrained=False)
    pwc.load_state_dict(S.pwc_state_dict(0, 1.0), strict=True)
    pwc = pwc.to(dev).eval()
    g = torch.Generator().manual_seed(31)
    data = [{'burst': torch.rand(5, 4, 48, 48, generator=g), 'frame_gt': torch.rand(3, 384, 384, generator=g) * 0.5,
             'burst_name': 'seq%02d' % i} for i in range(3)]
    exp = tmp_path / 'exp_real'
    exp.mkdir()
    (exp / '__init__.py').write_text('')
    (exp / 'one.py').write_text(
        'from deep_rawburst_sr_b200.evaluation.common_utils.network_param import NetworkParam\n'
        'def main():\n'
        "    return [NetworkParam(network_path='real_net.pth', unique_name='REAL')]\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.setenv('DBSR_PRETRAINED_NETS_DIR', str(ck.parent))
    monkeypatch.setenv('DBSR_SAVE_DATA_PATH', str(tmp_path / 'results'))
    with pytest.raises(ValueError):
        compute_score('exp_real.one')                      # the BurstSR RAW dataset classes are not part of the package
    live = compute_score('exp_real.one', dataset=data, alignment_net=pwc, batch_size=2, device=dev, verbose=False)
    assert save_results('exp_real.one', data, batch_size=2, device=dev) == {'REAL': 3}
    assert sorted(os.listdir(str(tmp_path / 'results' / 'burstsr' / 'REAL'))) == ['seq00.png', 'seq01.png', 'seq02.png']
    saved = compute_score('exp_real.one', load_saved=True, dataset=data, alignment_net=pwc, batch_size=2, device=dev, verbose=False)
    assert abs(live['REAL']['psnr'] - saved['REAL']['psnr']) < 1e-4 and abs(live['REAL']['ssim'] - saved['REAL']['ssim']) < 1e-5
