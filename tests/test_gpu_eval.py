"""GPU: the batched scoring loop of the SyntheticBurst protocol (evaluation/synburst/compute_score.py)
against the same protocol walked image by image on the CPU oracle: forward -> 14-bit quantisation -> PSNR / SSIM with
boundary_ignore -> mean over the set.  fp32 path: PSNR within 1e-3 dB, SSIM within 1e-5 (a prediction within 1.5e-7 of the
oracle's can flip single 14-bit codes); bf16 tensor-core path: PSNR within 0.02 dB (north_star), SSIM within 1e-3."""
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
