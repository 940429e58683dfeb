"""Result writer of the BurstSR protocol (reference evaluation/burstsr/save_results.py:33-69): the same 16-bit PNG files as
the SyntheticBurst writer, under `<save_data_path>/burstsr/<unique_name>`; items are the dicts of the reference's
`IndexedBurst` sampler (`'burst'`, `'frame_gt'`, `'burst_name'`).  The BurstSR dataset classes themselves (phone / DSLR RAW
containers with EXIF pickles) are out of scope: pass `dataset=`."""
from ..synburst.save_results import save_results as _write_all


def save_results(setting_name, dataset, batch_size: int = 16, device='cuda') -> dict:
    """-> {unique name: files written}"""
    from ...admin.environment import env_settings
    from ..synburst.compute_score import load_experiment
    base_results_dir = env_settings().save_data_path
    written = {}
    for n in load_experiment(setting_name, 'burstsr'):
        net = n.load_net()
        net.to(device).train(False)
        out_dir = '{}/burstsr/{}'.format(base_results_dir, n.get_unique_name())
        written[n.get_unique_name()] = _write_all(net, dataset, out_dir, batch_size=batch_size, device=device, burst_sz=n.burst_sz)
    return written
