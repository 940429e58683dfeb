"""The experiment the reference ships (evaluation/burstsr/experiments/dbsr_default.py): the published real-data network,
predictions saved under `DBSR_burstsr`."""
from ...common_utils.network_param import NetworkParam


def main():
    return [NetworkParam(network_path='dbsr_burstsr_default.pth', unique_name='DBSR_burstsr')]
