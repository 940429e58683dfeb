"""The experiment the reference ships (evaluation/synburst/experiments/dbsr_default.py): the published synthetic-data
network, predictions saved under `DBSR_syn`."""
from ...common_utils.network_param import NetworkParam


def main():
    return [NetworkParam(network_path='dbsr_synthetic_default.pth', unique_name='DBSR_syn')]
