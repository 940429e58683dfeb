"""`NetworkParam`: one network instance of an evaluation run, with the constructor arguments, naming rules and checkpoint
lookup of the reference's evaluation/common_utils/network_param.py:20-111 (an experiment file returns a list of these)."""
import os

from ...admin.environment import env_settings
from ...utils.loading import load_network


class NetworkParam:
    def __init__(self, module=None, parameter=None, epoch=None, burst_sz=None, display_name=None, unique_name=None,
                 network_path=None):
        """module / parameter: training module and setting the network was trained with (checkpoints under
        `<workspace_dir>/checkpoints/<module>/<parameter>`), epoch: which checkpoint (None: latest);
        network_path: a downloaded checkpoint instead -- absolute, or a file name inside `pretrained_nets_dir` -- which
        excludes module / parameter / epoch and requires `unique_name`;
        burst_sz: frames per burst used for the evaluation (None: the dataset's); display_name: name in reports;
        unique_name: name of the directory the predictions are saved under.  A NetworkParam with only `unique_name` stands for
        downloaded predictions in `<save_data_path>/<dataset>/<unique_name>`."""
        assert network_path is None or (module is None and parameter is None and epoch is None)
        assert network_path is None or (unique_name is not None)
        self.module, self.parameter, self.epoch = module, parameter, epoch
        self.display_name, self.unique_name = display_name, unique_name
        self.burst_sz = burst_sz
        self.network_path = network_path

    def load_net(self):
        if self.network_path is not None:
            path = self.network_path
            if not os.path.isabs(path):
                path = '{}/{}'.format(env_settings().pretrained_nets_dir, path)
            net, _ = load_network(path, return_dict=True)
        elif self.epoch is None:
            net, _ = load_network('{}/{}'.format(self.module, self.parameter), return_dict=True)
        else:
            net, _ = load_network('{}/{}'.format(self.module, self.parameter), checkpoint=self.epoch, return_dict=True)
        return net

    def get_display_name(self):
        return self.display_name if self.display_name is not None else self.get_unique_name()

    def get_unique_name(self):
        if self.unique_name is not None:
            return self.unique_name
        name = '{}_{}'.format(self.module, self.parameter)
        if self.epoch is not None:
            name = '{}_ep{:04d}'.format(name, self.epoch)
        if self.burst_sz is not None:
            name = '{}_bsz{:02d}'.format(name, self.burst_sz)
        return name
