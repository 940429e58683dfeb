"""`SyntheticBurstVal`: the SyntheticBurst validation set on disk, with the item contract and the file layout of the
reference's dataset/synthetic_burst_val_set.py:22-79 --

    <root>/bursts/<index:04d>/im_raw_<frame:02d>.png   16-bit, 4 channels: the packed RGGB frame * 2**14
    <root>/gt/<index:04d>/im_rgb.png                   16-bit RGB:        the linear ground truth * 2**14
    <root>/gt/<index:04d>/meta_info.pkl                pickle:            camera meta data of the burst

-- read with this package's PNG codec (`evaluation/synburst/save_results.read_png16`; the reference uses
`cv2.imread(..., cv2.IMREAD_UNCHANGED)`, whose channel order the codec reproduces) instead of OpenCV.  `write_burst` is the
inverse (the tests build a miniature set with it)."""
import os
import pickle as pkl

import numpy as np
import torch

from ..admin.environment import env_settings
from ..evaluation.synburst.save_results import read_png16, write_png16


class SyntheticBurstVal(torch.utils.data.Dataset):
    def __init__(self, root=None, initialize=True, num_bursts: int = 300, burst_size: int = 14):
        """root: dataset directory (default `env_settings().synburstval_dir`).  The published set has 300 bursts of 14 frames;
        `num_bursts` / `burst_size` exist for subsets."""
        self.root = env_settings().synburstval_dir if root is None else root
        self.burst_list = list(range(num_bursts))
        self.burst_size = burst_size

    def initialize(self):
        pass

    def __len__(self):
        return len(self.burst_list)

    def _read_burst_image(self, index, image_id):
        im = read_png16('{}/bursts/{:04d}/im_raw_{:02d}.png'.format(self.root, index, image_id))
        return torch.from_numpy(im.astype(np.float32)).permute(2, 0, 1).float() / (2 ** 14)

    def _read_gt_image(self, index):
        gt = read_png16('{}/gt/{:04d}/im_rgb.png'.format(self.root, index))
        return (torch.from_numpy(gt.astype(np.float32)) / 2 ** 14).permute(2, 0, 1).float()

    def _read_meta_info(self, index):
        with open('{}/gt/{:04d}/meta_info.pkl'.format(self.root, index), 'rb') as f:
            return pkl.load(f)

    def __getitem__(self, index):
        """-> burst [burst_size, 4, 48, 48] (R, G, G, B planes), gt [3, 384, 384] (linear), meta_info (+ 'burst_name')"""
        burst = torch.stack([self._read_burst_image(index, i) for i in range(self.burst_size)], 0)
        gt = self._read_gt_image(index)
        meta_info = self._read_meta_info(index)
        meta_info['burst_name'] = '{:04d}'.format(index)
        return burst, gt, meta_info


def write_burst(root: str, index: int, burst: torch.Tensor, gt: torch.Tensor, meta_info: dict) -> None:
    """store one item in the layout above (values are quantised to 14 bits like the published files)"""
    os.makedirs('{}/bursts/{:04d}'.format(root, index), exist_ok=True)
    os.makedirs('{}/gt/{:04d}'.format(root, index), exist_ok=True)
    q = lambda t: (t.clamp(0.0, 1.0) * 2 ** 14).round().to(torch.int32).permute(1, 2, 0).contiguous().numpy().astype(np.uint16)
    for i in range(burst.shape[0]):
        write_png16('{}/bursts/{:04d}/im_raw_{:02d}.png'.format(root, index, i), q(burst[i]))
    write_png16('{}/gt/{:04d}/im_rgb.png'.format(root, index), q(gt))
    with open('{}/gt/{:04d}/meta_info.pkl'.format(root, index), 'wb') as f:
        pkl.dump({k: v for k, v in meta_info.items() if k != 'burst_name'}, f)
