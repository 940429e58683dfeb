"""Result writer / reader of the SyntheticBurst protocol.

The reference's `save_results` (evaluation/synburst/save_results.py:33-68) runs the network one burst at a time and stores
`(pred.squeeze(0).permute(1, 2, 0).clamp(0, 1) * 2 ** 14).cpu().numpy().astype(np.uint16)` with `cv2.imwrite(<name>.png)`
(:65-68): a 16-bit, 3-channel PNG whose array channel 0 lands in the file's BLUE plane (OpenCV arrays are BGR).  Its
`compute_score(load_saved=True)` reads the files back with `cv2.imread(..., cv2.IMREAD_UNCHANGED)` and divides by 2 ** 14
(compute_score.py:100-104).  This module keeps that file format bit for bit at the pixel level -- files written here load in
the reference's reader and vice versa -- and changes the schedule:

  * the network writes the 14-bit int16 form straight from the predictor epilogue (`net.output_int16`), a whole batch per
    forward; the device -> host copy of batch i overlaps forward i + 1 (`HostPipeline`), and the PNG encoding (zlib, which
    releases the GIL) runs on a thread pool behind both;
  * the PNG codec is written out here (stdlib `zlib` + numpy; OpenCV is an un-vendored, unpinned dependency of the
    reference and is not needed at run time).  `read_png16` decodes every filter type of non-interlaced 16-bit RGB files, so
    it also reads what `cv2.imwrite` produced.
"""
from __future__ import annotations

import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np
import torch

_PNG_SIG = b'\x89PNG\r\n\x1a\n'


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack('>I', len(data)) + tag + data + struct.pack('>I', zlib.crc32(tag + data) & 0xFFFFFFFF)


def _file_order(channels: int):
    """OpenCV arrays are B, G, R (, A); PNG stores R, G, B (, A)"""
    return [2, 1, 0] if channels == 3 else [2, 1, 0, 3]


def write_png16(path: str, image: np.ndarray, level: int = 3) -> None:
    """image: uint16 [H, W, 3] or [H, W, 4] in the array order the reference hands to `cv2.imwrite` (channel 0 -> file BLUE
    plane; a fourth channel is the file's alpha plane -- the packed RGGB burst frames of the SyntheticBurst validation set are
    stored that way, dataset/synthetic_burst_val_set.py:44).  16-bit samples are stored big-endian; every scanline uses the `Up`
    filter (vectorised; smooth rows compress well)."""
    a = np.ascontiguousarray(image)
    if a.dtype != np.uint16 or a.ndim != 3 or a.shape[2] not in (3, 4):
        raise ValueError(f'write_png16 takes a uint16 [H, W, 3 or 4] array, got {a.dtype} {a.shape}')
    h, w, ch = a.shape
    rgb = np.ascontiguousarray(a[:, :, _file_order(ch)]).astype('>u2').view(np.uint8).reshape(h, w * 2 * ch)   # big-endian
    up = rgb.copy()
    up[1:] = rgb[1:] - rgb[:-1]                                                 # filter type 2 (Up), modulo 256
    rows = np.empty((h, 1 + w * 2 * ch), dtype=np.uint8)
    rows[:, 0] = 2
    rows[0, 0] = 0                                                              # first row: no filter (nothing above it)
    rows[:, 1:] = up
    rows[0, 1:] = rgb[0]
    ihdr = struct.pack('>IIBBBBB', w, h, 16, 2 if ch == 3 else 6, 0, 0, 0)      # 16 bit, colour type 2 (RGB) / 6 (RGBA), no interlace
    blob = _PNG_SIG + _chunk(b'IHDR', ihdr) + _chunk(b'IDAT', zlib.compress(rows.tobytes(), level)) + _chunk(b'IEND', b'')
    tmp = path + '.tmp'
    with open(tmp, 'wb') as f:
        f.write(blob)
    os.replace(tmp, path)                                                       # a reader never sees a half-written file


def _unfilter(raw: np.ndarray, h: int, stride: int, bpp: int) -> np.ndarray:
    """PNG scanline reconstruction (filters 0..4) of `h` rows of `stride` bytes, `bpp` bytes per pixel"""
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    for y in range(h):
        ft = int(raw[y, 0])
        line = raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:          # Sub: bytes of the same channel form a running sum along the row
            cur = (np.cumsum(line.reshape(-1, bpp), axis=0) & 255).reshape(-1)
        elif ft in (3, 4):     # Average / Paeth depend on the reconstructed left pixel: sequential in x, vector over bpp
            cur = np.zeros(stride, dtype=np.int32)
            left = np.zeros(bpp, dtype=np.int32)
            upleft = np.zeros(bpp, dtype=np.int32)
            for x in range(0, stride, bpp):
                up = prev[x:x + bpp]
                if ft == 3:
                    pred = (left + up) >> 1
                else:
                    p = left + up - upleft
                    pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - upleft)
                    pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, upleft))
                left = (line[x:x + bpp] + pred) & 255
                cur[x:x + bpp] = left
                upleft = up
        else:
            raise ValueError(f'bad PNG filter type {ft}')
        out[y] = cur
        prev = cur
    return out


def read_png16(path: str) -> np.ndarray:
    """-> uint16 [H, W, 3] (or [H, W, 4] for an RGBA file) in `cv2.imread(path, cv2.IMREAD_UNCHANGED)` order (channel 0 = file
    BLUE plane, channel 3 = alpha).  Reads 16-bit, non-interlaced RGB / RGBA PNGs (what `write_png16` and `cv2.imwrite` of a
    uint16 HxWx3 / HxWx4 array produce)."""
    with open(path, 'rb') as f:
        blob = f.read()
    if blob[:8] != _PNG_SIG:
        raise ValueError(f'{path}: not a PNG file')
    pos, idat, hdr = 8, [], None
    while pos < len(blob):
        n, tag = struct.unpack('>I4s', blob[pos:pos + 8])
        data = blob[pos + 8:pos + 8 + n]
        if tag == b'IHDR':
            hdr = struct.unpack('>IIBBBBB', data)
        elif tag == b'IDAT':
            idat.append(data)
        elif tag == b'IEND':
            break
        pos += 12 + n
    if hdr is None:
        raise ValueError(f'{path}: no IHDR chunk')
    w, h, depth, ctype, _comp, _filt, interlace = hdr
    if depth != 16 or ctype not in (2, 6) or interlace != 0:
        raise ValueError(f'{path}: expected a 16-bit non-interlaced RGB(A) PNG (got depth {depth}, colour type {ctype}, interlace {interlace})')
    ch = 3 if ctype == 2 else 4
    stride = w * 2 * ch
    raw = np.frombuffer(zlib.decompress(b''.join(idat)), dtype=np.uint8).reshape(h, 1 + stride)
    rgb = np.ascontiguousarray(_unfilter(raw, h, stride, 2 * ch)).view('>u2').reshape(h, w, ch)
    return np.ascontiguousarray(rgb[:, :, _file_order(ch)]).astype(np.uint16)


def prediction_to_array(pred_q: torch.Tensor) -> np.ndarray:
    """int16 [3, H, W] (14-bit quantised prediction, host tensor) -> uint16 [H, W, 3], the array of save_results.py:65-66"""
    assert pred_q.dtype == torch.int16 and pred_q.dim() == 3 and not pred_q.is_cuda
    return pred_q.permute(1, 2, 0).contiguous().numpy().astype(np.uint16)


def load_prediction(path: str, device=None) -> torch.Tensor:
    """saved PNG -> float prediction [1, 3, H, W] = value / 2 ** 14 (compute_score.py:100-104, `using_saved_results`)"""
    t = (torch.from_numpy(read_png16(path).astype(np.float32)) / 2 ** 14).permute(2, 0, 1).float().unsqueeze(0)
    return t.to(device) if device is not None else t


def saved_results_complete(out_dir: str, dataset) -> bool:
    """the reference's criterion for `using_saved_results` (compute_score.py:78-88): as many .png files as bursts"""
    return os.path.isdir(out_dir) and len([r for r in os.listdir(out_dir) if r[-3:] == 'png']) == len(dataset)


def save_results_for_setting(setting_name: str, dataset=None, batch_size: int = 32, device='cuda') -> dict:
    """The reference's `save_results(setting_name)` (evaluation/synburst/save_results.py:33-69): for every network of the
    experiment, write the predictions of the whole validation set to `<save_data_path>/synburst/<unique_name>/<burst_name>.png`.
    -> {unique name: files written}"""
    from ...admin.environment import env_settings
    from .compute_score import load_experiment
    if dataset is None:
        from ...dataset.synthetic_burst_val_set import SyntheticBurstVal
        dataset = SyntheticBurstVal()
    base_results_dir = env_settings().save_data_path
    written = {}
    for n in load_experiment(setting_name, 'synburst'):
        net = n.load_net()
        net.to(device).train(False)
        out_dir = '{}/synburst/{}'.format(base_results_dir, n.get_unique_name())
        written[n.get_unique_name()] = save_results(net, dataset, out_dir, batch_size=batch_size, device=device, burst_sz=n.burst_sz)
    return written


@torch.no_grad()
def save_results(net, dataset=None, out_dir: Optional[str] = None, batch_size: int = 32, device='cuda', burst_sz: Optional[int] = None,
                 workers: int = 8):
    """Run `net` over `dataset` (items `(burst [N, 4, H, W], gt, meta_info)`, the contract of `SyntheticBurstVal`) and write one
    `<burst_name>.png` per burst into `out_dir` -- the files of reference save_results.py:52-68.  Returns the number written.
    Called with an experiment NAME as the only argument it is the reference's `save_results(setting_name)` driver
    (`save_results_for_setting`)."""
    if isinstance(net, str):
        return save_results_for_setting(net, dataset, batch_size=batch_size, device=device)
    assert dataset is not None and out_dir is not None
    from ...pipeline import HostPipeline
    os.makedirs(out_dir, exist_ok=True)
    device = torch.device(device)
    was_q = getattr(net, 'output_int16', False)
    net.output_int16 = True
    pipe = HostPipeline(net, depth=2, device=device)
    pool = ThreadPoolExecutor(max_workers=workers)
    jobs, pending = [], []          # pending: (event, host_out, names) of submitted batches, at most depth of them
    slots = {}

    def flush(entry):
        ev, host_out, names = entry
        ev.synchronize()
        for i, name in enumerate(names):
            jobs.append(pool.submit(write_png16, os.path.join(out_dir, name + '.png'), prediction_to_array(host_out[i].clone())))

    try:
        r = getattr(getattr(getattr(net, 'decoder', None), 'upsample_layer', None), 'upsample_factor', 8)
        for bi, start in enumerate(range(0, len(dataset), batch_size)):
            items = [dataset[i] for i in range(start, min(start + batch_size, len(dataset)))]
            # tuples (burst, gt, meta_info) of SyntheticBurstVal, or the dicts of the BurstSR sampler
            raw = [it['burst'] if isinstance(it, dict) else it[0] for it in items]
            names = [it['burst_name'] if isinstance(it, dict) else it[2]['burst_name'] for it in items]
            bursts = torch.stack([b[:burst_sz] if burst_sz is not None else b for b in raw]).float()
            b, _n, _c, h, w = bursts.shape
            if len(pending) >= 2:
                flush(pending.pop(0))            # batch bi - 2 used the slot that is reused now: drain it first
            key = bi % 2
            if key not in slots or slots[key][0].shape != bursts.shape:
                slots[key] = (torch.empty(bursts.shape).pin_memory(), torch.empty((b, 3, r * h, r * w), dtype=torch.int16).pin_memory())
            host_in, host_out = slots[key]
            host_in.copy_(bursts)
            pending.append((pipe.submit(host_in, host_out), host_out, names))
        while pending:
            flush(pending.pop(0))
        for j in jobs:
            j.result()
    finally:
        pipe.drain()
        pool.shutdown(wait=True)
        net.output_int16 = was_q
    return len(jobs)


if __name__ == '__main__':
    import argparse
    parser = argparse.ArgumentParser(description='Save network predictions on the SyntheticBurst validation set')
    parser.add_argument('setting', type=str, help='Name of experiment setting')
    save_results_for_setting(parser.parse_args().setting)
