"""Host-buffer front end of the burst forward pass: bursts arrive in (pinned) host memory, results go back to host.

The reference's callers do `net(burst.cuda())` / `pred.cpu()` around every forward (evaluation loops of
`evaluation/synburst`, `util_scripts/`), which serialises the PCIe copies with the compute.  `HostPipeline` keeps the
same per-burst semantics but overlaps the three phases of consecutive submissions on three CUDA streams:

    copy-in stream :  H2D(i+1)                      (host_in  -> device input slot)
    compute stream :  forward(i)                    (DBSRNet on the device slot; CUDA-graph replay when enabled)
    copy-out stream:  D2H(i-1)                      (device result slot -> host_out)

`submit()` is asynchronous and returns a CUDA event that completes when `host_out` holds the prediction of that
submission; slots are recycled with event dependencies only (no host synchronisation inside `submit`).
"""
from __future__ import annotations

import torch

from . import ops


class HostPipeline:
    def __init__(self, net, depth: int = 2, device=None):
        assert depth >= 2
        self.net = net
        self.depth = depth
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        ops.require_device(torch.empty(1, device=self.device))
        self.compute = torch.cuda.current_stream(self.device)
        self.s_in = torch.cuda.Stream(device=self.device)
        self.s_out = torch.cuda.Stream(device=self.device)
        self._in = [None] * depth        # device input slots
        self._out = [None] * depth       # device result slots
        self._in_free = [None] * depth   # event: the forward that read input slot k has consumed it
        self._out_free = [None] * depth  # event: the D2H that read result slot k is complete
        self._i = 0

    def _engine(self):
        """the DBSREngine behind `net` when its forward is the fused engine path (then host copies go straight into / out of
        the engine's buffers), else None (any other module: `net(x)` + one device copy of the result)"""
        if getattr(self.net, '_fused_path', None) is not None and self.net._fused_path() and \
                not getattr(self.net, 'return_fusion_weights', False):
            eng = self.net.engine(self.device)
            return eng if eng.timers is None else None
        return None

    def submit(self, host_in: torch.Tensor, host_out: torch.Tensor) -> torch.cuda.Event:
        with torch.cuda.device(self.device):
            return self._submit(host_in, host_out)

    def _submit(self, host_in: torch.Tensor, host_out: torch.Tensor) -> torch.cuda.Event:
        """host_in: [B, N, 4, H, W] fp32 host tensor (pinned for a truly asynchronous copy); host_out: [B, 3, 8H, 8W] host
        tensor that receives `pred` -- fp32, or int16 when `net.output_int16` is set (the dtype must match the output).
        Returns the event to wait on before reading host_out."""
        if host_in.is_cuda or host_out.is_cuda:
            raise ValueError('HostPipeline moves HOST buffers; call the module directly for device tensors')
        k = self._i % self.depth
        self._i += 1
        eng = self._engine()
        quantize = bool(getattr(self.net, 'output_int16', False))
        want_dtype = torch.int16 if quantize else torch.float32
        if eng is not None and host_out.dtype != want_dtype:
            raise TypeError(f'host_out is {host_out.dtype} but the network output is {want_dtype} (net.output_int16={quantize})')
        graphed = eng is not None and bool(getattr(self.net, 'use_cuda_graph', False))
        if graphed:
            # slot k owns one CUDA graph of this shape: the H2D copy fills the graph's static input, the graph writes `pred`
            # into its static output, the D2H copy reads it from there -- no device-to-device copies around the forward
            with torch.cuda.stream(self.compute):
                _g, static_in, outs, _n = eng.graph_entry(tuple(host_in.shape), False, quantize, slot=k)
            dev_in, dev_out = static_in, outs[0]
            if self._in[k] is not dev_in or self._out[k] is not dev_out:
                self._in[k], self._out[k] = dev_in, dev_out
                self._in_free[k] = self._out_free[k] = None
                self.s_in.wait_stream(self.compute)       # the capture (and its warm-up) wrote both buffers
        else:
            if self._in[k] is None or self._in[k].shape != host_in.shape or self._in[k].dtype != torch.float32:
                self._in[k] = torch.empty(host_in.shape, dtype=torch.float32, device=self.device)
                self._in_free[k] = None
        # ---- H2D on the copy-in stream, once the previous user of this slot has consumed it
        if self._in_free[k] is not None:
            self.s_in.wait_event(self._in_free[k])
        with torch.cuda.stream(self.s_in):
            self._in[k].copy_(host_in, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.s_in)
        # ---- forward on the compute stream
        self.compute.wait_event(ev_in)
        with torch.cuda.stream(self.compute):
            if self._out_free[k] is not None:
                self.compute.wait_event(self._out_free[k])       # the D2H that last read this slot's result is complete
            if graphed:
                eng.forward_graphed(self._in[k], False, quantize, slot=k)
            elif eng is not None:
                B, _N, _c, H, W = host_in.shape
                shape = (B, 3, H * eng.up_r, W * eng.up_r)
                if self._out[k] is None or tuple(self._out[k].shape) != shape or self._out[k].dtype != want_dtype:
                    self._out[k] = torch.empty(shape, dtype=want_dtype, device=self.device)
                eng.forward(self._in[k], False, out={'pred': self._out[k]}, quantize=quantize)     # pred written in place
            else:
                pred, _aux = self.net(self._in[k])
                if self._out[k] is None or self._out[k].shape != pred.shape or self._out[k].dtype != pred.dtype:
                    self._out[k] = torch.empty_like(pred)
                if host_out.dtype != pred.dtype:
                    raise TypeError(f'host_out is {host_out.dtype} but the network output is {pred.dtype}')
                self._out[k].copy_(pred, non_blocking=True)
            self._in_free[k] = torch.cuda.Event()
            self._in_free[k].record(self.compute)
            ev_done = torch.cuda.Event()
            ev_done.record(self.compute)
        # ---- D2H on the copy-out stream
        self.s_out.wait_event(ev_done)
        with torch.cuda.stream(self.s_out):
            host_out.copy_(self._out[k], non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(self.s_out)
        self._out_free[k] = ev_out
        return ev_out

    def drain(self) -> None:
        self.s_in.synchronize()
        self.compute.synchronize()
        self.s_out.synchronize()
