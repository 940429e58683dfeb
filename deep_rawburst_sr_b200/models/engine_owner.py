"""Mixin of the drop-in modules that run on a cached `DBSREngine` (packed weights, workspaces, CUDA graphs).

The engine is a snapshot of the module's parameters, so it must be rebuilt whenever they change.  Three mechanisms:
  * `_apply` (`.to()`, `.cuda()`, `.half()`, ...) and `load_state_dict` on the module itself drop the engine;
  * every `engine()` call compares a fingerprint of the parameters -- the tuple of their autograd version counters -- with
    the one the engine was built from.  That catches everything that goes through a CHILD module instead of this one
    (`net.decoder.load_state_dict(...)`, `net.encoder.alignment_net.to(...)`), in-place updates (`p.copy_`, `p.add_`,
    an optimizer step) and `p.data = ...` (~25 us per call for the 231 tensors of DBSRNet);
  * `invalidate_engine()` for what no counter sees: writes through `p.data` / `p.detach()` views (`p.data.copy_(...)`) and
    Parameter objects REPLACED after the first forward (`conv.weight = nn.Parameter(...)`).
"""
import torch


class EngineOwner:
    _engine = None
    _engine_params = None
    _engine_fp = None

    def invalidate_engine(self):
        """drop the cached engine (packed weights, CUDA graphs); the next forward rebuilds it from the current parameters"""
        self._engine = None
        self._engine_params = None
        self._engine_fp = None
        return self

    def _apply(self, fn, *a, **k):
        self.invalidate_engine()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate_engine()
        return super().load_state_dict(*a, **k)

    def _weights_fingerprint(self):
        ps = self._engine_params
        if ps is None:
            ps = self._engine_params = list(self.parameters())
        return tuple(p._version for p in ps)

    def _engine_is_current(self, device, **attrs) -> bool:
        """True when the cached engine exists, lives on `device`, has the given attribute values and was built from the
        parameters as they are now"""
        e = self._engine
        if e is None or e.device != torch.device(device):
            return False
        for k, v in attrs.items():
            if getattr(e, k) != v:
                return False
        return self._engine_fp == self._weights_fingerprint()

    def _set_engine(self, engine):
        self._engine = engine
        self._engine_fp = self._weights_fingerprint()
        return engine
