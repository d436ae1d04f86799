"""Boosting and inhibition plugins -- device-backed mirrors of
``bithtm/regularizations.py`` (same class names, constructor arguments, defaults
and ``process``/``update`` methods; reference lines cited per method).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


class ExponentialBoosting:
    """regularizations.py:4-21.  ``duty_cycle`` lives in HBM (float32 [C])."""

    def __init__(self, output_dim, active_outputs, intensity=0.3, momentum=0.99):
        self.output_dim = output_dim
        self.active_outputs = active_outputs
        self.density = active_outputs / output_dim  # regularizations.py:9
        self.intensity = intensity
        self.momentum = momentum
        self._engine = None

    # -- constants exactly as NumPy evaluates them ----------
    def _constants(self):
        return dict(
            boost_coef=float(np.float32(-(self.intensity / self.density))),  # :16, weak python float -> f32
            duty_momentum=float(np.float32(self.momentum)),  # :20
            duty_increment=float(np.float32(1.0 - self.momentum)),  # :21
        )

    def _bind(self, engine):
        key = (self.intensity, self.density, self.momentum, id(engine))
        if getattr(self, "_bound_key", None) == key:
            return  # constants unchanged since the last call
        self._engine = engine
        for k, v in self._constants().items():
            setattr(engine.ctx, k, v)
        self._bound_key = key

    def _need_engine(self):
        if self._engine is None:
            raise RuntimeError("ExponentialBoosting is not attached to a network yet; construct a "
                               "bithtm_b200 SpatialPooler with it (boosting=...) first")
        return self._engine

    @property
    def duty_cycle(self):
        if self._engine is None:
            return np.zeros(self.output_dim, dtype=np.float32)  # :13
        return self._engine.buf["duty"].cpu().numpy()

    def process(self, input_activation):
        """regularizations.py:15-17 on the overlaps currently on the device (the
        argument is accepted for signature parity and uploaded if it is a host array)."""
        eng = self._need_engine()
        self._bind(eng)
        if input_activation is not None and not getattr(input_activation, "_bh_on_device", False):
            import torch

            host = np.asarray(input_activation)
            eng.buf["overlaps"].copy_(torch.from_numpy(host.astype(np.int32)).to(eng.device))
        nat.check(nat.lib.bh_boost(eng.ref, eng.stream), "bh_boost")
        return eng.buf["boosted"].cpu().numpy()

    def update(self, active_input):
        """regularizations.py:19-21."""
        eng = self._need_engine()
        self._bind(eng)
        eng_set_active(eng, active_input)
        nat.check(nat.lib.bh_duty_update(eng.ref, eng.stream), "bh_duty_update")


class GlobalInhibition:
    """regularizations.py:24-29 with a *defined* result: the k largest inputs, ties
    broken towards the lower column index, returned in ascending column index.
    (``np.argpartition``'s tie-break and output order depend on NumPy's SIMD
    dispatch)  Use the reference's own ``GlobalInhibition``
    object as ``inhibition=`` to reproduce a particular host's argpartition order:
    any object with ``process(boosted) -> active_column`` is accepted and run on the
    host."""

    _bh_native = True

    def __init__(self, active_outputs):
        self.active_outputs = active_outputs
        self._engine = None

    def _bind(self, engine):
        self._engine = engine

    def process(self, input_activation):
        eng = self._engine
        if eng is None:
            raise RuntimeError("GlobalInhibition is not attached to a network yet")
        if input_activation is not None and not getattr(input_activation, "_bh_on_device", False):
            import torch

            host = np.asarray(input_activation, dtype=np.float64) + 0.0  # -0.0 -> +0.0 (they compare equal)
            if host.shape != (eng.C_local,):
                raise ValueError(f"expected {eng.C_local} values, got shape {host.shape}")
            if (host < 0).any():
                # the device orders keys by their bit patterns as unsigned integers, which is the numeric
                # order for non-negative doubles only: map arbitrary doubles with the order-preserving
                # transform (negative: flip all bits, else: flip the sign bit)
                bits = host.view(np.int64)
                bits = np.where(bits < 0, ~bits, bits ^ np.int64(-2 ** 63))
                host = bits.view(np.float64)
            eng.buf["boosted"].copy_(torch.from_numpy(np.ascontiguousarray(host)).to(eng.device))
        if int(self.active_outputs) != eng.k:
            raise ValueError("active_outputs differs from the network's active_columns")
        nat.check(nat.lib.bh_inhibit(eng.ref, eng.stream), "bh_inhibit")
        sc = eng.scalars()
        cur = int(sc[nat.SC_STEP]) & 1
        k = eng.ctx.active_columns
        return eng.buf["active_cols"][cur * k:(cur + 1) * k].cpu().numpy().astype(np.int64)


def eng_set_active(eng, active_column):
    """Adopt an explicit ordered active-column list (host-inhibition mode)."""
    import torch

    cols = np.ascontiguousarray(np.asarray(active_column).reshape(-1), dtype=np.int32)
    if cols.size != eng.k:
        raise ValueError(f"expected {eng.k} active columns (the network's active_columns), got {cols.size}")
    dev = torch.from_numpy(cols).to(eng.device)
    nat.check(nat.lib.bh_set_active_columns(eng.ref, dev.data_ptr(), eng.stream), "bh_set_active_columns")
