"""Lock-step link between the device MT19937 stream and the global legacy ``np.random`` state
(the reference draws everything from it: networks.py:87, projections.py:120,235)."""

from __future__ import annotations

import numpy as np

from . import _native as nat


class _RngLink:
    """Keeps the device MT19937 state and the global ``np.random`` state in step.

    np.random.get_state()/set_state() cost ~40 us each, so the legacy global
    RandomState's raw state (624 key words + position, numpy/random/src/mt19937) is
    read and written in place through the address NumPy publishes for that purpose
    (``BitGenerator.ctypes.state_address``).  The Gaussian cache of the legacy
    generator is never touched, exactly as rand() does not touch it."""

    def __init__(self, mode="step"):
        assert mode in ("step", "lazy")
        self.mode = mode
        self._seeded = False
        import ctypes

        addr = np.random.mtrand._rand._bit_generator.ctypes.state_address
        self._live_key = np.ctypeslib.as_array((ctypes.c_uint32 * nat.MT_N).from_address(addr))
        self._live_pos = ctypes.c_int.from_address(addr + 4 * nat.MT_N)
        self._key = np.zeros(nat.MT_N, dtype=np.uint32)  # what the device continues from
        self._key_bytes = b""
        self._pos = -1

    def before(self, eng):
        if self.mode == "lazy" and self._seeded:
            return
        pos = self._live_pos.value
        if self._seeded and pos == self._pos and self._live_key.tobytes() == self._key_bytes:
            return  # nobody drew from np.random since our last write-back
        self._key[:] = self._live_key
        self._key_bytes = self._key.tobytes()
        self._pos = pos
        eng.set_rng_state(self._key, pos)
        self._seeded = True

    def adopt(self, eng, state):
        """Continue the device stream from an explicit ``np.random.get_state()`` tuple."""
        key, pos = np.asarray(state[1], dtype=np.uint32), int(state[2])
        self._key[:] = key
        self._key_bytes = self._key.tobytes()
        self._pos = pos
        eng.set_rng_state(self._key, pos)
        self._seeded = True

    def after(self, eng, summary=None):
        if self.mode == "lazy":
            return
        if summary is None:
            key, pos = eng.get_rng_state()
        else:
            k = eng.k
            tail = summary[4 + 4 * k:4 + 4 * k + nat.MT_N + 1]
            key, pos = tail[:nat.MT_N].view(np.uint32), int(tail[nat.MT_N])
        self._live_key[:] = key
        self._key_bytes = self._live_key.tobytes()
        self._pos = pos
        self._live_pos.value = pos

    def sync(self, eng):
        key, pos = eng.get_rng_state()
        self._live_key[:] = key
        self._key_bytes = self._live_key.tobytes()
        self._pos = pos
        self._live_pos.value = pos
