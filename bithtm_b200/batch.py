"""Independent streams batched on one GPU (BASELINE configs[3]).

Every stream is its own ``HierarchicalTemporalMemory`` (own permanence matrix, own
segments, own MT19937 stream -- the reference's per-step semantics, networks.py:31-32,
make the overlap a bit-GEMV per stream, so there is nothing to share between them).  A
small network's fused step kernel occupies one 16-CTA cluster, i.e. 16 of 148 SMs;
``StreamBatch`` puts one such launch per stream into ONE CUDA graph without dependencies
between them, so the streams advance side by side.  Streams are partitioned over GPUs
trivially (one ``StreamBatch`` per process / GPU, no collective).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


class StreamBatch:
    def __init__(self, networks):
        self.networks = list(networks)
        if not self.networks:
            raise ValueError("StreamBatch needs at least one network")
        engines = [h.engine for h in self.networks]
        dev = engines[0].device
        for h, e in zip(self.networks, engines):
            if e.device != dev:
                raise ValueError("all networks of a StreamBatch live on one device")
            if not e.ctx.fused_mode or e.ctx.ring_len <= 0:
                raise ValueError("StreamBatch needs networks with a fused step kernel and a device input ring "
                                 "(ring_len > 0)")
            if h.temporal_memory._rng.mode != "lazy":
                raise ValueError('StreamBatch needs rng_sync="lazy" networks (no per-step host round trip)')
        self.engines = engines
        self._graphs = {}

    def load_inputs(self, inputs, rng_states=None):
        """inputs[i]: bool [ring_len, input_dim] for stream i (uploaded into its device ring).
        rng_states[i]: the ``np.random.get_state()`` tuple stream i continues from (default: the
        global np.random state at this call, for every stream)."""
        for i, (h, e, x) in enumerate(zip(self.networks, self.engines, inputs)):
            if rng_states is not None:
                h.temporal_memory._rng.adopt(e, rng_states[i])
            else:
                h.temporal_memory._rng.before(e)  # adopt np.random's state once (lazy mode)
            e.load_ring(x)

    def _graph(self, steps, learning):
        key = (int(steps), bool(learning))
        if key not in self._graphs:
            import torch

            arr = (nat._CTXP * len(self.engines))(*[C.pointer(e.ctx) for e in self.engines])
            handle = C.c_void_p()
            eng = self.engines[0]
            side = torch.cuda.Stream(device=eng.device)
            side.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(side):
                nat.check(nat.lib.bh_batch_graph_create(arr, len(self.engines), key[0], int(key[1]), eng.stream,
                                                        C.byref(handle)), "bh_batch_graph_create")
            torch.cuda.current_stream(eng.device).wait_stream(side)
            self._graphs[key] = handle
        return self._graphs[key]

    def run(self, steps=1, learning=True):
        """Advance every stream by `steps` timesteps (asynchronous; inputs come from the rings)."""
        handle = self._graph(steps, learning)
        nat.check(nat.lib.bh_graph_launch(handle, self.engines[0].stream), "bh_graph_launch")
        for e in self.engines:
            e.epoch += steps

    def check_status(self):
        for e in self.engines:
            e.check_status()

    def __del__(self):
        try:
            for h in self._graphs.values():
                nat.lib.bh_graph_destroy(h)
        except Exception:
            pass
