"""bithtm_b200 -- B200 (sm_100a) implementation of bitHTM's spatial-pooler +
temporal-memory timestep behind the reference's ``bithtm.networks`` /
``projections`` / ``regularizations`` API (``bithtm/__init__.py:1-6``).

    import bithtm_b200 as bithtm
    htm = bithtm.HierarchicalTemporalMemory(input_dim, column_dim, cell_dim)
    sp_state, tm_state = htm.process(input_bits)

The arithmetic runs in hand-written CUDA kernels (``csrc/``) behind a C ABI
(``include/bithtm_b200.h``); there is no CPU fallback.
"""

from . import batch, networks, projections, regularizations  # noqa: F401

SpatialPooler = networks.SpatialPooler
TemporalMemory = networks.TemporalMemory
HierarchicalTemporalMemory = networks.HierarchicalTemporalMemory
StreamBatch = batch.StreamBatch  # extension: independent streams side by side on one GPU

__all__ = ["SpatialPooler", "TemporalMemory", "HierarchicalTemporalMemory", "StreamBatch", "networks",
           "projections", "regularizations", "batch"]
