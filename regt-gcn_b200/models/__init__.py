"""Drop-in ``models`` package for the RegT-GCN hot path (reference: models/__init__.py:1-13).

Only the classes on the accelerated path are provided: ``RegionalTemporalGCN`` (also used by the
reference for ``--model RandomTemporalGCN``), ``TemporalGCN`` and ``models.utils.TGCN``.
The reference's other baselines (GAT, GraphSAGE, STID, ...) are out of scope."""
from models.RegionalTemporalGCN import RegionalTemporalGCN
from models.TemporalGCN import TemporalGCN

__all__ = ["RegionalTemporalGCN", "TemporalGCN"]
