"""Drop-in namespace for the hot path: ``import normalisr_b200.normalisr as norm`` then
``norm.coex(dt, dc)`` / ``norm.de(dg, dt, dc)`` exactly as with
``import normalisr.normalisr as norm`` (reference src/normalisr/normalisr.py:3-9).
Only the association-testing entry points and their immediate consumer ``binnet`` are provided
here; the upstream steps (``lcpm``, ``normcov``, ``normvar`` ...) stay with the reference package
and their output feeds these functions unchanged."""
from .binnet import binnet
from .coex import coex
from .de import de

__all__ = ["coex", "de", "binnet"]
