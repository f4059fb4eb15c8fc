"""Drop-in namespace for the hot path: ``import normalisr_b200.normalisr as norm`` then
``norm.coex(dt, dc)`` / ``norm.de(dg, dt, dc)`` exactly as with
``import normalisr.normalisr as norm`` (reference src/normalisr/normalisr.py:3-9).
The association-testing entry points, their immediate consumer ``binnet`` and the step directly
upstream, ``normvar``, are provided here; the other steps (``lcpm``, ``normcov``, ``compute_var`` ...) stay with the reference package
and their output feeds these functions unchanged."""
from .binnet import binnet
from .coex import coex
from .de import de
from .lcpm import lcpm
from .norm import compute_var, normvar

__all__ = ["coex", "de", "binnet", "normvar", "lcpm", "compute_var"]
