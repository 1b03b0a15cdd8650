"""cgic_b200 -- B200-native (sm_100a) implementation of Control-GIC's VQ + entropy-coding hot path.

The directory is named `control-gic_b200`; import it as `cgic_b200` (the top-level shim
`cgic_b200.py` aliases it).  Python here is the host-side mirror of the reference's
operator API; all compute is in libcgic_b200.so (include/cgic_b200.h), loaded with ctypes.
"""
from . import _lib, ops, dist, inference  # noqa: F401
from .codec import BinaryCoding, HuffmanCoding  # noqa: F401
from .decoder import Normalize, SpatialNorm  # noqa: F401
from .entropy import Entropy, entropy_pair  # noqa: F401
from .model import CGIC, ReferenceEncoderHeads  # noqa: F401
from .quantize import VectorQuantize2  # noqa: F401
from .router import TripleGrainFixedEntropyRouter  # noqa: F401

__all__ = ["ops", "dist", "inference", "HuffmanCoding", "BinaryCoding", "Entropy", "entropy_pair", "CGIC", "ReferenceEncoderHeads",
           "VectorQuantize2", "TripleGrainFixedEntropyRouter", "SpatialNorm", "Normalize"]
