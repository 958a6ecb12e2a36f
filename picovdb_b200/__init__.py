"""picovdb_b200 -- B200-native exact cosine top-k engine behind the PicoVectorDB API.

``from picovdb_b200 import PicoVectorDB`` is a drop-in for ``from picovdb import PicoVectorDB``
(wensheng/picovdb) whose NumPy query path runs as hand-written sm_100a CUDA kernels.
"""
from .db import (  # noqa: F401
    PicoVectorDB,
    K_ID,
    K_METRICS,
    K_VECTOR,
    Float,
    _HAS_FAISS,
)

__all__ = ["PicoVectorDB", "K_METRICS", "K_ID", "K_VECTOR", "_HAS_FAISS"]
__version__ = "0.1.0"
