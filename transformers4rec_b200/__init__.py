"""t4r_b200: the session-sequence transformer hot path of Transformers4Rec,
hand-written for sm_100a behind the reference's module API.

    import transformers4rec_b200.torch as tr      # mirrors ``transformers4rec.torch``

The CUDA kernels live in ``libt4r_b200.so`` (built in-tree by ``build()``); there
is no CPU or eager fallback -- importing works anywhere, calling needs a B200.
"""
from ._lib import T4RError, build, load  # noqa: F401

__version__ = "0.1.0"
