"""proxmin_b200 -- B200-native implementation of proxmin's proximal-update hot path.

Drop-in for the reference's public names: ``pgm, adaprox, admm, sdmm, bsdmm``, every
``prox_*`` operator, ``AlternatingProjections`` and the sub-modules ``nmf``, ``utils``,
``algorithms``, ``operators`` (proxmin/__init__.py:1-4).  All arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/proxmin_b200.h.
"""
from .algorithms import pgm, adaprox, admm, sdmm, bsdmm  # noqa: F401
from .operators import *  # noqa: F401,F403
from .operators import AlternatingProjections  # noqa: F401
from . import nmf  # noqa: F401
from . import utils  # noqa: F401
from . import algorithms  # noqa: F401
from . import operators  # noqa: F401
from . import workloads  # noqa: F401

__version__ = "0.1.0"


def init_distributed(unique_id, world, rank):
    """Join the NCCL communicator of a one-process-per-GPU job (see bench.py for the id exchange)."""
    from . import _ffi

    _ffi.context().comm_init(unique_id, world, rank)
