"""Drop-in alias: ``import proxmin`` resolves to the B200-native implementation (package ``proxmin_b200``).

Same public surface as the reference package: ``pgm, adaprox, admm, sdmm, bsdmm``, the ``prox_*`` operators,
``AlternatingProjections`` and the sub-modules ``nmf, utils, algorithms, operators``.
"""
import sys as _sys

import proxmin_b200 as _impl
from proxmin_b200 import *  # noqa: F401,F403
from proxmin_b200 import adaprox, admm, algorithms, bsdmm, nmf, operators, pgm, sdmm, utils  # noqa: F401

for _name in ("nmf", "utils", "algorithms", "operators"):
    _sys.modules[__name__ + "." + _name] = getattr(_impl, _name)
__version__ = _impl.__version__
