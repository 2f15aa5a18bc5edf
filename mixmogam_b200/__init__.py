"""
mixmogam_b200 -- B200-native EMMAX hot path (kinship -> REML -> SNP scan) behind mixmogam's Python API.

    from mixmogam_b200 import kinship, linear_models as lm, hdf5_data

The numerics run in hand-written sm_100a CUDA (libmixmogam_b200.so, C ABI in include/mixmogam_b200.h);
there is no CPU fallback: without the built library or without a B200 the calls raise.
"""
from . import _lib            # noqa: F401  (fails loudly if the shared library is missing)
from . import kinship         # noqa: F401
from . import linear_models   # noqa: F401
from . import hdf5_data       # noqa: F401
from ._lib import (Context, DeviceMatrix, MmgError, PackedGenotypes, get_context, load_library, pack_genotypes,  # noqa: F401
                   resident)

__version__ = '0.1.0'
