"""cell2cell_b200 -- the B200-native (sm_100a) hot path of Tensor-cell2cell behind the ``cell2cell.tensor`` API.

    from cell2cell_b200.tensor import InteractionTensor, PreBuiltTensor, build_context_ccc_tensor

The CUDA kernels live in ``csrc/`` and are reached through the C ABI of ``include/c2c_b200.h`` (``libc2c_b200.so``,
built in-tree by ``python -m cell2cell_b200._build`` / ``__graft_entry__.build()``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import tensor  # noqa: E402,F401
from . import analysis, io, spatial  # noqa: E402,F401
