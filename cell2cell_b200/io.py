"""Persistence helpers of the tensor objects: same functions as ``cell2cell/io/save_data.py:8-28`` and
``cell2cell/io/read_data.py:467-583``.  The pickles hold host numpy arrays (see
``BaseTensor.__getstate__``); loading puts the tensor, mask and ``loc_*`` arrays back on a CUDA device -- there is no CPU
path, so ``device`` must name a CUDA device (or be None: the current one)."""
import os
import pickle
from collections import OrderedDict

import pandas as pd

from . import engine

__all__ = ["export_variable_with_pickle", "load_variable_with_pickle", "load_tensor", "load_tensor_factors"]

_MAX_BYTES = 2 ** 31 - 1


def export_variable_with_pickle(variable, filename):
    """Pickle ``variable`` into ``filename`` in chunks of < 2 GiB (save_data.py:8-28)."""
    bytes_out = pickle.dumps(variable, protocol=pickle.HIGHEST_PROTOCOL)
    with open(filename, "wb") as f_out:
        for idx in range(0, len(bytes_out), _MAX_BYTES):
            f_out.write(bytes_out[idx:idx + _MAX_BYTES])
    print(filename, " was correctly saved.")


def load_variable_with_pickle(filename):
    """Counterpart of ``export_variable_with_pickle`` (read_data.py:467-493)."""
    bytes_in = bytearray(0)
    input_size = os.path.getsize(filename)
    with open(filename, "rb") as f_in:
        for _ in range(0, input_size, _MAX_BYTES):
            bytes_in += f_in.read(_MAX_BYTES)
    return pickle.loads(bytes_in)


def load_tensor(filename, backend=None, device=None):
    """Loads a communication tensor saved with ``write_file`` / ``export_variable_with_pickle`` (read_data.py:496-557).

    ``backend`` is accepted for signature compatibility: this implementation has one backend (its own CUDA kernels), so
    anything but None / 'pytorch' raises."""
    if backend not in (None, "pytorch"):
        raise ValueError("cell2cell_b200 has a single (CUDA) backend; got backend=%r" % (backend,))
    dev = engine.require_cuda(device)
    interaction_tensor = load_variable_with_pickle(filename)      # __setstate__ puts the arrays on the current device
    if hasattr(interaction_tensor, "to_device"):
        interaction_tensor.to_device(dev)
    return interaction_tensor


def load_tensor_factors(filename):
    """Factor loadings exported with ``export_factor_loadings``: one sheet per tensor dimension (read_data.py:560-583)."""
    xls = pd.ExcelFile(filename)
    factors = OrderedDict()
    for sheet_name in xls.sheet_names:
        factors[sheet_name] = xls.parse(sheet_name, index_col=0)
    return factors
