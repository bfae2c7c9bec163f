"""Golden fixtures of ``dataframes_to_tensor`` produced by the REFERENCE'S OWN CODE (cell2cell/tensor/external_scores.py,
imported through oracle/ref_harness.py).  Run in the build container only:

    python tests/golden/make_golden_dataframes.py        ->  tests/golden/dataframes.npz

The inputs are regenerated from seeds by ``tests/cases.long_tables``.
"""
import importlib
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_harness import load_reference  # noqa: E402
from tests.cases import LONG_TABLE_OPTIONS, long_tables  # noqa: E402

SEEDS = (0, 1)


def main():
    warnings.simplefilter("ignore")
    load_reference()
    ext = importlib.import_module("cell2cell.tensor.external_scores")
    store = {}
    for seed in SEEDS:
        tables = long_tables(seed)
        for i, kw in enumerate(LONG_TABLE_OPTIONS):
            t = ext.dataframes_to_tensor({k: v.copy() for k, v in tables.items()}, "source", "target", "ligand", "receptor", "score", **kw)
            key = "%d|%d" % (seed, i)
            store[key + "/tensor"] = np.asarray(t.tensor, dtype=np.float64)
            store[key + "/has_mask"] = np.array(t.mask is not None)
            if t.mask is not None:
                store[key + "/mask"] = np.asarray(t.mask).astype(np.uint8)
            store[key + "/loc_nans"] = np.asarray(t.loc_nans).astype(np.uint8)
            store[key + "/loc_zeros"] = np.asarray(t.loc_zeros).astype(np.uint8)
            store[key + "/names"] = np.array(json.dumps(t.order_names))
    np.savez_compressed(os.path.join(HERE, "dataframes.npz"), **store)
    print("wrote", len(store), "arrays")


if __name__ == "__main__":
    main()
