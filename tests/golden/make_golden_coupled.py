"""Golden fixtures of the coupled factorization produced by the REFERENCE'S OWN ``coupled_non_negative_parafac``
(cell2cell/tensor/coupled_factorization.py:123-482, imported unmodified through ``oracle/ref_harness.load_reference_coupled``;
its tensorly primitives are the numpy restatements of oracle/ncp_oracle.py -- tensorly is not installable here).  Run in the
build container only:

    python tests/golden/make_golden_coupled.py        ->  tests/golden/coupled.npz

The inputs are regenerated from seeds by ``tests/cases.coupled_inputs``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_harness import load_reference_coupled  # noqa: E402
from tests.cases import COUPLED_CASES, coupled_inputs  # noqa: E402


def main():
    ref = load_reference_coupled()
    store = {}
    for name, case in COUPLED_CASES.items():
        x1, x2, m1, m2 = coupled_inputs(name)
        mm = ref._process_mode_mapping(x1, x2, case["mode_mapping"])
        f64 = lambda m: None if m is None else m.astype(np.float64)  # noqa: E731
        cp1, cp2, (e1, e2) = ref.coupled_non_negative_parafac(
            x1.astype(np.float64), x2.astype(np.float64), case["rank"], mm, mask1=f64(m1), mask2=f64(m2),
            n_iter_max=case["n_iter_max"], init="random", tol=case["tol"], random_state=7, return_errors=True,
            **case.get("kwargs", {}))
        for k, cp in ((1, cp1), (2, cp2)):
            store["%s/weights%d" % (name, k)] = np.asarray(cp[0])
            for i, f in enumerate(cp[1]):
                store["%s/factors%d_%d" % (name, k, i)] = np.asarray(f)
        store[name + "/errors1"] = np.asarray(e1, dtype=np.float64)
        store[name + "/errors2"] = np.asarray(e2, dtype=np.float64)
        print(name, "iterations", len(e1), "errors", float(e1[-1]), float(e2[-1]))
    np.savez_compressed(os.path.join(HERE, "coupled.npz"), **store)


if __name__ == "__main__":
    main()
