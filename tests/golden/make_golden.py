"""Generates the golden fixtures of the tensor-build stage by running the REFERENCE'S OWN CODE (read-only at
/root/reference, imported through oracle/ref_harness.py).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/build_toy.npz, tests/golden/build_hard.npz, tests/golden/build_hard32.npz.  The inputs are not stored: they are
regenerated from seeds by tests/cases.py, which this script shares with the tests.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_harness import load_reference  # noqa: E402
from tests.cases import OPTION_GRID, hard_case, toy_ppi, toy_rnaseq  # noqa: E402

HARD_SEEDS = (0, 1)
HARD32_SEEDS = (2, 3)


def run_reference(ref, mats, ppi, **kw):
    t = ref.tensor.InteractionTensor([m.copy() for m in mats], ppi.copy(), verbose=False, **kw)
    return dict(tensor=np.asarray(t.tensor), mask=None if t.mask is None else np.asarray(t.mask),
                loc_nans=np.asarray(t.loc_nans), loc_zeros=np.asarray(t.loc_zeros),
                names=json.dumps(t.order_names), genes=json.dumps(list(t.genes)))


def pack(store, key, res):
    store[key + "/tensor"] = res["tensor"]
    store[key + "/has_mask"] = np.array(res["mask"] is not None)
    if res["mask"] is not None:
        store[key + "/mask"] = res["mask"].astype(np.uint8)
    store[key + "/loc_zeros"] = res["loc_zeros"].astype(np.uint8)
    store[key + "/names"] = np.array(res["names"])
    store[key + "/genes"] = np.array(res["genes"])


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    toy = {}
    for score in ("expression_product", "expression_mean", "expression_gmean"):
        for cplx in (True, False):
            kw = dict(communication_score=score, complex_sep="&" if cplx else None)
            pack(toy, "%s|%d" % (score, cplx), run_reference(ref, [toy_rnaseq()], toy_ppi(cplx), **kw))
    # two contexts with different cell sets -> mask under how='outer'
    a = toy_rnaseq()
    b = toy_rnaseq().drop(columns=["C2"]).drop(index=["Protein-D"]) * 2
    for how in ("inner", "outer", "outer_genes", "outer_cells"):
        kw = dict(communication_score="expression_gmean", complex_sep="&", how=how)
        pack(toy, "two|" + how, run_reference(ref, [a, b], toy_ppi(True), **kw))
    np.savez_compressed(os.path.join(HERE, "build_toy.npz"), **toy)

    hard = {}
    for seed in HARD_SEEDS:
        mats, ppi = hard_case(seed)
        for i, opt in enumerate(OPTION_GRID):
            for up in ((True, False) if i % 12 == 0 else (True,)):
                res = run_reference(ref, mats, ppi, complex_sep="&", upper_letter_comparison=up, **opt)
                pack(hard, "%d|%d|%d" % (seed, i, up), res)
    np.savez_compressed(os.path.join(HERE, "build_hard.npz"), **hard)
    # fp32-representable inputs (float64 frames): what the GPU path is compared with after rounding to fp32
    hard32 = {}
    for seed in HARD32_SEEDS:
        mats, ppi = hard_case(seed, fp32_values=True)
        for i, opt in enumerate(OPTION_GRID):
            res = run_reference(ref, mats, ppi, complex_sep="&", **opt)
            pack(hard32, "%d|%d|1" % (seed, i), res)
    np.savez_compressed(os.path.join(HERE, "build_hard32.npz"), **hard32)
    print("toy cases:", len([k for k in toy if k.endswith("/tensor")]),
          "hard cases:", len([k for k in hard if k.endswith("/tensor")]))


if __name__ == "__main__":
    main()
