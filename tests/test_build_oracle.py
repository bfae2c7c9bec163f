"""The tensor-build oracle (oracle/build_oracle.py) against the golden fixtures made by the reference's own code
(tests/golden/make_golden.py) -- bit for bit -- and, when /root/reference is present, against the live reference."""
import json
import os

import numpy as np
import pytest

from oracle.build_oracle import interaction_tensor_oracle
from oracle.ref_harness import load_reference, reference_available
from tests.cases import OPTION_GRID, hard_case, toy_ppi, toy_rnaseq

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _check(store, key, res, exact=True):
    # 'mean' / 'gmean' over the subunits of a complex sum in the iteration order of a Python *set* of names
    # (rnaseq.py:178-179), which changes with the process's string-hash seed: for >= 3 subunits the reference itself is
    # only reproducible to the last fp64 bit, so those cases are compared to 2 ulp(fp64) instead of bit for bit.
    if exact:
        np.testing.assert_array_equal(res["tensor"], store[key + "/tensor"])
    else:
        np.testing.assert_allclose(res["tensor"], store[key + "/tensor"], rtol=5e-16, atol=0)
    assert res["tensor"].dtype == store[key + "/tensor"].dtype
    assert bool(store[key + "/has_mask"]) == (res["mask"] is not None)
    if res["mask"] is not None:
        np.testing.assert_array_equal(res["mask"].astype(np.uint8), store[key + "/mask"])
    np.testing.assert_array_equal(res["loc_zeros"].astype(np.uint8), store[key + "/loc_zeros"])
    assert json.loads(str(store[key + "/names"])) == [list(x) for x in res["order_names"]]
    assert json.loads(str(store[key + "/genes"])) == list(res["genes"])


def test_toy_golden():
    store = np.load(os.path.join(GOLD, "build_toy.npz"))
    for score in ("expression_product", "expression_mean", "expression_gmean"):
        for cplx in (True, False):
            res = interaction_tensor_oracle([toy_rnaseq()], toy_ppi(cplx), communication_score=score,
                                            complex_sep="&" if cplx else None)
            _check(store, "%s|%d" % (score, cplx), res)
    a = toy_rnaseq()
    b = toy_rnaseq().drop(columns=["C2"]).drop(index=["Protein-D"]) * 2
    for how in ("inner", "outer", "outer_genes", "outer_cells"):
        res = interaction_tensor_oracle([a, b], toy_ppi(True), communication_score="expression_gmean", complex_sep="&", how=how)
        _check(store, "two|" + how, res)


def test_toy_known_answer():
    # min(C, E) (x) F for the complex row of the toy data
    res = interaction_tensor_oracle([toy_rnaseq()], toy_ppi(True), communication_score="expression_product", complex_sep="&")
    assert res["order_names"][1][7] == "PROTEIN-C&PROTEIN-E^PROTEIN-F"
    expect = [[60, 22, 32, 10, 24], [30, 11, 16, 5, 12], [30, 11, 16, 5, 12], [810, 297, 432, 135, 324], [450, 165, 240, 75, 180]]
    np.testing.assert_array_equal(res["tensor"][0, 7], np.asarray(expect))


@pytest.mark.parametrize("seed", [0, 1])
def test_hard_golden(seed):
    store = np.load(os.path.join(GOLD, "build_hard.npz"))
    mats, ppi = hard_case(seed)
    for i, opt in enumerate(OPTION_GRID):
        for up in ((True, False) if i % 12 == 0 else (True,)):
            res = interaction_tensor_oracle(mats, ppi, complex_sep="&", upper_letter_comparison=up, **opt)
            _check(store, "%d|%d|%d" % (seed, i, up), res, exact=opt["complex_agg_method"] == "min")


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_live_reference_fresh_seed():
    ref = load_reference()
    mats, ppi = hard_case(5, n_contexts=3, n_genes=12, n_cells=4, n_lr=15)
    for opt in OPTION_GRID[::7]:
        t = ref.tensor.InteractionTensor([m.copy() for m in mats], ppi.copy(), complex_sep="&", verbose=False, **opt)
        res = interaction_tensor_oracle(mats, ppi, complex_sep="&", **opt)
        np.testing.assert_array_equal(res["tensor"], np.asarray(t.tensor))
        if t.mask is None:
            assert res["mask"] is None
        else:
            np.testing.assert_array_equal(res["mask"], np.asarray(t.mask))
        assert [list(x) for x in t.order_names] == [list(x) for x in res["order_names"]]
