"""Host half of the product's tensor build (cell2cell_b200/tensor/build_index.py) against the reference-made golden
fixtures: the index arrays it hands to ``c2c_build_ccc`` are run through a numpy emulation of the kernel's gather
(tests/emulate.py) and must reproduce the reference's tensor, mask and names."""
import json
import os

import numpy as np
import pytest

from cell2cell_b200.tensor import build_index as bi
from tests.cases import OPTION_GRID, hard_case, toy_ppi, toy_rnaseq
from tests.emulate import emulate_build, emulate_group

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def host_build(mats, ppi, how="inner", outer_fraction=0.0, communication_score="expression_mean", complex_sep=None,
               complex_agg_method="min", upper_letter_comparison=True, interaction_columns=("A", "B"), group_ppi_by=None,
               group_ppi_method="gmean"):
    """Mirror of cell2cell_b200.tensor.tensor._build with the device calls replaced by the numpy emulation."""
    complexes = bi.complex_table(ppi, complex_sep, interaction_columns)[4] if complex_sep is not None else None
    tables = [bi.context_rows(list(df.index), complexes, upper_letter_comparison) for df in mats]
    names = [t.names for t in tables]
    columns = [list(df.columns) for df in mats]
    genes = bi.ordered(names[0], bi.select_elements(names, how in ("outer", "outer_genes"), outer_fraction))
    cells = bi.ordered(columns[0], bi.select_elements(columns, how in ("outer", "outer_cells"), outer_fraction))
    f = bi.filter_lr_pairs(ppi, genes, complex_sep, upper_letter_comparison, interaction_columns)
    lig, rec = f[interaction_columns[0]].tolist(), f[interaction_columns[1]].tolist()
    idx = bi.resolve(tables, columns, genes, cells, lig, rec)
    X = emulate_build([df.values for df in mats], idx, len(mats), len(lig), len(cells), communication_score, complex_agg_method)
    if group_ppi_by is not None:
        keys, off, rows = bi.group_rows(f, group_ppi_by)
        X = emulate_group(X, off, rows, group_ppi_method)
        ppi_names = list(keys)
    else:
        ppi_names = [a + "^" + b for a, b in zip(lig, rec)]
    mask = None if how == "inner" else (~np.isnan(X)).astype(np.uint8)
    return np.nan_to_num(X), mask, genes, cells, ppi_names


def _check(store, key, out):
    X, mask, genes, cells, ppi_names = out
    np.testing.assert_allclose(X, store[key + "/tensor"], rtol=5e-16, atol=0)
    assert bool(store[key + "/has_mask"]) == (mask is not None)
    if mask is not None:
        np.testing.assert_array_equal(mask, store[key + "/mask"])
    names = json.loads(str(store[key + "/names"]))
    assert names[1] == list(ppi_names) and names[2] == list(cells) and names[3] == list(cells)
    assert json.loads(str(store[key + "/genes"])) == list(genes)


def test_toy():
    store = np.load(os.path.join(GOLD, "build_toy.npz"))
    for score in ("expression_product", "expression_mean", "expression_gmean"):
        for cplx in (True, False):
            _check(store, "%s|%d" % (score, cplx),
                   host_build([toy_rnaseq()], toy_ppi(cplx), communication_score=score, complex_sep="&" if cplx else None))
    a = toy_rnaseq()
    b = toy_rnaseq().drop(columns=["C2"]).drop(index=["Protein-D"]) * 2
    for how in ("inner", "outer", "outer_genes", "outer_cells"):
        _check(store, "two|" + how,
               host_build([a, b], toy_ppi(True), communication_score="expression_gmean", complex_sep="&", how=how))


@pytest.mark.parametrize("seed", [0, 1])
def test_hard(seed):
    store = np.load(os.path.join(GOLD, "build_hard.npz"))
    mats, ppi = hard_case(seed)
    for i, opt in enumerate(OPTION_GRID):
        for up in ((True, False) if i % 12 == 0 else (True,)):
            _check(store, "%d|%d|%d" % (seed, i, up),
                   host_build(mats, ppi, complex_sep="&", upper_letter_comparison=up, **opt))


def test_partner_outside_universe_raises():
    mats = [toy_rnaseq()]
    with pytest.raises(KeyError):
        tables = [bi.plain_rows(list(df.index)) for df in mats]
        bi.resolve(tables, [list(mats[0].columns)], list(mats[0].index), list(mats[0].columns), ["Protein-A"], ["NOPE"])


def test_lpt_and_groups():
    import pandas as pd
    ppi = pd.DataFrame({"A": list("abcd"), "B": list("efgh"), "g": ["y", "x", "y", "x"]})
    keys, off, rows = bi.group_rows(ppi, "g")
    assert keys == ["x", "y"] and off.tolist() == [0, 2, 4] and rows.tolist() == [1, 3, 0, 2]


def test_context_tables_share_the_descriptor_of_equal_indexes():
    """Contexts with equal gene indexes share ONE ContextRows (and, in ``resolve``, one set of first / count rows and
    sub_rows): same index arrays, hence the same tensor, as resolving every context on its own."""
    mats, ppi = hard_case(3)
    mats = list(mats) + [mats[0].copy(), mats[1].iloc[::-1].copy(), mats[0].copy()]       # equal, permuted (not equal), equal
    complexes = bi.complex_table(ppi, "&", ("A", "B"))[4]
    shared = bi.context_tables([df.index for df in mats], complexes, True)
    each = [bi.context_rows(list(df.index), complexes, True) for df in mats]
    n0 = len(mats) - 3
    assert shared[n0] is shared[0] and shared[n0 + 2] is shared[0] and shared[n0 + 1] is not shared[1]
    for a, b in zip(shared, each):
        assert a.names == b.names and dict(a.rows) == dict(b.rows)
    names = [t.names for t in each]
    columns = [list(df.columns) for df in mats]
    genes = bi.ordered(names[0], bi.select_elements(names, True, 0.0))
    cells = bi.ordered(columns[0], bi.select_elements(columns, True, 0.0))
    f = bi.filter_lr_pairs(ppi, genes, "&", True, ("A", "B"))
    lig, rec = f["A"].tolist(), f["B"].tolist()
    ia = bi.resolve(shared, columns, genes, cells, lig, rec)
    ib = bi.resolve(each, columns, genes, cells, lig, rec)
    vals = [df.values for df in mats]
    Xa = emulate_build(vals, ia, len(mats), len(lig), len(cells), "expression_gmean", "min")
    Xb = emulate_build(vals, ib, len(mats), len(lig), len(cells), "expression_gmean", "min")
    np.testing.assert_array_equal(np.isnan(Xa), np.isnan(Xb))
    np.testing.assert_array_equal(np.nan_to_num(Xa), np.nan_to_num(Xb))
    assert len(ia["sub_rows"]) < len(ib["sub_rows"])
