"""``dataframes_to_tensor`` (cell2cell/tensor/external_scores.py:12-215): the host stage against the golden names produced by
the reference's own code (CPU), the whole function against its golden tensors / masks (GPU)."""
import json
import os

import numpy as np
import pytest

from tests.cases import LONG_TABLE_OPTIONS, long_tables

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataframes.npz")
COLS = ("source", "target", "ligand", "receptor", "score")


def _cases():
    return [(seed, i) for seed in (0, 1) for i in range(len(LONG_TABLE_OPTIONS))]


@pytest.mark.parametrize("seed,i", _cases())
def test_names_and_coordinates_match_reference(seed, i):
    from cell2cell_b200.tensor.external_scores import dataframe_coordinates
    gold = np.load(GOLD)
    kw = dict(LONG_TABLE_OPTIONS[i])
    kw.pop("lr_fill"); kw.pop("cell_fill")
    names, coords = dataframe_coordinates(long_tables(seed), *COLS, **kw)
    want = json.loads(str(gold["%d|%d/names" % (seed, i)]))
    assert names == want
    ref = gold["%d|%d/tensor" % (seed, i)]
    # every kept coordinate carries the reference's value (duplicates already aggregated)
    for c, (l, s, r, v) in enumerate(coords):
        ok = ~np.isnan(s) & ~np.isnan(r)
        got = ref[c, l[ok], s[ok].astype(int), r[ok].astype(int)]
        np.testing.assert_array_equal(got, v[ok])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,i", _cases())
def test_dataframes_to_tensor_matches_reference(seed, i):
    torch = pytest.importorskip("torch")
    from cell2cell_b200.tensor.external_scores import dataframes_to_tensor
    gold = np.load(GOLD)
    key = "%d|%d" % (seed, i)
    t = dataframes_to_tensor(long_tables(seed), *COLS, **LONG_TABLE_OPTIONS[i])
    assert [list(n) for n in t.order_names] == json.loads(str(gold[key + "/names"]))
    ref = gold[key + "/tensor"]
    got = t.tensor.cpu().numpy()
    np.testing.assert_array_equal(got, np.nan_to_num(ref).astype(np.float32))     # scores are 4-decimal numbers: exact in fp32 rounding
    if bool(gold[key + "/has_mask"]):
        assert np.array_equal(t.mask.cpu().numpy(), gold[key + "/mask"])
    else:
        assert t.mask is None
    assert np.array_equal(t.loc_nans.cpu().numpy(), gold[key + "/loc_nans"])
    assert np.array_equal(t.loc_zeros.cpu().numpy(), gold[key + "/loc_zeros"])


@pytest.mark.gpu
def test_subset_and_concatenate_follow_reference_semantics():
    """subset.py:88-181 / tensor_manipulation.py:9-107 on device tensors: same element selection, order and duplicate rules
    as numpy ``take`` / ``concatenate`` on the host copies."""
    torch = pytest.importorskip("torch")
    from cell2cell_b200.tensor import PreBuiltTensor, concatenate_interaction_tensors, subset_tensor
    rs = np.random.RandomState(0)
    x = rs.rand(3, 6, 4, 4).astype(np.float32)
    x[0, 1, 2, :] = np.nan
    m = (rs.rand(3, 6, 4, 4) > 0.2).astype(np.uint8)
    names = [["c1", "c2", "c3"], ["lr%d" % i for i in range(6)], list("ABCD"), list("ABCD")]
    t = PreBuiltTensor(x, order_names=names, mask=m)
    sub = subset_tensor(t, {0: ["c3", "c1", "zz"], 1: ["lr4", "lr0"], 3: []})
    assert sub.order_names[0] == ["c3", "c1"] and sub.order_names[1] == ["lr4", "lr0"] and sub.order_names[3] == list("ABCD")
    want = np.nan_to_num(x).take([2, 0], axis=0).take([4, 0], axis=1)
    assert np.array_equal(sub.tensor.cpu().numpy(), want)
    assert np.array_equal(sub.mask.cpu().numpy(), m.take([2, 0], axis=0).take([4, 0], axis=1))
    assert tuple(sub.loc_nans.shape) == tuple(sub.tensor.shape) and sub.tl_object is None and sub.rank is None
    ordered = subset_tensor(t, {0: ["c3", "c1"]}, original_order=True)
    assert ordered.order_names[0] == ["c1", "c3"]
    # concatenation along the context axis with a repeated context name: duplicates dropped on request (first kept)
    t2 = PreBuiltTensor(x[:2] * 2.0, order_names=[["c3", "c9"]] + names[1:], mask=m[:2])
    cat = concatenate_interaction_tensors([t, t2], axis=0, order_labels=["a", "b", "c", "d"], remove_duplicates=True)
    assert cat.order_names[0] == ["c1", "c2", "c3", "c9"]
    want = np.concatenate([np.nan_to_num(x), np.nan_to_num(x[:2] * 2.0)], axis=0).take([0, 1, 2, 4], axis=0)
    assert np.array_equal(cat.tensor.cpu().numpy(), want)
    assert np.array_equal(cat.mask.cpu().numpy(), np.concatenate([m, m[:2]], axis=0).take([0, 1, 2, 4], axis=0))
    plain = concatenate_interaction_tensors([t, PreBuiltTensor(x, order_names=names)], axis=1, order_labels=["a", "b", "c", "d"])
    assert plain.mask is None and plain.tensor.shape[1] == 12


def _tiled_template(dist, shape, source_axis, target_axis):
    """The reference's own template construction (spatial/filtering.py:52-70), kept in numpy as the checker."""
    original_order = list(range(len(shape)))
    new_order = []
    template = dist
    for i, size in enumerate(shape):
        if i != source_axis and i != target_axis:
            template = [template] * size
            new_order.insert(0, i)
    template = np.array(template)
    new_order += [source_axis, target_axis]
    return template.transpose([new_order.index(i) for i in original_order])


@pytest.mark.gpu
@pytest.mark.parametrize("source_axis,target_axis", [(2, 3), (3, 2)])
def test_dist_filter_tensor_follows_reference(source_axis, target_axis):
    """spatial/filtering.py:6-82: scores of cell pairs outside [min_dist, max_dist] become real zeros."""
    pd = pytest.importorskip("pandas")
    from cell2cell_b200.spatial import dist_filter_tensor
    from cell2cell_b200.tensor import PreBuiltTensor
    rs = np.random.RandomState(3)
    cells = list("ABCDE")
    x = rs.rand(3, 7, 5, 5).astype(np.float32)
    x[1, 2, 3, :] = np.nan
    x[0, 0, 0, 0] = 0.0
    names = [["c1", "c2", "c3"], ["lr%d" % i for i in range(7)], cells, cells]
    t = PreBuiltTensor(x, order_names=names)
    d = rs.rand(6, 6) * 10
    distances = pd.DataFrame(d, index=list("FEDCBA"), columns=list("BFACED"))      # other order, one extra group
    out = dist_filter_tensor(t, distances, max_dist=6.0, min_dist=1.0, source_axis=source_axis, target_axis=target_axis)
    sub = distances.loc[cells, cells]
    keep = ((1.0 <= sub) & (sub <= 6.0)).astype(int).values
    want = np.nan_to_num(x) * _tiled_template(keep, x.shape, source_axis, target_axis)
    assert np.array_equal(out.tensor.cpu().numpy(), want.astype(np.float32))
    loc_nans = np.isnan(x).astype(int)
    assert np.array_equal(out.loc_nans.cpu().numpy(), loc_nans)
    assert np.array_equal(out.loc_zeros.cpu().numpy().astype(int), ((want == 0).astype(int) - loc_nans > 0).astype(int))
    assert np.array_equal(t.tensor.cpu().numpy(), np.nan_to_num(x))               # the input object is untouched
    with pytest.raises(AssertionError):
        dist_filter_tensor(t, distances.drop(index="A"), max_dist=6.0)


@pytest.mark.gpu
def test_interactions_to_tensor_equals_direct_build():
    """tensor.py:1318-1419: per-context pipeline objects -> one InteractionTensor over the union of their PPI lists."""
    from types import SimpleNamespace
    from cell2cell_b200.synthetic import synthetic_contexts, synthetic_ppi
    from cell2cell_b200.tensor import InteractionTensor, interactions_to_tensor
    ppi = synthetic_ppi(40, 80)
    contexts = synthetic_contexts(3, 80, 5, outer=True)
    pipes = []
    for c, m in enumerate(contexts):
        part = ppi.iloc[c * 10:(c + 2) * 10]                                        # overlapping slices: duplicates dropped
        pipes.append(SimpleNamespace(complex_sep="&", complex_agg_method="min", interaction_columns=("A", "B"),
                                     analysis_setup={"cci_type": "directed" if c else "undirected"},
                                     ppi_data=part, ref_ppi=part, aggregated_expression=m, rnaseq_data=m))
    got = interactions_to_tensor(pipes, experiment="single_cell", context_names=["a", "b", "c"], how="outer",
                                 communication_score="expression_gmean", verbose=False)
    import pandas as pd
    merged = pd.concat([p.ppi_data for p in pipes]).drop_duplicates().reset_index(drop=True)
    want = InteractionTensor(contexts, merged, context_names=["a", "b", "c"], how="outer", complex_sep="&",
                             communication_score="expression_gmean", verbose=False)
    assert got.order_names == want.order_names and len(got.order_names[1]) == len(merged)
    assert np.array_equal(got.tensor.cpu().numpy(), want.tensor.cpu().numpy())
    assert np.array_equal(got.mask.cpu().numpy(), want.mask.cpu().numpy())
    with pytest.raises(ValueError):
        interactions_to_tensor(pipes, experiment="spatial")
