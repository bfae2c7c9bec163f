"""CPU oracle for the tensor-build stage: a plain numpy/pandas (float64) restatement of the reference's
``InteractionTensor`` build.

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, never by the product package ``cell2cell_b200``.

PARITY PINNED: ``tests/golden/build_*.npz`` hold outputs of the reference's own code (``cell2cell/tensor/tensor.py``,
imported through ``oracle/ref_harness.py`` in the build container) for the toy data of
``cell2cell/datasets/toy_data.py:5-76`` and for seeded random cases over every ``how`` x ``communication_score`` x
``complex_agg_method`` x ``group_ppi_method`` option; ``tests/test_build_oracle.py`` checks this restatement against
them bit for bit (float64 values, int64 mask, names), and -- when /root/reference is present -- against the live
reference on fresh random inputs.

Each function cites the reference lines it follows.
"""
import warnings
from collections import Counter, OrderedDict

import numpy as np
import pandas as pd

__all__ = ["complexes_from_ppi", "add_complex_rows", "filter_ppi", "score_matrix", "aggregate_group",
           "build_context_ccc_tensor_oracle", "interaction_tensor_oracle"]


def complexes_from_ppi(ppi_data, complex_sep, interaction_columns=("A", "B")):
    """``get_genes_from_complexes`` (cell2cell/preprocessing/ppi.py:370-444): dict complex name -> set(subunits), plus
    the plain names and the subunit unions of either column."""
    col_a, col_b = interaction_columns
    plain = {col_a: set(), col_b: set()}
    subunits = {col_a: set(), col_b: set()}
    complexes = OrderedDict()
    for a, b in zip(ppi_data[col_a].tolist(), ppi_data[col_b].tolist()):
        for col, name in ((col_a, a), (col_b, b)):
            if complex_sep in name:
                parts = set(name.split(complex_sep))
                complexes[name] = parts
                subunits[col] |= parts
            else:
                plain[col].add(name)
    return plain[col_a], subunits[col_a], plain[col_b], subunits[col_b], complexes


def add_complex_rows(rnaseq, complexes, agg_method="min"):
    """``add_complexes_to_expression`` (cell2cell/preprocessing/rnaseq.py:144-196): one new row per complex; all
    subunits present -> min / mean (NaN-skipping, pandas default) / exp(mean(log)) over the subunit rows, otherwise a row
    of zeros (``:195``).  Rows are appended one at a time, so an earlier complex row is visible to a later one."""
    out = rnaseq.copy()
    for name, parts in complexes.items():
        parts = list(parts)
        if all(p in out.index for p in parts):
            sub = out.loc[parts, :]
            if agg_method == "min":
                out.loc[name] = sub.min().values.tolist()
            elif agg_method == "mean":
                out.loc[name] = sub.mean().values.tolist()
            elif agg_method == "gmean":
                out.loc[name] = sub.apply(lambda col: np.exp(np.mean(np.log(col)))).values.tolist()
            # any other string: the reference builds a ValueError without raising it (:193) -> the row is not added
        else:
            out.loc[name] = [0] * out.shape[1]
    return out


def filter_ppi(ppi_data, proteins, complex_sep=None, upper_letter_comparison=True, interaction_columns=("A", "B")):
    """``filter_ppi_by_proteins`` (ppi.py:185-244) -> ``filter_complex_ppi_by_proteins`` (ppi.py:447-523), including the
    ``elif`` at :511-515 (a complex that qualifies on side A is never added to side B)."""
    col_a, col_b = interaction_columns
    ppi = ppi_data.copy()
    if upper_letter_comparison:
        ppi[col_a] = [str(x).upper() for x in ppi[col_a]]
        ppi[col_b] = [str(x).upper() for x in ppi[col_b]]
        prots = set(str(p).upper() for p in proteins)
    else:
        prots = set(proteins)
    if complex_sep is None:
        keep = ppi[col_a].isin(prots) & ppi[col_b].isin(prots)
        return ppi[keep].reset_index(drop=True)
    plain_a, sub_a, plain_b, sub_b, complexes = complexes_from_ppi(ppi, complex_sep, interaction_columns)
    ok_a = set(plain_a & prots)
    ok_b = set(plain_b & prots)
    sub_ok_a = sub_a & prots
    sub_ok_b = sub_b & prots
    for name, parts in complexes.items():
        if all(p in sub_ok_a for p in parts):
            ok_a.add(name)
        elif all(p in sub_ok_b for p in parts):
            ok_b.add(name)
    keep = ppi[col_a].isin(ok_a) & ppi[col_b].isin(ok_b)
    return ppi.loc[keep].reset_index(drop=True)


def score_matrix(v, w, communication_score):
    """``compute_ccc_matrix`` (cell2cell/core/communication_scores.py:159-203); rows = senders, cols = receivers."""
    v = np.asarray(v)
    w = np.asarray(w)
    if communication_score == "expression_product":
        return v[:, None] * w[None, :]
    if communication_score == "expression_mean":
        return (v[:, None] * np.ones(w.shape)[None, :] + np.ones(v.shape)[:, None] * w[None, :]) / 2.0
    if communication_score == "expression_gmean":
        with np.errstate(invalid="ignore"):
            return np.sqrt(v[:, None] * w[None, :])
    raise ValueError("Not a valid communication_score")


def aggregate_group(mats, method):
    """``aggregate_ccc_matrices`` (communication_scores.py:206-244).  'gmean' is ``scipy.stats.mstats.gmean``; with the
    scipy the reference runs on here (1.18.1, the version the golden fixtures were made with) that is a plain
    ``exp(mean(log(x)))`` along axis 0: a zero in the group gives 0, a NaN or a negative value gives NaN (older scipy
    releases masked such entries instead -- the fixtures pin the behaviour restated here).  'sum' / 'mean' are
    ``np.nansum`` / ``np.nanmean`` (an all-NaN group sums to 0 and averages to NaN)."""
    mats = np.asarray(mats, dtype=float)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        if method == "gmean":
            from scipy.stats.mstats import gmean   # same third-party call as the reference (its last bit depends on scipy's mean)
            return np.asarray(gmean(mats))
        if method == "sum":
            return np.nansum(mats, axis=0)
        if method == "mean":
            return np.nanmean(mats, axis=0)
    raise ValueError("Not a valid method")


def _universe(lists, how_outer, fraction):
    """``get_element_abundances`` + ``get_elements_over_fraction`` (find_elements.py:32-76) or plain intersection."""
    if not how_outer:
        return set.intersection(*map(set, lists))
    counts = Counter()
    for l in lists:
        counts.update(set(l))
    total = len(lists)
    return set(k for k, c in counts.items() if c / total >= fraction)


def build_context_ccc_tensor_oracle(rnaseq_matrices, ppi_data, how="inner", outer_fraction=0.0,
                                    communication_score="expression_product", complex_sep=None,
                                    upper_letter_comparison=True, interaction_columns=("A", "B"), group_ppi_by=None,
                                    group_ppi_method="gmean"):
    """``build_context_ccc_tensor`` (cell2cell/tensor/tensor.py:973-1146) with ``generate_ccc_tensor`` (:1149-1202) and
    ``aggregate_ccc_tensor`` (:1205-1252) inlined.  Returns (tensor float64/int64, genes, cells, ppi_names, mask)."""
    idxs = [list(df.index) for df in rnaseq_matrices]
    cols = [list(df.columns) for df in rnaseq_matrices]
    if how not in ("inner", "outer", "outer_genes", "outer_cells"):
        raise ValueError('Provide a valid way to build the tensor; "how" must be "inner", "outer", "outer_genes" or "outer_cells"')
    genes = _universe(idxs, how in ("outer", "outer_genes"), outer_fraction)
    cells = _universe(cols, how in ("outer", "outer_cells"), outer_fraction)
    genes = idxs[0] if set(idxs[0]) == genes else sorted(genes)      # :1106-1109
    cells = cols[0] if set(cols[0]) == cells else sorted(cells)      # :1111-1114
    ppi = filter_ppi(ppi_data, genes, complex_sep, upper_letter_comparison, interaction_columns)
    col_a, col_b = interaction_columns
    lig = ppi[col_a].tolist()
    rec = ppi[col_b].tolist()
    per_context = []
    for df in rnaseq_matrices:
        expr = df.reindex(genes).reindex(cells, axis="columns")       # :1126  (NaN where absent)
        mats = []
        for a, b in zip(lig, rec):
            mats.append(score_matrix(expr.loc[a, :].values, expr.loc[b, :].values, communication_score).tolist())
        per_context.append(mats)
    if group_ppi_by is not None:
        groups = [(g, list(sub.index)) for g, sub in ppi.groupby(group_ppi_by)]
        ppi_names = [g for g, _ in groups]
        grouped = []
        for mats in per_context:
            arr = np.array(mats)
            grouped.append([aggregate_group(arr[rows], group_ppi_method).tolist() for _, rows in groups])
        per_context = grouped
    else:
        ppi_names = [a + "^" + b for a, b in zip(lig, rec)]
    if how != "inner":
        mask = (~np.isnan(np.asarray(per_context, dtype=float))).astype(int) if group_ppi_by is not None \
            else (~np.isnan(np.asarray(per_context))).astype(int)
    else:
        mask = None
    if group_ppi_by is not None:
        tensor = np.nan_to_num(np.asarray(per_context, dtype=float))
    else:
        tensor = np.nan_to_num(per_context)
    return tensor, genes, cells, ppi_names, mask


def interaction_tensor_oracle(rnaseq_matrices, ppi_data, context_names=None, how="inner", outer_fraction=0.0,
                              communication_score="expression_mean", complex_sep=None, complex_agg_method="min",
                              upper_letter_comparison=True, interaction_columns=("A", "B"), group_ppi_by=None,
                              group_ppi_method="gmean"):
    """``InteractionTensor.__init__`` (cell2cell/tensor/tensor.py:796-883).  Returns a dict with tensor, mask, loc_nans,
    loc_zeros, genes, cells, order_names."""
    if complex_sep is not None:
        complexes = complexes_from_ppi(ppi_data, complex_sep, interaction_columns)[4]   # ORIGINAL-case names (:812-816)
        mats = [add_complex_rows(df, complexes, complex_agg_method) for df in rnaseq_matrices]
    else:
        mats = [df.copy() for df in rnaseq_matrices]
    if upper_letter_comparison:
        for df in mats:
            df.index = [i.upper() for i in df.index]                                     # :821-823
    mats = [df[~df.index.duplicated(keep="first")] for df in mats]                       # :826
    tensor, genes, cells, ppi_names, mask = build_context_ccc_tensor_oracle(
        mats, ppi_data, how, outer_fraction, communication_score, complex_sep, upper_letter_comparison,
        interaction_columns, group_ppi_by, group_ppi_method)
    loc_nans = np.zeros(tensor.shape, dtype=int) if mask is None else 1 - np.array(mask)
    loc_zeros = (((tensor == 0).astype(int) - loc_nans) > 0).astype(int)                  # :846-847
    if context_names is None:
        context_names = ["C-" + str(i) for i in range(1, len(mats) + 1)]
    return dict(tensor=tensor, mask=mask, loc_nans=loc_nans, loc_zeros=loc_zeros, genes=genes, cells=cells,
                order_names=[context_names, ppi_names, cells, cells])
