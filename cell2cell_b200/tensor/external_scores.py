"""On-ramp from long-format communication-score tables (LIANA, MEBOCOST, ...) to a ``PreBuiltTensor``: same function,
signature and result as ``cell2cell/tensor/external_scores.py:12-215`` ("next" row 3).

The reference builds the tensor with one pandas filter + pivot + two reindexes per (context, LR pair).  Here the tables
are turned into integer coordinates once per context (``dataframe_coordinates``: the host stage, plain pandas / numpy, no
device needed) and the dense ``[context, LR pair, sender, receiver]`` array is filled on the device by three scatters:

* an LR pair absent from a context                       -> the whole matrix is ``lr_fill``             (external_scores.py:186-187)
* sender (receiver) that never appears for that LR pair
  in that context                                        -> its row (column) is ``cell_fill``           (:192, ``reindex(fill_value=)``)
* sender and receiver both appear but not as a pair      -> NaN (what ``DataFrame.pivot`` leaves)       (:191)
* duplicated (sender, receiver, ligand, receptor) rows   -> aggregated with ``dup_aggregation``         (:188-190)

Masks follow the reference: ``how != 'inner'`` -> ``mask = ~isnan``; ``how == 'inner'`` -> no mask (NaN positions, if any,
still end up in ``loc_nans`` through ``PreBuiltTensor``).
"""
import numpy as np
import pandas as pd
import torch

from .. import engine
from . import build_index as bi
from .tensor import PreBuiltTensor

__all__ = ["dataframes_to_tensor", "dataframe_coordinates"]

_HOWS = ("inner", "outer", "outer_lrs", "outer_cells")


def dataframe_coordinates(context_df_dict, sender_col, receiver_col, ligand_col, receptor_col, score_col, how="inner",
                          outer_fraction=0.0, lr_sep="^", dup_aggregation="max", context_order=None, sort_elements=True):
    """Host stage: element selection / ordering (external_scores.py:103-173) and integer coordinates of every kept entry.

    Returns ``(order_names, coords)``: ``order_names = [contexts, lr_pairs, senders, receivers]`` and ``coords`` = one tuple
    ``(l, s, r, value)`` of numpy arrays per context (entries whose LR pair, sender or receiver was not selected are
    dropped -- the reference drops them with its ``reindex``)."""
    if context_order is None:
        sort_context = sort_elements
        context_order = list(context_df_dict.keys())
    else:
        assert all([c in context_df_dict.keys() for c in context_order]), \
            "The list 'context_order' must contain all context names contained in the keys of 'context_df_dict'"
        assert len(context_order) == len(context_df_dict.keys()), \
            "Each context name must be contained only once in the list 'context_order'"
        sort_context = False
    cols = [sender_col, receiver_col, ligand_col, receptor_col, score_col]
    assert all([c in df.columns for c in cols for df in context_df_dict.values()]), \
        "One or more columns (among 'sender_col', 'receiver_col', 'ligand_col', 'receptor_col', and 'score_col') are " \
        "missing in at least one dataframe included in 'context_df_dict'"
    if how not in _HOWS:
        raise ValueError("Not a valid input for parameter 'how'")

    tables = {}
    for k, v in context_df_dict.items():
        df = v[cols].copy()
        df["LRs"] = df[ligand_col].astype(str) + lr_sep + df[receptor_col].astype(str)
        tables[k] = df
    df_lrs = [list(pd.unique(tables[k]["LRs"])) for k in context_order]
    df_senders = [list(pd.unique(tables[k][sender_col])) for k in context_order]
    df_receivers = [list(pd.unique(tables[k][receiver_col])) for k in context_order]

    def inner(lists):
        return list(set.intersection(*map(set, lists)))

    def outer(lists):
        return bi.select_elements(lists, True, outer_fraction)

    lr_pairs = outer(df_lrs) if how in ("outer", "outer_lrs") else inner(df_lrs)
    sender_cells = outer(df_senders) if how in ("outer", "outer_cells") else inner(df_senders)
    receiver_cells = outer(df_receivers) if how in ("outer", "outer_cells") else inner(df_receivers)
    if sort_elements:
        if sort_context:
            context_order = sorted(context_order)
        lr_pairs, sender_cells, receiver_cells = sorted(lr_pairs), sorted(sender_cells), sorted(receiver_cells)
    else:
        lr_pairs, sender_cells, receiver_cells = list(lr_pairs), list(sender_cells), list(receiver_cells)

    lr_pos = {n: i for i, n in enumerate(lr_pairs)}
    s_pos = {n: i for i, n in enumerate(sender_cells)}
    r_pos = {n: i for i, n in enumerate(receiver_cells)}
    coords = []
    for k in context_order:
        df = tables[k]
        keys = [sender_col, receiver_col, ligand_col, receptor_col]
        if df[keys].duplicated().any():
            assert dup_aggregation in ["max", "min", "mean", "median"], "Please use a valid option for `dup_aggregation`."
            df = getattr(df.groupby(keys + ["LRs"])[[score_col]], dup_aggregation)().reset_index()
        l = df["LRs"].map(lr_pos).to_numpy(dtype=np.float64, na_value=np.nan)
        keep_l = ~np.isnan(l)
        # presence of a sender / receiver for an LR pair is decided on ALL rows of that pair (the pivot), before the
        # selected cells are picked out by the reindex
        s_all = df[sender_col].map(s_pos).to_numpy(dtype=np.float64, na_value=np.nan)
        r_all = df[receiver_col].map(r_pos).to_numpy(dtype=np.float64, na_value=np.nan)
        vals = df[score_col].to_numpy(dtype=np.float64)
        coords.append((l[keep_l].astype(np.int64), s_all[keep_l], r_all[keep_l], vals[keep_l]))
    return [list(context_order), lr_pairs, sender_cells, receiver_cells], coords


def dataframes_to_tensor(context_df_dict, sender_col, receiver_col, ligand_col, receptor_col, score_col, how="inner",
                         outer_fraction=0.0, lr_fill=np.nan, cell_fill=np.nan, lr_sep="^", dup_aggregation="max",
                         context_order=None, order_labels=None, sort_elements=True, device=None):
    """Generates a ``PreBuiltTensor`` (CUDA) from a dictionary of long-format dataframes, one per context."""
    dev = engine.require_cuda(device)
    names, coords = dataframe_coordinates(context_df_dict, sender_col, receiver_col, ligand_col, receptor_col, score_col,
                                          how=how, outer_fraction=outer_fraction, lr_sep=lr_sep,
                                          dup_aggregation=dup_aggregation, context_order=context_order,
                                          sort_elements=sort_elements)
    if order_labels is None:
        order_labels = ["Contexts", "Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
    C, L, S, R = (len(n) for n in names)
    with torch.cuda.device(dev):
        tensor = torch.empty((C, L, S, R), dtype=torch.float64, device=dev)
        for c, (l, s, r, v) in enumerate(coords):
            lt = torch.as_tensor(l, device=dev)
            lr_present = torch.zeros(L, dtype=torch.bool, device=dev)
            lr_present[lt] = True
            s_ok, r_ok = ~np.isnan(s), ~np.isnan(r)
            s_seen = torch.zeros((L, S), dtype=torch.bool, device=dev)
            r_seen = torch.zeros((L, R), dtype=torch.bool, device=dev)
            s_seen[torch.as_tensor(l[s_ok], device=dev), torch.as_tensor(s[s_ok].astype(np.int64), device=dev)] = True
            r_seen[torch.as_tensor(l[r_ok], device=dev), torch.as_tensor(r[r_ok].astype(np.int64), device=dev)] = True
            both = s_seen[:, :, None] & r_seen[:, None, :]
            plane = torch.where(both, torch.full((), float("nan"), dtype=torch.float64, device=dev),
                                torch.full((), float(cell_fill), dtype=torch.float64, device=dev))
            plane = torch.where(lr_present[:, None, None], plane, torch.full((), float(lr_fill), dtype=torch.float64, device=dev))
            ok = s_ok & r_ok
            plane[torch.as_tensor(l[ok], device=dev), torch.as_tensor(s[ok].astype(np.int64), device=dev),
                  torch.as_tensor(r[ok].astype(np.int64), device=dev)] = torch.as_tensor(v[ok], device=dev)
            tensor[c] = plane
        if how != "inner":
            nan = torch.isnan(tensor)
            mask = (~nan).to(torch.uint8)
            loc_nans = nan.to(torch.uint8)
        else:
            mask = None
            loc_nans = torch.zeros(tensor.shape, dtype=torch.uint8, device=dev)
    return PreBuiltTensor(tensor=tensor, order_names=names, order_labels=order_labels, mask=mask, loc_nans=loc_nans.cpu().numpy(),
                          device=dev)
