"""Distance filter for communication tensors: same function and behaviour as ``cell2cell/spatial/filtering.py:6-82``
("next" row 3), on CUDA tensors.

The reference tiles the K x K 0/1 keep-matrix into a full tensor-sized template on the host and multiplies
(filtering.py:56-76).  Here the keep-matrix goes to the device once and is broadcast over the other modes inside the
multiply; nothing tensor-sized is built besides the result.
"""
import numpy as np
import torch

__all__ = ["dist_filter_tensor"]


def dist_filter_tensor(interaction_tensor, distances, max_dist, min_dist=0, source_axis=2, target_axis=3):
    """Copy of ``interaction_tensor`` with the scores of (sender, receiver) pairs whose distance lies outside
    ``[min_dist, max_dist]`` made real zeros (filtering.py:6-82).

    ``distances`` is a square DataFrame holding every sender (index) and receiver (columns) of the tensor.
    """
    assert all([cell in distances.index for cell in
                interaction_tensor.order_names[source_axis]]), "Distances not provided for all sender cells"
    assert all([cell in distances.columns for cell in
                interaction_tensor.order_names[target_axis]]), "Distances not provided for all receiver cells"
    assert source_axis != target_axis, "source_axis and target_axis must differ"
    dist_df = distances.loc[interaction_tensor.order_names[source_axis], interaction_tensor.order_names[target_axis]]
    keep = ((min_dist <= dist_df) & (dist_df <= max_dist)).values                   # filtering.py:49-50
    new = interaction_tensor.copy()
    X = new.tensor
    ndim = X.dim()
    src, tgt = source_axis % ndim, target_axis % ndim
    view = [1] * ndim
    view[src], view[tgt] = keep.shape
    k = torch.as_tensor(np.ascontiguousarray(keep if src < tgt else keep.T), device=X.device)
    new.tensor = X * k.view(view).to(X.dtype)         # a new tensor object: the cached ||X||^2 of X is not carried over
    # pairs removed by distance are real zeros, not missing values (filtering.py:80)
    new.loc_zeros = ((new.tensor == 0) & (new.loc_nans == 0)).to(torch.uint8)
    return new
