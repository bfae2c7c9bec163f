"""Index utilities on the tensor objects: same functions and behaviour as ``cell2cell/tensor/subset.py:9-221`` and
``cell2cell/tensor/tensor_manipulation.py:9-107`` ("next" row 3), on CUDA tensors: the gathers are
``index_select`` / ``cat`` on the device instead of a round trip through numpy.

One reference quirk is NOT kept: ``subset_tensor`` there subsets ``tensor``, ``mask`` and ``order_names`` but leaves the
``loc_nans`` / ``loc_zeros`` arrays at their old shape (subset.py:137-178); here they are subset too, so
``sparsity_fraction`` / ``missing_fraction`` stay usable on the result.
"""
import copy

import torch

from .tensor import PreBuiltTensor

__all__ = ["find_element_indexes", "subset_tensor", "subset_metadata", "concatenate_interaction_tensors"]


def _find_duplicates(names):
    """cell2cell/preprocessing/find_elements.py:8-29: {element: [positions]} for elements occurring more than once."""
    pos = {}
    for i, n in enumerate(names):
        pos.setdefault(n, []).append(i)
    return {k: v for k, v in pos.items() if len(v) > 1}


def find_element_indexes(interaction_tensor, elements, axis=0, remove_duplicates=True, keep="first", original_order=False):
    """Positions of ``elements`` along ``axis`` (subset.py:9-85)."""
    assert axis < len(interaction_tensor.tensor.shape), "List index out of range. 'axis' must be one of the axis in the tensor."
    assert axis < len(interaction_tensor.order_names), \
        "List index out of range. interaction_tensor.order_names must have element names for each axis of the tensor."
    names = list(interaction_tensor.order_names[axis])
    elements = sorted(set(elements), key=list(elements).index)
    if original_order:
        elements = set(elements).intersection(set(names))
        elements = sorted(elements, key=names.index)
    to_exclude = []
    if remove_duplicates:
        dup = _find_duplicates(names)
        if len(dup) > 0:
            if keep == "first":
                for v in dup.values():
                    to_exclude.extend(v[1:])
            elif keep == "last":
                for v in dup.values():
                    to_exclude.extend(v[:-1])
            elif not keep:
                for v in dup.values():
                    to_exclude.extend(v)
            else:
                raise ValueError("Not a valid option was selected for the parameter `keep`")
    where = {}
    for i, n in enumerate(names):
        where.setdefault(n, []).append(i)
    indexes = sum([where.get(e, []) for e in elements], [])
    excl = set(to_exclude)
    return [i for i in indexes if i not in excl]


def subset_tensor(interaction_tensor, subset_dict, remove_duplicates=True, keep="first", original_order=False):
    """Copy of ``interaction_tensor`` restricted to the elements of ``subset_dict = {axis: [names]}`` (subset.py:88-181);
    the previous factorization is reset."""
    sub = copy.deepcopy(interaction_tensor)
    sub.rank = None
    sub.tl_object = None
    sub.factors = None
    tensor, mask = sub.tensor, sub.mask
    extras = {k: getattr(sub, k, None) for k in ("_loc_nans", "_loc_zeros")}
    axis_idxs = {}
    for k, v in subset_dict.items():
        if k < tensor.dim():
            if len(v) != 0:
                idx = find_element_indexes(sub, v, axis=k, remove_duplicates=remove_duplicates, keep=keep,
                                           original_order=original_order)
                if len(idx) == 0:
                    print("No elements found for axis {}. It will return an empty tensor.".format(k))
                axis_idxs[k] = idx
        else:
            print("Axis {} is out of index, not considering elements in this axis.".format(k))
    sub.order_names = [list(n) for n in sub.order_names]
    for k, v in axis_idxs.items():
        idx = torch.as_tensor(v, dtype=torch.long, device=tensor.device)
        tensor = tensor.index_select(k, idx)
        sub.order_names[k] = [sub.order_names[k][i] for i in v]
        if mask is not None:
            mask = mask.index_select(k, idx)
        for name, arr in extras.items():
            if isinstance(arr, torch.Tensor):
                extras[name] = arr.index_select(k, idx)
    sub.tensor = tensor.contiguous()
    sub.mask = None if mask is None else mask.contiguous()
    for name, arr in extras.items():
        setattr(sub, name, arr.contiguous() if isinstance(arr, torch.Tensor) else arr)
    return sub


def subset_metadata(tensor_metadata, interaction_tensor, sample_col="Element"):
    """Metadata rows of the elements still in ``interaction_tensor``, in its order (subset.py:184-221)."""
    out = []
    for i, meta in enumerate(tensor_metadata):
        if meta is not None:
            tmp = meta.set_index(sample_col).loc[interaction_tensor.order_names[i], :].reset_index()
            out.append(tmp)
        else:
            out.append(None)
    return out


def concatenate_interaction_tensors(interaction_tensors, axis, order_labels, remove_duplicates=False, keep="first", mask=None,
                                    device=None):
    """Concatenation of tensors that share every other axis (tensor_manipulation.py:9-107); a mask is kept only when every
    input has one (or one is given)."""
    ndim = len(interaction_tensors[0].tensor.shape)
    assert all(ndim == len(t.tensor.shape) for t in interaction_tensors[1:]), "Tensors must have same number of dimensions"
    for i in range(ndim):
        if i != axis:
            elements = list(interaction_tensors[0].order_names[i])
            for t in interaction_tensors[1:]:
                assert elements == list(t.order_names[i]), "Tensors must have the same elements in the other axes."
    dev = interaction_tensors[0].tensor.device if device is None else torch.device(device)
    concat = torch.cat([t.tensor.to(dev) for t in interaction_tensors], dim=axis)
    if mask is not None:
        assert tuple(mask.shape) == tuple(concat.shape), \
            "Mask must have the same shape of the concatenated tensor. Here: {}".format(tuple(concat.shape))
    elif all(t.mask is not None for t in interaction_tensors):
        mask = torch.cat([t.mask.to(dev) for t in interaction_tensors], dim=axis)
    order_names = []
    for i in range(ndim):
        if i == axis:
            names = []
            for t in interaction_tensors:
                names += list(t.order_names[i])
        else:
            names = list(interaction_tensors[0].order_names[i])
        order_names.append(names)
    out = PreBuiltTensor(tensor=concat, order_names=order_names, order_labels=order_labels, mask=mask, device=dev)
    if remove_duplicates:
        out = subset_tensor(out, {axis: order_names[axis]}, remove_duplicates=remove_duplicates, keep=keep, original_order=False)
    return out
