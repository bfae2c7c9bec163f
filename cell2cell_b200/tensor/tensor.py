"""Tensor objects of Tensor-cell2cell on B200: same classes, signatures and attributes as ``cell2cell/tensor/tensor.py``.

``InteractionTensor`` / ``build_context_ccc_tensor`` build the 4-way (context x LR pair x sender x receiver) tensor and
its missing-value mask on the device (host: name logic in ``build_index``; device: ``c2c_build_ccc`` /
``c2c_group_aggregate``); ``BaseTensor.compute_tensor_factorization`` and ``BaseTensor.elbow_rank_selection`` drive the
device NCP.  The tensor lives on the GPU as fp32 (what the reference's own GPU path -- tensorly's pytorch backend --
holds: docs GPU-Example prints ``dtype: float32``), the mask as uint8; ``loc_nans`` / ``loc_zeros`` are derived on first
use instead of being stored as two more full-size int64 arrays.  There is no CPU fallback: without a CUDA device or the
built library every constructor raises (contrast cell2cell/tensor/tensor.py:205-214).
"""
import copy
import pickle
from collections import OrderedDict

import numpy as np
import pandas as pd
import torch
from tqdm.auto import tqdm

from .. import engine
from . import build_index as bi
from .cp import CPTensor
from .elbow import smooth_curve
from .factorization import (_compute_elbow, _compute_norm_error, _compute_tensor_factorization,
                            _multiple_runs_elbow_analysis, _run_elbow_analysis, as_device_mask, as_device_tensor)

__all__ = ["BaseTensor", "InteractionTensor", "PreBuiltTensor", "build_context_ccc_tensor", "generate_tensor_metadata",
           "interactions_to_tensor"]

_SCORES = ("expression_product", "expression_mean", "expression_gmean")


class BaseTensor:
    """Base class holding the factorization API (cell2cell/tensor/tensor.py:20-688)."""

    def __init__(self):
        self.communication_score = None
        self.how = None
        self.outer_fraction = None
        self.tensor = None
        self.genes = None
        self.cells = None
        self.order_names = [None, None, None, None]
        self.order_labels = None
        self.tl_object = None
        self.norm_tl_object = None
        self.factors = None
        self.rank = None
        self.mask = None
        self.explained_variance_ = None
        self.explained_variance_ratio_ = None
        self._loc_nans = None
        self._loc_zeros = None
        self.elbow_metric = None
        self.elbow_metric_mean = None
        self.elbow_metric_raw = None

    # ---- derived full-size arrays ------------------------------------------------------------------------------------
    @property
    def loc_nans(self):
        """uint8 CUDA tensor, 1 where a NaN was assigned when building the tensor (tensor.py:842-845)."""
        if self._loc_nans is None and self.tensor is not None:
            if self.mask is None:
                self._loc_nans = torch.zeros(self.tensor.shape, dtype=torch.uint8, device=self.tensor.device)
            else:
                self._loc_nans = (1 - self.mask).to(torch.uint8)
        return self._loc_nans

    @loc_nans.setter
    def loc_nans(self, value):
        self._loc_nans = value

    @property
    def loc_zeros(self):
        """uint8 CUDA tensor, 1 at real zeros (zeros that are not NaN positions) (tensor.py:846-847)."""
        if self._loc_zeros is None and self.tensor is not None:
            self._loc_zeros = ((self.tensor == 0) & (self.loc_nans == 0)).to(torch.uint8)
        return self._loc_zeros

    @loc_zeros.setter
    def loc_zeros(self, value):
        self._loc_zeros = value

    def copy(self):
        """Deep copy of this object."""
        return copy.deepcopy(self)

    @property
    def shape(self):
        return tuple(self.tensor.shape) if hasattr(self.tensor, "shape") else ()

    def __getstate__(self):
        # pickled like the reference's object (io/save_data.py:8-28), with device tensors as host numpy arrays
        state = dict(self.__dict__)
        for k in ("tensor", "mask", "_loc_nans", "_loc_zeros"):
            v = state.get(k)
            if isinstance(v, torch.Tensor):
                state[k] = v.detach().cpu().numpy()
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if isinstance(self.tensor, np.ndarray) and torch.cuda.is_available():
            self.tensor = as_device_tensor(self.tensor)
            if isinstance(self.mask, np.ndarray):
                self.mask = as_device_mask(self.mask, self.tensor)
            for k in ("_loc_nans", "_loc_zeros"):
                v = getattr(self, k, None)
                if isinstance(v, np.ndarray):
                    setattr(self, k, torch.as_tensor(v).to(self.tensor.device))

    def write_file(self, filename):
        """Exports this object into a pickle file (tensor.py:181-192)."""
        with open(filename, "wb") as fh:
            pickle.dump(self, fh, protocol=pickle.HIGHEST_PROTOCOL)
        print(filename, " was correctly saved.")

    def to_device(self, device):
        """Moves the tensor (and mask) to another CUDA device; raises for anything else (tensor.py:194-214 falls back to
        the CPU with a print -- this implementation has no CPU path)."""
        dev = engine.require_cuda(device)
        self.tensor = self.tensor.to(dev)
        if self.mask is not None:
            self.mask = self.mask.to(dev)
        self._loc_nans = None if self._loc_nans is None else self._loc_nans.to(dev)
        self._loc_zeros = None if self._loc_zeros is None else self._loc_zeros.to(dev)

    # ---- factorization ----------------------------------------------------------------------------------------------
    def compute_tensor_factorization(self, rank, tf_type="non_negative_cp", init="svd", svd="truncated_svd",
                                     random_state=None, runs=1, normalize_loadings=True, var_ordered_factors=True,
                                     n_iter_max=100, tol=10e-7, verbose=False, batch_runs=True, **kwargs):
        """Best-of-``runs`` factorization; sets ``tl_object``, ``norm_tl_object``, ``factors``, ``rank``,
        ``explained_variance_`` and ``explained_variance_ratio_`` (tensor.py:216-362).

        The runs are independent: with ``batch_runs`` as many as fit 28 factor columns are factorized side by side in one
        device run (one read of the tensor per pass serves all of them; MU, unmasked, host-seeded random init -- what
        ``run_tensor_cell2cell_pipeline(tf_optimization='robust')`` asks for with its 100 runs), and the units are spread
        over the GPUs of the job when ``torch.distributed`` is initialised.  Same result as run by run."""
        tensor_dim = len(self.tensor.shape)
        kwargs = dict(kwargs or {})
        kwargs["return_errors"] = True
        from ..parallel import run_units
        from .factorization import _factorize_batch, _factorize_multi, _multi_eligible, _pack_runs
        from .._lib import C2CError

        def one(run):
            seed = None if random_state is None else random_state + run
            local_tf, errors = _compute_tensor_factorization(tensor=self.tensor, rank=rank, tf_type=tf_type, init=init,
                                                             svd=svd, random_state=seed, mask=self.mask,
                                                             n_iter_max=n_iter_max, tol=tol, verbose=verbose, **kwargs)
            if self.mask is None:
                err = float(errors[-1])
            else:
                err = float(_compute_norm_error(self.tensor, local_tf, self.mask))
            return (err, local_tf)

        batchable = (batch_runs and runs > 1 and tf_type == "non_negative_cp" and self.mask is None and init == "random" and
                     random_state is not None and set(kwargs) <= {"return_errors", "cvg_criterion"} and 2 * int(rank) <= 28 and
                     tensor_dim >= 3)

        def one_bin(unit):
            try:
                res = _factorize_batch(self.tensor, [(rank, random_state + run) for _, run in unit], n_iter_max, tol,
                                       cvg_criterion=kwargs.get("cvg_criterion", "abs_rec_error"))
                return [(run, float(errs[-1]), tf_) for (_, run), (tf_, errs) in zip(unit, res)]
            except C2CError:                      # shape outside the fused update kernels: run by run on the device instead
                return [(run,) + one(run) for _, run in unit]

        multi = batchable_multi = False
        if (batch_runs and runs > 1 and tf_type == "non_negative_cp" and self.mask is None and init == "random" and
                random_state is not None and set(kwargs) <= {"return_errors", "cvg_criterion"} and int(rank) <= 32):
            batchable_multi = _multi_eligible(self.tensor, [int(rank)] * runs, n_iter_max)

        def one_share(unit):
            try:
                res = _factorize_multi(self.tensor, [(int(rank), random_state + run) for run in unit], n_iter_max, tol,
                                       cvg_criterion=kwargs.get("cvg_criterion", "abs_rec_error"))
                return [(run, float(errs[-1]), tf_) for run, (tf_, errs) in zip(unit, res)]
            except C2CError as err:
                if getattr(err, "code", 0) != -61:
                    raise
                return [(run,) + one(run) for run in unit]

        if batchable_multi:
            # small tensor: every run is one CTA of a single launch (per GPU of the job)
            from ..parallel import world
            _, size = world()
            shares = [list(range(k, runs, size)) for k in range(size) if k < runs]
            parts = run_units(shares, one_share, gather="object")
            results = [(err, tf_) for _, err, tf_ in sorted((r for part in parts for r in part), key=lambda r: r[0])]
            multi = True
        if multi:
            pass
        elif batchable:
            parts = run_units(_pack_runs([(int(rank), run) for run in range(runs)]), one_bin, gather="object")
            results = [(err, tf_) for _, err, tf_ in sorted((r for part in parts for r in part), key=lambda r: r[0])]
        else:
            results = run_units(list(range(runs)), one, gather="object") if runs > 1 else [one(0)]
        best_err, tf = np.inf, None
        for err, local_tf in results:
            if best_err > err:
                best_err, tf = err, local_tf
        if runs > 1:
            print("Best model has a normalized error of: {0:.3f}".format(best_err))

        self.tl_object = tf
        if normalize_loadings:
            self.norm_tl_object = _cp_normalize(self.tl_object)

        factor_names = ["Factor {}".format(i) for i in range(1, rank + 1)]
        if self.order_labels is None:
            if tensor_dim == 4:
                order_labels = ["Contexts", "Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
            elif tensor_dim > 4:
                order_labels = ["Contexts-{}".format(i + 1) for i in range(tensor_dim - 3)] + \
                               ["Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
            elif tensor_dim == 3:
                order_labels = ["Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
            else:
                raise ValueError("Too few dimensions in the tensor")
        else:
            assert len(self.order_labels) == tensor_dim, \
                "The length of order_labels must match the number of orders/dimensions in the tensor"
            order_labels = self.order_labels

        if normalize_loadings:
            weights, factors = self.norm_tl_object
            weights = np.asarray(weights)
            if var_ordered_factors:
                w_order = weights.argsort()[::-1]
                factors = [np.asarray(f)[:, w_order] for f in factors]
                self.explained_variance_ratio_ = weights[w_order] / sum(weights)
            else:
                factors = [np.asarray(f) for f in factors]
                self.explained_variance_ratio_ = weights / sum(weights)
        else:
            weights, factors = self.tl_object
            self.explained_variance_ratio_ = None

        self.explained_variance_ = self.explained_variance()
        self.factors = OrderedDict(zip(order_labels,
                                       [pd.DataFrame(np.asarray(f), index=idx, columns=factor_names)
                                        for f, idx in zip(factors, self.order_names)]))
        self.rank = rank

    def elbow_rank_selection(self, upper_rank=50, runs=20, tf_type="non_negative_cp", init="random", svd="truncated_svd",
                             metric="error", random_state=None, n_iter_max=100, tol=10e-7, automatic_elbow=True,
                             manual_elbow=None, smooth=False, mask=None, ci="std", figsize=(4, 2.25), fontsize=14,
                             filename=None, output_fig=True, verbose=False, **kwargs):
        """Elbow analysis over ranks 1..``upper_rank`` (tensor.py:364-568).  Returns ``(fig, loss)``."""
        assert metric in ["similarity", "error"], "`metric` must be either 'similarity' or 'error'"
        ylabel = {"similarity": "Similarity\n(1-CorrIndex)", "error": "Normalized Error"}
        if verbose:
            print("Running Elbow Analysis")
        if mask is None and self.mask is not None:
            mask = self.mask
        if metric == "similarity":
            assert runs > 1, "`runs` must be greater than 1 when `metric` = 'similarity'"
        fig = None
        if runs == 1:
            loss = _run_elbow_analysis(tensor=self.tensor, upper_rank=upper_rank, tf_type=tf_type, init=init, svd=svd,
                                       random_state=random_state, mask=mask, n_iter_max=n_iter_max, tol=tol,
                                       verbose=verbose, **kwargs)
            loss = [(l[0], l[1].item()) for l in loss]
            all_loss = np.array([[l[1] for l in loss]])
            if automatic_elbow:
                if smooth:
                    loss = [(i + 1, l) for i, l in enumerate(smooth_curve([l[1] for l in loss]))]
                rank = int(_compute_elbow(loss))
            else:
                rank = manual_elbow
            if output_fig:
                fig = _plotting().plot_elbow(loss=loss, elbow=rank, figsize=figsize, ylabel=ylabel[metric],
                                             fontsize=fontsize, filename=filename)
        elif runs > 1:
            all_loss = _multiple_runs_elbow_analysis(tensor=self.tensor, upper_rank=upper_rank, runs=runs, tf_type=tf_type,
                                                     init=init, svd=svd, metric=metric, random_state=random_state,
                                                     mask=mask, n_iter_max=n_iter_max, tol=tol, verbose=verbose, **kwargs)
            loss = np.nanmean(all_loss, axis=0).tolist()
            if smooth:
                loss = smooth_curve(loss)
            loss = [(i + 1, l) for i, l in enumerate(loss)]
            if automatic_elbow:
                rank = int(_compute_elbow(loss))
            else:
                rank = manual_elbow
            if output_fig:
                fig = _plotting().plot_multiple_run_elbow(all_loss=all_loss, ci=ci, elbow=rank, figsize=figsize,
                                                          ylabel=ylabel[metric], smooth=smooth, fontsize=fontsize,
                                                          filename=filename)
        else:
            assert runs > 0, "Input runs must be an integer greater than 0"

        self.rank = rank
        self.elbow_metric = metric
        self.elbow_metric_mean = loss
        self.elbow_metric_raw = all_loss
        if self.rank is not None:
            assert isinstance(rank, int), "rank must be an integer."
            print("The rank at the elbow is: {}".format(self.rank))
        return fig, loss

    def get_top_factor_elements(self, order_name, factor_name, top_number=10):
        """Top elements of one factor of one order (tensor.py:570-592)."""
        return self.factors[order_name][factor_name].sort_values(ascending=False).head(top_number)

    def export_factor_loadings(self, filename):
        """One Excel sheet per order (tensor.py:594-606); needs an Excel writer engine (openpyxl)."""
        writer = pd.ExcelWriter(filename)
        for k, v in self.factors.items():
            v.to_excel(writer, sheet_name=k)
        writer.close()
        print("Loadings of the tensor factorization were successfully saved into {}".format(filename))

    def excluded_value_fraction(self):
        """Fraction of entries excluded by the mask (tensor.py:608-624)."""
        if self.mask is None:
            print("The interaction tensor does not have masked values")
            return 0.0
        return 1.0 - float(self.mask.sum(dtype=torch.int64).item()) / float(self.tensor.numel())

    def sparsity_fraction(self):
        """Fraction of real zeros (tensor.py:626-640)."""
        if self.tensor is None:
            print("The interaction tensor does not have zeros")
            return 0.0
        return float(self.loc_zeros.sum(dtype=torch.int64).item()) / float(self.tensor.numel())

    def missing_fraction(self):
        """Fraction of entries that were NaN (tensor.py:642-657)."""
        if self.tensor is None:
            print("The interaction tensor does not have missing values")
            return 0.0
        return float(self.loc_nans.sum(dtype=torch.int64).item()) / float(self.tensor.numel())

    def explained_variance(self):
        """sklearn-style explained variance of the factorization (tensor.py:659-688), from one fused pass of sums.

        With a mask the reference sets ``rec_tensor = tensor * mask`` (tensor.py:675), so its numerator is exactly 0 and
        the score is 1.0 whenever the masked tensor is not constant; that behaviour is reproduced, not corrected."""
        assert self.tl_object is not None, "Must run compute_tensor_factorization before using this method."
        weights, factors = self.tl_object
        rank = int(np.asarray(factors[0]).shape[1])
        s = getattr(self.tl_object, "recon_sums_", None)        # left by the factorization's own final pass over this tensor
        if s is None or self.mask is not None:
            fs = []
            for i, f in enumerate(factors):
                f = np.asarray(f, dtype=np.float64)
                if i == 0:
                    f = f * np.asarray(weights, dtype=np.float64)[None, :]
                fs.append(engine.pack_factor(f.astype(np.float32), rank, self.tensor.device))
            s = engine.recon_sums(self.tensor, self.mask, fs, rank)
        n = float(self.tensor.numel())
        sum_d, sum_dd, sum_x, sum_xx = s[0], s[1], s[2], s[3]
        if self.mask is not None:
            sum_d, sum_dd = 0.0, 0.0                      # tensor*mask - tensor*mask
        num = np.sqrt(max(sum_dd - sum_d * sum_d / n, 0.0))
        den = np.sqrt(max(sum_xx - sum_x * sum_x / n, 0.0))
        if den == 0.0:
            return 0.0
        return float(1.0 - num / den)


def _cp_normalize(cp):
    """tensorly ``cp_normalize`` (tensor.py:327) through ``c2c_cp_normalize``."""
    weights, factors = cp
    rank = int(np.asarray(factors[0]).shape[1])
    dev = engine.require_cuda(getattr(cp, "device", None))
    fs = [engine.pack_factor(np.asarray(f, dtype=np.float32), rank, dev) for f in factors]
    w = engine.cp_normalize(fs, rank, torch.as_tensor(np.asarray(weights, dtype=np.float32)))
    return CPTensor(w.cpu().numpy(), [engine.unpack_factor(f, rank).cpu().numpy() for f in fs], device=dev)


def _plotting():
    try:
        from . import plotting
        plotting.require_matplotlib()
        return plotting
    except ImportError as e:
        raise ImportError("output_fig=True needs matplotlib (not installed): pass output_fig=False. (%s)" % e)


def _to_f32(df):
    return np.ascontiguousarray(df.to_numpy(dtype=np.float32, na_value=np.nan) if hasattr(df, "to_numpy") else df,
                                dtype=np.float32)


def _build(tables, rnaseq_matrices, ppi_data, how, outer_fraction, communication_score, complex_sep, complex_agg_method,
           upper_letter_comparison, interaction_columns, group_ppi_by, group_ppi_method, device, context_slab=None):
    """``context_slab=(lo, hi)``: the gene / cell / LR-pair selection is made over ALL contexts (names only, host side), but
    only the matrices of contexts ``lo..hi-1`` are uploaded and only that slab ``[hi - lo, L, K, K]`` of the tensor is built
    -- the contexts are independent (the reference loops over them, tensor.py:1126-1129), so a rank of a context-sharded
    factorization builds exactly its own shard."""
    if communication_score not in _SCORES:
        raise ValueError("Not a valid communication_score")
    if how not in ("inner", "outer", "outer_genes", "outer_cells"):
        raise ValueError('Provide a valid way to build the tensor; "how" must be "inner", "outer", "outer_genes" or "outer_cells"')
    if group_ppi_by is not None and group_ppi_method not in ("gmean", "sum", "mean"):
        raise ValueError("Not a valid method")
    dev = engine.require_cuda(device)
    names = [t.names for t in tables]
    columns = [list(df.columns) for df in rnaseq_matrices]
    genes = bi.select_elements(names, how in ("outer", "outer_genes"), outer_fraction)
    cells = bi.select_elements(columns, how in ("outer", "outer_cells"), outer_fraction)
    genes = bi.ordered(names[0], genes)
    cells = bi.ordered(columns[0], cells)
    ppi = bi.filter_lr_pairs(ppi_data, genes, complex_sep, upper_letter_comparison, interaction_columns)
    col_a, col_b = interaction_columns
    lig, rec = ppi[col_a].tolist(), ppi[col_b].tolist()
    if context_slab is not None:
        lo, hi = (int(v) for v in context_slab)
        if not 0 <= lo < hi <= len(tables):
            raise ValueError("context_slab %r is not a non-empty range of the %d contexts" % (context_slab, len(tables)))
        tables, columns, rnaseq_matrices = tables[lo:hi], columns[lo:hi], rnaseq_matrices[lo:hi]
    C, L, K = len(tables), len(lig), len(cells)
    idx = bi.resolve(tables, columns, genes, cells, lig, rec)
    mats = [_to_f32(df) for df in rnaseq_matrices]
    off = np.zeros(C, dtype=np.int64)
    ld = np.zeros(C, dtype=np.int32)
    pos = 0
    for c, m in enumerate(mats):
        off[c], ld[c] = pos, m.shape[1]
        pos += m.size
    expr = np.concatenate([m.reshape(-1) for m in mats]) if pos else np.zeros(1, np.float32)
    grouped = group_ppi_by is not None
    with_mask = how != "inner" or grouped
    # no multi-subunit partner in this build -> the aggregation method is irrelevant: take the exact-fp32 ('min') path
    multi = (idx["a_count"].max(initial=0) > 1) or (idx["b_count"].max(initial=0) > 1)
    X, mask = engine.build_ccc(expr, off, ld, idx["cell_col"], idx["a_first"], idx["a_count"], idx["b_first"],
                               idx["b_count"], idx["sub_rows"], C, L, K, communication_score,
                               complex_agg_method if multi else "min", with_mask, dev)
    if grouped:
        keys, g_off, g_rows = bi.group_rows(ppi, group_ppi_by)
        X, mask = engine.group_aggregate(X, mask, g_off, g_rows, group_ppi_method)
        ppi_names = list(keys)
    else:
        ppi_names = [a + "^" + b for a, b in zip(lig, rec)]
    if how == "inner":
        mask = None
    return X, genes, cells, ppi_names, mask


def build_context_ccc_tensor(rnaseq_matrices, ppi_data, how="inner", outer_fraction=0.0,
                             communication_score="expression_product", complex_sep=None, upper_letter_comparison=True,
                             interaction_columns=("A", "B"), group_ppi_by=None, group_ppi_method="gmean", verbose=True,
                             device=None, context_slab=None):
    """Same contract as the reference's ``build_context_ccc_tensor`` (tensor.py:973-1146): returns
    ``(tensor, genes, cells, ppi_names, mask)`` -- tensor fp32 ``[C, L, K, K]`` and mask uint8 (None for how='inner') as
    CUDA tensors.  The matrices are used as given: protein complexes must already be rows (``InteractionTensor`` adds
    them), and -- as in the reference -- only the PPI names are upper-cased by ``upper_letter_comparison``.
    ``context_slab=(lo, hi)`` (not in the reference) builds only those contexts' slab of the same tensor."""
    tables = [bi.plain_rows(list(df.index)) for df in rnaseq_matrices]
    return _build(tables, rnaseq_matrices, ppi_data, how, outer_fraction, communication_score, complex_sep, "min",
                  upper_letter_comparison, interaction_columns, group_ppi_by, group_ppi_method, device, context_slab)


class InteractionTensor(BaseTensor):
    """4D-communication tensor built from per-context expression matrices and a ligand-receptor list
    (tensor.py:691-883)."""

    def __init__(self, rnaseq_matrices, ppi_data, order_labels=None, context_names=None, how="inner", outer_fraction=0.0,
                 communication_score="expression_mean", complex_sep=None, complex_agg_method="min",
                 upper_letter_comparison=True, interaction_columns=("A", "B"), group_ppi_by=None,
                 group_ppi_method="gmean", device=None, verbose=True, context_slab=None):
        """``context_slab=(lo, hi)`` (not in the reference): this object holds only the slab of contexts ``lo..hi-1`` of the
        tensor the full list would give -- same genes, cells, LR pairs and values -- e.g. one rank's shard of a
        context-sharded factorization (``cell2cell_b200.sharded``); ``full_contexts`` keeps the full context count."""
        if group_ppi_by is not None:
            assert group_ppi_by in ppi_data.columns, \
                "Using {} for grouping PPIs is not possible. Not present among columns in ppi_data".format(group_ppi_by)
        BaseTensor.__init__(self)
        if complex_sep is not None:
            if verbose:
                print("Getting expression values for protein complexes")
            if complex_agg_method not in ("min", "mean", "gmean"):
                raise ValueError("{} is not a valid agg_method".format(complex_agg_method))
            complexes = bi.complex_table(ppi_data, complex_sep, interaction_columns)[4]
        else:
            complexes = None
            complex_agg_method = "min"
        tables = bi.context_tables([df.index for df in rnaseq_matrices], complexes, upper_letter_comparison)
        if verbose:
            print("Building tensor for the provided context")
        tensor, genes, cells, ppi_names, mask = _build(tables, rnaseq_matrices, ppi_data, how, outer_fraction,
                                                       communication_score, complex_sep, complex_agg_method,
                                                       upper_letter_comparison, interaction_columns, group_ppi_by,
                                                       group_ppi_method, device, context_slab)
        if context_names is None:
            context_names = ["C-" + str(i) for i in range(1, len(rnaseq_matrices) + 1)]
        self.full_contexts = len(rnaseq_matrices)
        self.full_context_names = list(context_names)
        self.context_slab = None if context_slab is None else (int(context_slab[0]), int(context_slab[1]))
        if context_slab is not None:
            context_names = list(context_names)[self.context_slab[0]:self.context_slab[1]]
        self.communication_score = communication_score
        self.how = how
        self.outer_fraction = outer_fraction
        self.tensor = tensor
        self.mask = mask
        self.genes = genes
        self.cells = cells
        self.order_labels = order_labels
        self.order_names = [context_names, ppi_names, self.cells, self.cells]


class PreBuiltTensor(BaseTensor):
    """Wraps an existing N-way array (N >= 3) (tensor.py:886-970)."""

    def __init__(self, tensor, order_names, order_labels=None, mask=None, loc_nans=None, device=None, shared_upload=False):
        """``shared_upload=True`` (not in the reference; needs ``torch.distributed`` on NCCL): every rank of the job passes the
        same host fp32 tensor, so each uploads 1 / world of it and an all-gather over NVLink builds the replicas."""
        BaseTensor.__init__(self)
        t = tensor.detach() if isinstance(tensor, torch.Tensor) else np.asarray(tensor)
        if device is None and isinstance(t, torch.Tensor) and t.is_cuda:
            device = t.device
        X = as_device_tensor(t, device, shared_upload=shared_upload)   # NaN / inf survive the cast to fp32; scanned on the device
        # one reduction pass decides whether the clean-up passes (loc_nans, nan_to_num: tensor.py:933-946) are needed at all:
        # sum(x^2) is finite only if every entry is (it is also the ||X||^2 the factorization needs, so it is kept)
        nx2 = engine.sumsq(X)
        finite = bool(np.isfinite(nx2))
        if finite:
            X._c2c_norm_x2 = nx2
        if finite and loc_nans is None:
            self._loc_nans = None                         # derived on first use: all zeros
            self._loc_zeros = None                        # derived on first use: (tensor == 0)
            self._no_nans = True
            self.tensor = X
        else:
            nan = torch.isnan(X)
            if loc_nans is not None:
                assert tuple(loc_nans.shape) == tuple(X.shape), "`loc_nans` and `tensor` must be of the same shape"
                nan = nan | (torch.as_tensor(np.asarray(loc_nans)).to(X.device) > 0)
            self._loc_nans = nan.to(torch.uint8)
            self._loc_zeros = ((X == 0) & ~nan).to(torch.uint8)
            self.tensor = torch.nan_to_num(X)
        self.mask = as_device_mask(mask, self.tensor)
        self.order_names = order_names
        if order_labels is None:
            self.order_labels = ["Dimension-{}".format(i + 1) for i in range(len(self.tensor.shape))]
        else:
            self.order_labels = order_labels
        assert len(self.tensor.shape) == len(self.order_labels), \
            "The length of order_labels must match the number of orders/dimensions in the tensor"

    @property
    def loc_nans(self):
        if self._loc_nans is None and getattr(self, "_no_nans", False):
            self._loc_nans = torch.zeros(self.tensor.shape, dtype=torch.uint8, device=self.tensor.device)
        return self._loc_nans

    @loc_nans.setter
    def loc_nans(self, value):
        self._loc_nans = value


_DEFAULT_TAIL = ["Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]


def _default_categories(n_orders):
    """Category used for an order whose elements get no metadata of their own (tensor.py:1283-1292)."""
    if n_orders < 3:
        raise ValueError("Too few dimensions in the tensor")
    if n_orders == 3:
        return list(_DEFAULT_TAIL)
    if n_orders == 4:
        return ["Contexts"] + _DEFAULT_TAIL
    return ["Contexts-{}".format(i + 1) for i in range(n_orders - 3)] + _DEFAULT_TAIL


def generate_tensor_metadata(interaction_tensor, metadata_dicts, fill_with_order_elements=True):
    """One ``(Element, Category)`` DataFrame per order of the tensor (tensor.py:1255-1315).

    ``metadata_dicts[i]`` maps the elements of order ``i`` to their category, or is None: the order then gets no table
    (``fill_with_order_elements=False``), or one whose category is the element itself (fewer than 75 elements) or the order's
    label (the tensor's ``order_labels``, else the default names)."""
    names_per_order = interaction_tensor.order_names
    n_orders = len(interaction_tensor.tensor.shape)
    assert n_orders == len(metadata_dicts), \
        "metadata_dicts should be of the same size as the number of orders/dimensions in the tensor"
    labels = interaction_tensor.order_labels
    if labels is None:
        labels = _default_categories(n_orders)
    else:
        assert len(labels) == n_orders, "The length of order_labels must match the number of orders/dimensions in the tensor"
    tables = []
    for label, names, mapping in zip(labels, names_per_order, metadata_dicts):
        if mapping is None and not fill_with_order_elements:
            tables.append(None)
            continue
        elements = pd.Index(names)
        if mapping is not None:
            category = [mapping[e] for e in elements]
        elif len(elements) < 75:
            category = elements
        else:
            category = [label] * len(elements)
        tables.append(pd.DataFrame({"Element": elements, "Category": category}))
    return tables


def interactions_to_tensor(interactions, experiment="single_cell", context_names=None, how="inner", outer_fraction=0.0,
                           communication_score="expression_product", upper_letter_comparison=True, verbose=True):
    """4D tensor from a list of interaction pipelines, one per context (tensor.py:1318-1419).

    The pipelines (``BulkInteractions`` / ``SingleCellInteractions``, cell2cell/analysis/cell2cell_pipelines.py) are outside
    this package; any object with their attributes works: ``complex_sep``, ``complex_agg_method``, ``interaction_columns``,
    ``analysis_setup['cci_type']``, ``ppi_data`` / ``ref_ppi`` and ``aggregated_expression`` (single cell) or
    ``rnaseq_data`` (bulk).
    """
    ppis = []
    rnaseq_matrices = []
    first = interactions[0]
    for inter in interactions:
        ppis.append(inter.ref_ppi if inter.analysis_setup["cci_type"] == "undirected" else inter.ppi_data)
        if experiment == "single_cell":
            rnaseq_matrices.append(inter.aggregated_expression)
        elif experiment == "bulk":
            rnaseq_matrices.append(inter.rnaseq_data)
        else:
            raise ValueError("experiment must be 'single_cell' or 'bulk'")
    ppi_data = pd.concat(ppis).drop_duplicates().reset_index(drop=True)
    return InteractionTensor(rnaseq_matrices=rnaseq_matrices, ppi_data=ppi_data, context_names=context_names, how=how,
                             outer_fraction=outer_fraction, complex_sep=first.complex_sep,
                             complex_agg_method=first.complex_agg_method, interaction_columns=first.interaction_columns,
                             communication_score=communication_score, upper_letter_comparison=upper_letter_comparison,
                             verbose=verbose)


_ = tqdm
