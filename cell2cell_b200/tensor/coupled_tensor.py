"""``CoupledInteractionTensor``: same class, methods and attributes as ``cell2cell/tensor/coupled_tensor.py:22-720``, with the
tensors resident on the GPU (fp32, masks uint8) and the factorization done by ``coupled_factorization`` on the device.
There is no CPU fallback (contrast ``to_device`` at coupled_tensor.py:546-562, which prints and stays on the CPU)."""
import copy
import pickle
from collections import OrderedDict

import numpy as np
import pandas as pd
import torch
from tqdm.auto import tqdm

from .. import engine
from .coupled_factorization import (_combined_error, _compute_coupled_tensor_factorization,
                                    _multiple_runs_coupled_elbow_analysis, _process_mode_mapping, _run_coupled_elbow_analysis,
                                    _split_modes, _validate_tensors)
from .elbow import smooth_curve
from .factorization import _compute_elbow, _compute_norm_error, as_device_mask, as_device_tensor
from .tensor import _cp_normalize, _plotting

__all__ = ["CoupledInteractionTensor"]

_DEVICE_ATTRS = ("tensor1", "tensor2", "mask1", "mask2", "_loc_nans1", "_loc_nans2", "_loc_zeros1", "_loc_zeros2")


def _default_labels(ndim):
    if ndim == 4:
        return ["Contexts", "Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
    if ndim > 4:
        return ["Contexts-{}".format(i + 1) for i in range(ndim - 3)] + ["Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
    if ndim == 3:
        return ["Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
    raise ValueError("Too few dimensions in the tensor")


class CoupledInteractionTensor:
    """Simultaneous non-negative CP decomposition of two interaction tensors that share some of their modes
    (``mode_mapping = {'shared': [(t1_mode, t2_mode), ...]}``, or the non-shared mode(s) of two tensors of the same order)."""

    def __init__(self, tensor1, tensor2, mode_mapping, tensor1_name="Tensor1", tensor2_name="Tensor2", balance_errors=True,
                 device=None):
        self.mode_mapping = _process_mode_mapping(tensor1.tensor, tensor2.tensor, mode_mapping)
        _validate_tensors(tensor1, tensor2, self.mode_mapping)
        dev = engine.require_cuda(device) if device is not None else None
        self.tensor1 = as_device_tensor(tensor1.tensor, dev)
        self.tensor2 = as_device_tensor(tensor2.tensor, self.tensor1.device)
        self.tensor1_name = tensor1_name
        self.tensor2_name = tensor2_name
        self.balance_errors = balance_errors
        self.order_names1 = [list(n) for n in tensor1.order_names]
        self.order_names2 = [list(n) for n in tensor2.order_names]
        self.order_labels1 = list(tensor1.order_labels) if tensor1.order_labels else None
        self.order_labels2 = list(tensor2.order_labels) if tensor2.order_labels else None
        self.mask1 = as_device_mask(tensor1.mask, self.tensor1)
        self.mask2 = as_device_mask(tensor2.mask, self.tensor2)
        self.tl_object1 = None
        self.tl_object2 = None
        self.norm_tl_object1 = None
        self.norm_tl_object2 = None
        self.factors1 = None
        self.factors2 = None
        self.factors = None
        self.rank = None
        self.explained_variance_ = None
        self.explained_variance_ratio_ = None
        self.elbow_metric = None
        self.elbow_metric_mean = None
        self.elbow_metric_raw = None
        # loc_nans / loc_zeros of the two tensors: taken over when the source objects hold them, else derived on first use
        self._loc_nans1 = self._on_device(getattr(tensor1, "_loc_nans", None), self.tensor1)
        self._loc_nans2 = self._on_device(getattr(tensor2, "_loc_nans", None), self.tensor2)
        self._loc_zeros1 = self._on_device(getattr(tensor1, "_loc_zeros", None), self.tensor1)
        self._loc_zeros2 = self._on_device(getattr(tensor2, "_loc_zeros", None), self.tensor2)
        self._no_nans1 = bool(getattr(tensor1, "_no_nans", False))      # PreBuiltTensor: the array held no NaN (tensor.py:933-946)
        self._no_nans2 = bool(getattr(tensor2, "_no_nans", False))

    @staticmethod
    def _on_device(v, like):
        if v is None:
            return None
        return torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).to(like.device)

    # ---- derived full-size arrays (coupled_tensor.py:118-137) ---------------------------------------------------------
    def _loc_nans(self, k):
        cur = getattr(self, "_loc_nans%d" % k)
        if cur is None:
            t, m = getattr(self, "tensor%d" % k), getattr(self, "mask%d" % k)
            if m is None or getattr(self, "_no_nans%d" % k, False):
                cur = torch.zeros(t.shape, dtype=torch.uint8, device=t.device)
            else:
                cur = (1 - m).to(torch.uint8)
            setattr(self, "_loc_nans%d" % k, cur)
        return cur

    def _loc_zeros(self, k):
        cur = getattr(self, "_loc_zeros%d" % k)
        if cur is None:
            t = getattr(self, "tensor%d" % k)
            cur = ((t == 0) & (self._loc_nans(k) == 0)).to(torch.uint8)
            setattr(self, "_loc_zeros%d" % k, cur)
        return cur

    loc_nans1 = property(lambda self: self._loc_nans(1))
    loc_nans2 = property(lambda self: self._loc_nans(2))
    loc_zeros1 = property(lambda self: self._loc_zeros(1))
    loc_zeros2 = property(lambda self: self._loc_zeros(2))

    # ---- factorization ----------------------------------------------------------------------------------------------
    def _loss(self, cp1, cp2, errors1, errors2):
        error1 = float(errors1[-1]) if self.mask1 is None else float(_compute_norm_error(self.tensor1, cp1, self.mask1))
        error2 = float(errors2[-1]) if self.mask2 is None else float(_compute_norm_error(self.tensor2, cp2, self.mask2))
        return _combined_error(error1, error2, self.tensor1.shape, self.tensor2.shape, self.mode_mapping, self.balance_errors)

    def compute_tensor_factorization(self, rank, tf_type="coupled_non_negative_cp", init="svd", svd="truncated_svd",
                                     random_state=None, runs=1, normalize_loadings=True, var_ordered_factors=True,
                                     n_iter_max=100, tol=10e-7, verbose=False, **kwargs):
        """Best-of-``runs`` coupled factorization (coupled_tensor.py:140-265); sets ``tl_object1/2``, ``norm_tl_object1/2``,
        ``factors1``, ``factors2``, the unified ``factors``, ``rank``, ``explained_variance_(ratio_)``."""
        kwargs = dict(kwargs or {})
        kwargs["return_errors"] = True
        from ..parallel import run_units

        def one(run):
            seed = None if random_state is None else random_state + run
            cp1, cp2, (errors1, errors2) = _compute_coupled_tensor_factorization(
                tensor1=self.tensor1, tensor2=self.tensor2, rank=rank, mode_mapping=self.mode_mapping, mask1=self.mask1,
                mask2=self.mask2, tf_type=tf_type, init=init, svd=svd, random_state=seed, n_iter_max=n_iter_max, tol=tol,
                verbose=verbose, balance_errors=self.balance_errors, **kwargs)
            return (self._loss(cp1, cp2, errors1, errors2), cp1, cp2)

        results = run_units(list(range(runs)), one, gather="object") if runs > 1 else [one(0)]
        best_error, best_cp1, best_cp2 = np.inf, None, None
        for err, cp1, cp2 in results:
            if err < best_error:
                best_error, best_cp1, best_cp2 = err, cp1, cp2
        if runs > 1:
            print("Best coupled model has a combined normalized error of: {:.3f}".format(best_error))
        self.tl_object1 = best_cp1
        self.tl_object2 = best_cp2
        self.rank = rank
        self._create_factor_dataframes(normalize_loadings, var_ordered_factors)
        self.explained_variance_ = self.explained_variance()

    def _create_factor_dataframes(self, normalize_loadings, var_ordered_factors):
        """Per-tensor and unified factor DataFrames (coupled_tensor.py:267-361)."""
        factor_names = ["Factor {}".format(i) for i in range(1, self.rank + 1)]
        if self.order_labels1 is None:
            self.order_labels1 = _default_labels(len(self.tensor1.shape))
        if self.order_labels2 is None:
            self.order_labels2 = _default_labels(len(self.tensor2.shape))
        if normalize_loadings:
            self.norm_tl_object1 = _cp_normalize(self.tl_object1)
            self.norm_tl_object2 = _cp_normalize(self.tl_object2)
            weights1, factors1 = self.norm_tl_object1
            weights2, factors2 = self.norm_tl_object2
            weights1, weights2 = np.asarray(weights1), np.asarray(weights2)
            avg_weights = (weights1 + weights2) / 2
            if var_ordered_factors:
                w_order = avg_weights.argsort()[::-1]
                factors1 = [np.asarray(f)[:, w_order] for f in factors1]
                factors2 = [np.asarray(f)[:, w_order] for f in factors2]
                self.explained_variance_ratio_ = avg_weights[w_order] / sum(avg_weights)
            else:
                factors1 = [np.asarray(f) for f in factors1]
                factors2 = [np.asarray(f) for f in factors2]
                self.explained_variance_ratio_ = avg_weights / sum(avg_weights)
        else:
            factors1 = [np.asarray(f) for f in self.tl_object1[1]]
            factors2 = [np.asarray(f) for f in self.tl_object2[1]]
            self.explained_variance_ratio_ = None
        self.factors1 = OrderedDict(zip(self.order_labels1, [pd.DataFrame(f, index=idx, columns=factor_names)
                                                             for f, idx in zip(factors1, self.order_names1)]))
        self.factors2 = OrderedDict(zip(self.order_labels2, [pd.DataFrame(f, index=idx, columns=factor_names)
                                                             for f, idx in zip(factors2, self.order_names2)]))
        shared_pairs = self.mode_mapping.get("shared", [])
        only1, only2 = _split_modes(len(self.tensor1.shape), len(self.tensor2.shape), shared_pairs)
        self.factors = OrderedDict()
        for a, _ in shared_pairs:                               # shared factors: tensor 1's version
            self.factors[self.order_labels1[a]] = self.factors1[self.order_labels1[a]]
        for mode in only1:
            self.factors[self.order_labels1[mode]] = self.factors1[self.order_labels1[mode]]
        for mode in only2:
            label = self.order_labels2[mode]
            unified = "{}_{}".format(label, self.tensor2_name) if label in self.factors.keys() else label
            self.factors[unified] = self.factors2[label]

    def elbow_rank_selection(self, upper_rank=50, runs=20, tf_type="coupled_non_negative_cp", init="random",
                             svd="truncated_svd", metric="error", random_state=None, n_iter_max=100, tol=10e-7,
                             automatic_elbow=True, manual_elbow=None, smooth=False, mask1=None, mask2=None, ci="std",
                             figsize=(4, 2.25), fontsize=14, filename=None, output_fig=True, verbose=False, **kwargs):
        """Elbow analysis on the combined error of the coupled factorization (coupled_tensor.py:363-459).
        Returns ``(fig, loss)``."""
        assert metric in ["error"], "`metric` must be 'error' for coupled factorization"
        ylabel = "Normalized Error"
        if verbose:
            print("Running Coupled Elbow Analysis")
        mask1 = self.mask1 if mask1 is None else mask1
        mask2 = self.mask2 if mask2 is None else mask2
        common = dict(tensor1=self.tensor1, tensor2=self.tensor2, mode_mapping=self.mode_mapping, upper_rank=upper_rank,
                      tf_type=tf_type, init=init, svd=svd, random_state=random_state, mask1=mask1, mask2=mask2,
                      n_iter_max=n_iter_max, tol=tol, verbose=verbose, balance_errors=self.balance_errors)
        if runs == 1:
            loss = _run_coupled_elbow_analysis(**common, **kwargs)
            loss = [(l[0], l[1].item() if hasattr(l[1], "item") else l[1]) for l in loss]
            all_loss = np.array([[l[1] for l in loss]])
        else:
            all_loss = _multiple_runs_coupled_elbow_analysis(runs=runs, metric=metric, **common, **kwargs)
            loss = np.nanmean(all_loss, axis=0).tolist()
            loss = [(i + 1, l) for i, l in enumerate(loss)]
        if smooth:
            loss = [(i + 1, l) for i, l in enumerate(smooth_curve([l[1] for l in loss]))]
        rank = int(_compute_elbow(loss)) if automatic_elbow else manual_elbow
        fig = None
        if output_fig:
            if runs == 1:
                fig = _plotting().plot_elbow(loss=loss, elbow=rank, figsize=figsize, ylabel=ylabel, fontsize=fontsize,
                                             filename=filename)
            else:
                fig = _plotting().plot_multiple_run_elbow(all_loss=all_loss, ci=ci, elbow=rank, figsize=figsize, ylabel=ylabel,
                                                          smooth=smooth, fontsize=fontsize, filename=filename)
        self.rank = rank
        self.elbow_metric = metric
        self.elbow_metric_mean = loss
        self.elbow_metric_raw = all_loss
        if self.rank is not None:
            print("The rank at the elbow is: {}".format(self.rank))
        return fig, loss

    def explained_variance(self):
        """Size-weighted mean of the two tensors' explained variances (coupled_tensor.py:461-503), each from one fused pass
        of fp64 sums (``c2c_recon_sums``) instead of a materialised reconstruction."""
        if self.tl_object1 is None or self.tl_object2 is None:
            raise ValueError("Must run compute_tensor_factorization first")

        def one(tensor, mask, cp):
            weights, factors = cp
            rank = int(np.asarray(factors[0]).shape[1])
            s = getattr(cp, "recon_sums_", None) if mask is None else None
            if s is None:
                fs = []
                for i, f in enumerate(factors):
                    f = np.asarray(f, dtype=np.float64)
                    if i == 0:
                        f = f * np.asarray(weights, dtype=np.float64)[None, :]
                    fs.append(engine.pack_factor(f.astype(np.float32), rank, tensor.device))
                s = engine.recon_sums(tensor, mask, fs, rank)
            n = float(tensor.numel())                        # means run over every entry, masked ones count as zeros
            num = np.sqrt(max(s[1] - s[0] * s[0] / n, 0.0))
            den = np.sqrt(max(s[3] - s[2] * s[2] / n, 0.0))
            return 0.0 if den == 0.0 else float(1.0 - num / den)

        ev1 = one(self.tensor1, self.mask1, self.tl_object1)
        ev2 = one(self.tensor2, self.mask2, self.tl_object2)
        n1, n2 = float(self.tensor1.numel()), float(self.tensor2.numel())
        return (ev1 * n1 + ev2 * n2) / (n1 + n2)

    def get_top_factor_elements(self, order_name, factor_name, top_number=10, tensor="unified"):
        """Top elements of one factor of one order (coupled_tensor.py:505-520)."""
        if tensor == "unified":
            factors = self.factors
        elif tensor == "tensor1":
            factors = self.factors1
        elif tensor == "tensor2":
            factors = self.factors2
        else:
            raise ValueError("tensor must be 'unified', 'tensor1', or 'tensor2'")
        if order_name not in factors:
            raise ValueError("Order '{}' not found in {} factors".format(order_name, tensor))
        return factors[order_name][factor_name].sort_values(ascending=False).head(top_number)

    def export_factor_loadings(self, filename, save_separate=False):
        """One Excel sheet per order (coupled_tensor.py:522-539); needs an Excel writer engine (openpyxl)."""
        writer = pd.ExcelWriter(filename)
        if save_separate:
            for k, v in self.factors1.items():
                v.to_excel(writer, sheet_name="T1_{}".format(k))
            for k, v in self.factors2.items():
                v.to_excel(writer, sheet_name="T2_{}".format(k))
        else:
            for k, v in self.factors.items():
                v.to_excel(writer, sheet_name="{}".format(k))
        writer.close()
        print("Coupled tensor factor loadings saved to {}".format(filename))

    @property
    def shape(self):
        return (tuple(self.tensor1.shape), tuple(self.tensor2.shape))

    def to_device(self, device):
        """Moves both tensors and their masks to another CUDA device; raises for anything else."""
        dev = engine.require_cuda(device)
        for k in _DEVICE_ATTRS:
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.to(dev))

    def copy(self):
        return copy.deepcopy(self)

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in _DEVICE_ATTRS:
            v = state.get(k)
            if isinstance(v, torch.Tensor):
                state[k] = v.detach().cpu().numpy()
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if torch.cuda.is_available():
            for k in _DEVICE_ATTRS:
                v = getattr(self, k, None)
                if isinstance(v, np.ndarray):
                    setattr(self, k, torch.as_tensor(v).to(engine.require_cuda()))

    def write_file(self, filename):
        """Exports this object into a pickle file (coupled_tensor.py:569-572)."""
        with open(filename, "wb") as fh:
            pickle.dump(self, fh, protocol=pickle.HIGHEST_PROTOCOL)
        print(filename, " was correctly saved.")

    # ---- fractions (coupled_tensor.py:574-671) ----------------------------------------------------------------------
    def _fraction(self, which, tensor):
        if tensor in ("tensor1", "tensor2"):
            return which(1 if tensor == "tensor1" else 2)
        if tensor == "both":
            return {"tensor1": which(1), "tensor2": which(2)}
        if tensor == "combined":
            n1, n2 = float(self.tensor1.numel()), float(self.tensor2.numel())
            return (which(1) * n1 + which(2) * n2) / (n1 + n2)
        raise ValueError("tensor must be 'tensor1', 'tensor2', 'both', or 'combined'")

    def excluded_value_fraction(self, tensor="both"):
        def which(k):
            m, t = getattr(self, "mask%d" % k), getattr(self, "tensor%d" % k)
            return 0.0 if m is None else 1.0 - float(m.sum(dtype=torch.int64).item()) / float(t.numel())
        return self._fraction(which, tensor)

    def sparsity_fraction(self, tensor="both"):
        def which(k):
            return float(self._loc_zeros(k).sum(dtype=torch.int64).item()) / float(getattr(self, "tensor%d" % k).numel())
        return self._fraction(which, tensor)

    def missing_fraction(self, tensor="both"):
        def which(k):
            return float(self._loc_nans(k).sum(dtype=torch.int64).item()) / float(getattr(self, "tensor%d" % k).numel())
        return self._fraction(which, tensor)

    def reorder_metadata(self, metadata1, metadata2):
        """Metadata in the order of the unified ``factors``: shared modes, tensor1-specific, tensor2-specific
        (coupled_tensor.py:673-720)."""
        if self.factors is None:
            raise ValueError("Must run compute_tensor_factorization first to determine factor ordering")
        shared_pairs = self.mode_mapping.get("shared", [])
        only1, only2 = _split_modes(len(self.tensor1.shape), len(self.tensor2.shape), shared_pairs)
        if len(metadata1) != len(self.tensor1.shape):
            raise ValueError("metadata1 must have {} elements, got {}".format(len(self.tensor1.shape), len(metadata1)))
        if len(metadata2) != len(self.tensor2.shape):
            raise ValueError("metadata2 must have {} elements, got {}".format(len(self.tensor2.shape), len(metadata2)))
        return [metadata1[a] for a, _ in shared_pairs] + [metadata1[m] for m in only1] + [metadata2[m] for m in only2]


_ = tqdm
