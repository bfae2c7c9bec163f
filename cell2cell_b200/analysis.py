"""``run_tensor_cell2cell_pipeline``: the caller of the hot path, same signature and behaviour as
``cell2cell/analysis/tensor_pipelines.py:10-215`` (SURVEY.md sections 3.4 and 8f row 4): elbow analysis when no rank is given,
then the (best-of-runs) factorization, then the optional exports.

* ``tf_optimization='regular'``: 10 elbow runs, 1 factorization run, 100 iterations, tol 1e-7;
  ``'robust'``: 20 elbow runs, 100 factorization runs, 500 iterations, tol 1e-8 -- the independent runs are packed side by
  side on the device and spread over the GPUs of the job (``BaseTensor.compute_tensor_factorization`` /
  ``_multiple_runs_elbow_analysis``).
* ``backend`` is accepted for compatibility; there is one backend here (the CUDA library), so anything but None / 'pytorch'
  raises.  ``device`` must be a CUDA device.
* figures need matplotlib (absent from the build image): with ``output_fig=True`` and no matplotlib an ImportError says so.
"""
from .tensor.plotting import tensor_factors_plot

__all__ = ["run_tensor_cell2cell_pipeline"]


def run_tensor_cell2cell_pipeline(interaction_tensor, tensor_metadata, copy_tensor=False, rank=None,
                                  tf_optimization="regular", random_state=None, backend=None, device=None,
                                  elbow_metric="error", smooth_elbow=False, upper_rank=25, tf_init="random",
                                  tf_svd="numpy_svd", cmaps=None, sample_col="Element", group_col="Category",
                                  fig_fontsize=14, output_folder=None, output_fig=True, fig_format="pdf", **kwargs):
    """Runs the basic Tensor-cell2cell pipeline; returns the tensor object (or its copy) holding the results."""
    if copy_tensor:
        interaction_tensor = interaction_tensor.copy()
    dim = len(interaction_tensor.tensor.shape)

    if output_folder is None:
        elbow_filename = tf_filename = loading_filename = None
    else:
        elbow_filename = output_folder + "/Elbow.{}".format(fig_format)
        tf_filename = output_folder + "/Tensor-Factorization.{}".format(fig_format)
        loading_filename = output_folder + "/Loadings.xlsx"

    if cmaps is None:
        cmap_5d = ["tab10", "viridis", "Dark2_r", "tab20", "tab20"]
        cmap_4d = ["plasma", "Dark2_r", "tab20", "tab20"]
        if dim == 5:
            cmaps = cmap_5d
        elif dim <= 4:
            cmaps = cmap_4d[-dim:]
        else:
            raise ValueError("Tensor of dimension higher to 5 is not supported")
    assert len(cmaps) == dim, "`cmap` must be of the same len of dimensions in the tensor."

    if tf_optimization == "robust":
        elbow_runs, tf_runs, tol, n_iter_max = 20, 100, 1e-8, 500
    elif tf_optimization == "regular":
        elbow_runs, tf_runs, tol, n_iter_max = 10, 1, 1e-7, 100
    else:
        raise ValueError("`factorization_type` must be either 'robust' or 'regular'.")

    if backend not in (None, "pytorch"):
        raise ValueError("cell2cell_b200 has a single (CUDA) backend; got backend=%r" % (backend,))
    if device is not None:
        interaction_tensor.to_device(device=device)

    if rank is None:
        print("Running Elbow Analysis")
        interaction_tensor.elbow_rank_selection(upper_rank=upper_rank, runs=elbow_runs, init=tf_init, svd=tf_svd,
                                                automatic_elbow=True, metric=elbow_metric, output_fig=output_fig,
                                                smooth=smooth_elbow, random_state=random_state, fontsize=fig_fontsize,
                                                filename=elbow_filename, tol=tol, n_iter_max=n_iter_max, **kwargs)
        rank = interaction_tensor.rank

    print("Running Tensor Factorization")
    interaction_tensor.compute_tensor_factorization(rank=rank, init=tf_init, svd=tf_svd, random_state=random_state,
                                                    runs=tf_runs, normalize_loadings=True, tol=tol, n_iter_max=n_iter_max,
                                                    **kwargs)

    if output_folder is not None:
        print("Generating Outputs")
        interaction_tensor.export_factor_loadings(loading_filename)

    if output_fig:
        tensor_factors_plot(interaction_tensor=interaction_tensor, metadata=tensor_metadata, sample_col=sample_col,
                            group_col=group_col, meta_cmaps=cmaps, fontsize=fig_fontsize, filename=tf_filename)
    return interaction_tensor
