"""Elbow figures for ``elbow_rank_selection(output_fig=True)`` (cell2cell/plotting/tensor_plot.py:489-611).

matplotlib is optional (it is not installed in the build image): the import happens on first use, and the factorization
path never needs it -- pass ``output_fig=False`` on a box without it.
"""
import numpy as np

__all__ = ["require_matplotlib", "plot_elbow", "plot_multiple_run_elbow", "tensor_factors_plot"]


def require_matplotlib():
    import matplotlib  # noqa: F401
    return matplotlib


def plot_elbow(loss, elbow=None, figsize=(4, 2.25), ylabel="Normalized Error", fontsize=14, filename=None):
    from matplotlib import pyplot as plt
    fig = plt.figure(figsize=figsize)
    plt.plot(*zip(*loss))
    plt.tick_params(axis="both", labelsize=fontsize)
    plt.xlabel("Rank", fontsize=int(1.2 * fontsize))
    plt.ylabel(ylabel, fontsize=int(1.2 * fontsize))
    if elbow is not None:
        _, y = zip(*loss)
        plt.plot(*loss[elbow - 1], "ro")
    plt.tight_layout()
    if filename is not None:
        plt.savefig(filename, dpi=300, bbox_inches="tight")
    return fig


def plot_multiple_run_elbow(all_loss, elbow=None, ci="95%", figsize=(4, 2.25), ylabel="Normalized Error", fontsize=14,
                            smooth=False, filename=None):
    from matplotlib import pyplot as plt
    fig = plt.figure(figsize=figsize)
    x = np.arange(1, all_loss.shape[1] + 1)
    mean = np.nanmean(all_loss, axis=0)
    std = np.nanstd(all_loss, axis=0)
    if smooth:
        from .elbow import smooth_curve
        mean = smooth_curve(mean)
        std = smooth_curve(std)
    if ci == "95%":
        coeff = 1.96
    elif ci == "std":
        coeff = 1.0
    else:
        raise ValueError("Specify a correct ci. Either '95%' or 'std'")
    plt.plot(x, mean, "ob")
    plt.plot(x, mean, "-b")
    plt.fill_between(x, mean - coeff * std, mean + coeff * std, color="b", alpha=0.2)
    plt.tick_params(axis="both", labelsize=fontsize)
    plt.xlabel("Rank", fontsize=int(1.2 * fontsize))
    plt.ylabel(ylabel, fontsize=int(1.2 * fontsize))
    if elbow is not None:
        plt.plot(x[elbow - 1], mean[elbow - 1], "ro")
    plt.tight_layout()
    if filename is not None:
        plt.savefig(filename, dpi=300, bbox_inches="tight")
    return fig


def _group_colors(groups, cmap):
    """One colour per distinct metadata group, evenly spaced on ``cmap`` (plotting/aesthetics.py:get_colors_from_labels)."""
    from matplotlib import colormaps
    labels = sorted(set(groups), key=str)
    cm = colormaps[cmap]
    n = max(len(labels) - 1, 1)
    return {lab: cm(i / n) for i, lab in enumerate(labels)}


def tensor_factors_plot(interaction_tensor, order_labels=None, metadata=None, sample_col="Element", group_col="Category",
                        meta_cmaps=None, fontsize=20, plot_legend=True, filename=None):
    """Loadings of every factor (rows) in every tensor dimension (columns) as bar plots, bars coloured by the metadata group
    of their element (cell2cell/plotting/tensor_plot.py:14-104, the figure ``run_tensor_cell2cell_pipeline`` ends with).
    Element re-ordering / dimension sorting of the reference's plot are not offered.  Returns ``(fig, axes)``."""
    from matplotlib import pyplot as plt
    from matplotlib.patches import Patch
    factors = interaction_tensor.factors
    assert factors is not None, "First run the method 'compute_tensor_factorization' in your InteractionTensor"
    dim = len(factors)
    if order_labels is None:
        order_labels = list(factors.keys())
    assert dim == len(order_labels), "The lenght of factor_labels must match the order of the tensor (order {})".format(dim)
    rank = list(factors.values())[0].shape[1]
    if metadata is None:
        metadata = [None] * dim
    if meta_cmaps is None:
        meta_cmaps = ["gist_rainbow"] * dim
    assert len(metadata) == dim and len(meta_cmaps) == dim, "Provide a metadata and a cmap for each order"
    group_colors, element_colors = [], []
    for m, cmap in zip(metadata, meta_cmaps):
        if m is None or cmap is None:
            group_colors.append(None)
            element_colors.append(None)
            continue
        gc = _group_colors(m[group_col].tolist(), cmap)
        group_colors.append(gc)
        element_colors.append({e: gc[g] for e, g in zip(m[sample_col], m[group_col])})
    fig, axes = plt.subplots(nrows=rank, ncols=dim, figsize=(10, int(rank * 1.2 + 1)), sharex="col", squeeze=False)
    for d, df in enumerate(factors.values()):
        for r, name in enumerate(df.columns):
            ax = axes[r, d]
            colors = None if element_colors[d] is None else [element_colors[d].get(e, (0.5, 0.5, 0.5, 1.0)) for e in df.index]
            ax.bar(range(df.shape[0]), df[name].values.tolist(), color=colors)
            ax.spines["top"].set_visible(False)
            ax.spines["right"].set_visible(False)
            ax.tick_params(axis="x", which="both", length=0, labelbottom=False)
            ax.tick_params(axis="both", labelsize=fontsize)
            if d == 0:
                ax.set_ylabel(name, fontsize=int(1.2 * fontsize))
        axes[-1, d].set_xlabel(order_labels[d], fontsize=int(1.2 * fontsize), labelpad=fontsize)
    fig.align_ylabels(axes[:, 0])
    plt.tight_layout()
    if plot_legend:
        x0 = 1.05
        for d in range(dim):
            if group_colors[d] is None:
                continue
            handles = [Patch(facecolor=c, label=str(g)) for g, c in group_colors[d].items()]
            lgd = axes[0, -1].legend(handles=handles, title=order_labels[d], loc="upper left", bbox_to_anchor=(x0, 1.2),
                                     fontsize=fontsize, title_fontsize=fontsize, frameon=False)
            axes[0, -1].add_artist(lgd)
            x0 += 0.9
    if filename is not None:
        plt.savefig(filename, dpi=300, bbox_inches="tight")
    return fig, axes
