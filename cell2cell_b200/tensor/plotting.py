"""Elbow figures for ``elbow_rank_selection(output_fig=True)`` (cell2cell/plotting/tensor_plot.py:489-611).

matplotlib is optional (it is not installed in the build image): the import happens on first use, and the factorization
path never needs it -- pass ``output_fig=False`` on a box without it.
"""
import numpy as np

__all__ = ["require_matplotlib", "plot_elbow", "plot_multiple_run_elbow"]


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
