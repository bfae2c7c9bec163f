"""Seeded synthetic inputs of the shapes BASELINE.json names.

The reference ships ``cell2cell/datasets/random_data.py:13-137`` (normal-distributed TPM matrices + resampled PPI lists)
for the same purpose; these generators produce the sparse, non-negative, complex-bearing inputs the benchmark configs
call for, byte-identical for the GPU path, the oracle and the CPU baseline.
"""
import numpy as np
import pandas as pd

__all__ = ["synthetic_contexts", "synthetic_ppi", "synthetic_case", "synthetic_lowrank_tensor", "CONFIGS"]

# name -> (contexts C, LR pairs L, cell types K, how)
CONFIGS = {
    "toy":     dict(C=4,   L=300,  K=6,  how="inner"),
    "balf":    dict(C=12,  L=1054, K=6,  how="inner"),
    "asd":     dict(C=13,  L=1900, K=17, how="outer"),
    "atlas":   dict(C=60,  L=2000, K=40, how="inner"),
    "spatial": dict(C=500, L=1000, K=30, how="inner"),
}


def synthetic_ppi(n_lr, n_genes, complex_fraction=0.1, seed=7, groups=0):
    """``n_lr`` unique (ligand, receptor) pairs over genes ``G00000..``; a fraction of the receptors are 2-subunit
    complexes ``Ga&Gb``.  ``groups`` > 0 adds a ``pathway`` column with that many labels (for ``group_ppi_by``)."""
    rs = np.random.RandomState(seed)
    names = np.array(["G%05d" % i for i in range(n_genes)])
    seen = set()
    lig, rec = [], []
    while len(lig) < n_lr:
        a, b = rs.randint(0, n_genes, size=2)
        if a == b or (a, b) in seen:
            continue
        seen.add((a, b))
        lig.append(names[a])
        if rs.rand() < complex_fraction:
            c = rs.randint(0, n_genes)
            while c == b:
                c = rs.randint(0, n_genes)
            rec.append(names[b] + "&" + names[c])
        else:
            rec.append(names[b])
    ppi = pd.DataFrame({"A": lig, "B": rec})
    ppi["score"] = 1.0
    if groups:
        ppi["pathway"] = ["P%03d" % g for g in rs.randint(0, groups, size=n_lr)]
    return ppi


def synthetic_contexts(n_contexts, n_genes, n_cells, outer=False, seed=1000, dtype=np.float32):
    """Per context ``gamma(0.5, 1) * (U > 0.3)`` (sparse, non-negative, real zeros).  With ``outer`` every 3rd context
    loses one cell type and every 4th loses 2 % of its genes (-> NaN / mask = 0 under how='outer')."""
    genes = ["G%05d" % i for i in range(n_genes)]
    cells = ["CT%02d" % i for i in range(n_cells)]
    out = []
    for c in range(n_contexts):
        rs = np.random.RandomState(seed + c)
        e = rs.gamma(0.5, 1.0, size=(n_genes, n_cells)) * (rs.rand(n_genes, n_cells) > 0.3)
        df = pd.DataFrame(e.astype(dtype), index=genes, columns=cells)
        if outer and c % 3 == 2:
            df = df.drop(columns=[cells[(c // 3) % n_cells]])
        if outer and c % 4 == 3:
            drop = rs.choice(n_genes, size=max(1, n_genes // 50), replace=False)
            df = df.drop(index=[genes[i] for i in drop])
        out.append(df)
    return out


def synthetic_case(name, scale=1.0):
    """(rnaseq_matrices, ppi_data, kwargs) for one of ``CONFIGS``; ``scale`` < 1 shrinks L (and G) for CPU-sized tests."""
    cfg = CONFIGS[name]
    n_lr = max(4, int(round(cfg["L"] * scale)))
    n_genes = 2 * n_lr
    ppi = synthetic_ppi(n_lr, n_genes)
    mats = synthetic_contexts(cfg["C"], n_genes, cfg["K"], outer=(cfg["how"] != "inner"))
    kwargs = dict(how=cfg["how"], communication_score="expression_gmean", complex_sep="&", complex_agg_method="min",
                  verbose=False)
    return mats, ppi, kwargs


def synthetic_lowrank_tensor(shape, rank=5, noise=0.05, seed=1, dtype=np.float32):
    """Rank-``rank`` gamma CP tensor + uniform noise: the factorization-only generator."""
    rs = np.random.RandomState(seed)
    factors = [rs.gamma(1.0, 1.0, size=(s, rank)) for s in shape]
    letters = "abcdefgh"[: len(shape)]
    expr = ",".join(l + "r" for l in letters) + "->" + letters
    x = np.einsum(expr, *factors, optimize=True)
    x = x / x.mean()
    x += noise * rs.random_sample(x.shape)
    return np.ascontiguousarray(x, dtype=dtype)
