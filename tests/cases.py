"""Seeded small inputs that exercise the build quirks Q1-Q6 of SURVEY.md section 8a (shared by the golden generator and
the tests, so a fixture only has to store the reference's *outputs*)."""
import numpy as np
import pandas as pd


def toy_rnaseq():
    """The 6 x 5 toy matrix of the reference (cell2cell/datasets/toy_data.py:5-31) -- data values restated."""
    data = np.asarray([[5, 10, 8, 15, 2], [15, 5, 20, 1, 30], [18, 12, 5, 40, 20], [9, 30, 22, 5, 2],
                       [2, 1, 1, 27, 15], [30, 11, 16, 5, 12]])
    df = pd.DataFrame(data, index=["Protein-" + c for c in "ABCDEF"], columns=["C1", "C2", "C3", "C4", "C5"])
    df.index.name = "gene_id"
    return df


def toy_ppi(prot_complex=True):
    """The 10-row toy PPI list of the reference (cell2cell/datasets/toy_data.py:34-76)."""
    pairs = [("A", "B"), ("B", "C"), ("C", "A"), ("B", "B"), ("B", "A"), ("E", "F"), ("F", "F"),
             ("C&E" if prot_complex else "C", "F"), ("B", "E"), ("A&B" if prot_complex else "A", "F")]
    def nm(x):
        return "&".join("Protein-" + p for p in x.split("&"))
    ppi = pd.DataFrame([[nm(a), nm(b)] for a, b in pairs], columns=["A", "B"])
    return ppi.assign(score=1.0)


def hard_case(seed, n_contexts=4, n_genes=14, n_cells=5, n_lr=24, dtype=np.float64, fp32_values=False):
    """Mixed-case gene names, complexes on both sides (some with a subunit missing from some context or from the whole
    universe), a complex name used on both sides (the ``elif`` quirk), duplicated LR rows, contexts with different gene /
    cell sets and orders, real zeros, and a ``pathway`` column for grouping."""
    rs = np.random.RandomState(seed)
    genes = ["Gene%d" % i if i % 3 else "gene%d" % i for i in range(n_genes)]
    cells = ["T%d" % i for i in range(n_cells)]
    mats = []
    for c in range(n_contexts):
        p_drop = 0.15 if seed % 2 == 0 else 0.06
        g_keep = [g for g in genes if rs.rand() > p_drop]
        c_keep = [k for k in cells if rs.rand() > p_drop] or cells[:2]
        if c and rs.rand() < 0.5:
            g_keep = list(rs.permutation(g_keep))
        if c and rs.rand() < 0.5:
            c_keep = list(rs.permutation(c_keep))
        e = rs.gamma(0.7, 2.0, size=(len(g_keep), len(c_keep))) * (rs.rand(len(g_keep), len(c_keep)) > 0.25)
        if fp32_values:                      # float64 frames holding fp32-representable values (GPU parity inputs)
            e = e.astype(np.float32).astype(np.float64)
        mats.append(pd.DataFrame(e.astype(dtype), index=g_keep, columns=c_keep))
    def pick():
        r = rs.rand()
        if r < 0.6:
            return genes[rs.randint(n_genes)]
        if r < 0.9:
            a, b = rs.choice(n_genes, 2, replace=False)
            return genes[a] + "&" + genes[b]
        a, b, c3 = rs.choice(n_genes, 3, replace=False)
        return genes[a] + "&" + genes[b] + "&" + genes[c3]
    rows = [(pick(), pick()) for _ in range(n_lr)]
    rows.append((rows[0][1], rows[0][0]))          # same names on the other side (complex 'elif' quirk when complex)
    rows.append(rows[1])                            # duplicated LR row is kept
    rows.append(("GENE1&NOTAGENE", genes[2]))       # complex with a subunit nobody expresses
    ppi = pd.DataFrame(rows, columns=["A", "B"])
    ppi["score"] = 1.0
    ppi["pathway"] = ["P%d" % rs.randint(5) for _ in range(len(rows))]
    return mats, ppi


OPTION_GRID = [
    dict(how=how, communication_score=score, complex_agg_method=agg, group_ppi_by=grp, group_ppi_method=gm)
    for how in ("inner", "outer", "outer_genes", "outer_cells")
    for score in ("expression_product", "expression_mean", "expression_gmean")
    for agg in ("min", "mean", "gmean")
    for grp, gm in ((None, "gmean"), ("pathway", "gmean"), ("pathway", "sum"), ("pathway", "mean"))
]


def long_tables(seed, n_contexts=3, n_cells=5, n_lrs=8, dup=True):
    """Seeded long-format communication-score tables (one per context) for ``dataframes_to_tensor``: every context misses a
    cell type and a few LR pairs, LR pairs do not cover every cell pair, a few rows are duplicated."""
    import pandas as pd
    rs = np.random.RandomState(seed)
    cells = ["CT%d" % i for i in range(n_cells)]
    lrs = [("L%d" % i, "R%d" % (i % 3)) for i in range(n_lrs)]
    out = {}
    for c in range(n_contexts):
        ctx_cells = [x for i, x in enumerate(cells) if i != (c % n_cells)]
        ctx_lrs = [x for i, x in enumerate(lrs) if (i + c) % 4 != 0]
        rows = []
        for lig, rec in ctx_lrs:
            senders = [x for x in ctx_cells if rs.rand() > 0.15]
            receivers = [x for x in ctx_cells if rs.rand() > 0.15]
            for s in senders:
                for r in receivers:
                    if rs.rand() > 0.2:
                        rows.append((s, r, lig, rec, float(np.round(rs.rand() * (rs.rand() > 0.1), 4))))
        df = pd.DataFrame(rows, columns=["source", "target", "ligand", "receptor", "score"])
        if dup and len(df) > 4:
            extra = df.iloc[rs.choice(len(df), 4, replace=False)].copy()
            extra["score"] = np.round(extra["score"] + rs.rand(4), 4)
            df = pd.concat([df, extra], ignore_index=True)
        out["ctx%d" % (n_contexts - c)] = df.sample(frac=1.0, random_state=seed + c).reset_index(drop=True)
    return out


LONG_TABLE_OPTIONS = [dict(how=h, outer_fraction=f, lr_fill=lf, cell_fill=cf, dup_aggregation=d)
                      for h, f in (("inner", 0.0), ("outer", 0.0), ("outer", 0.6), ("outer_lrs", 0.0), ("outer_cells", 0.5))
                      for lf, cf in ((np.nan, np.nan), (0.0, np.nan), (np.nan, 0.0))
                      for d in ("max", "min")]


# ----------------------------------------------------------------------------------------------------------------------
# Coupled Tensor Component Analysis (cell2cell/tensor/coupled_factorization.py): seeded pairs of tensors sharing modes
# ----------------------------------------------------------------------------------------------------------------------
COUPLED_CASES = {
    # name: shape1, shape2, mode_mapping, rank, n_iter_max, tol, extra keyword arguments, masked tensors
    "same_order_int_mapping": dict(shape1=(4, 7, 5, 5), shape2=(4, 9, 5, 5), mode_mapping=1, rank=3, n_iter_max=25, tol=-1.0),
    "different_orders": dict(shape1=(6, 8, 5, 4), shape2=(6, 8, 7), mode_mapping={"shared": [(0, 0), (1, 1)]}, rank=4,
                             n_iter_max=20, tol=-1.0),
    "tensor2_all_shared": dict(shape1=(5, 10, 6, 6), shape2=(10, 6, 6), mode_mapping={"shared": [(1, 0), (2, 1), (3, 2)]},
                               rank=2, n_iter_max=20, tol=-1.0),
    "crossed_pairs": dict(shape1=(5, 6, 7), shape2=(7, 4, 5), mode_mapping={"shared": [(0, 2), (2, 0)]}, rank=3,
                          n_iter_max=15, tol=-1.0),
    "masked_first": dict(shape1=(4, 7, 5, 5), shape2=(4, 9, 5, 5), mode_mapping=[1], rank=3, n_iter_max=15, tol=-1.0,
                         masked=(True, False)),
    "masked_both": dict(shape1=(4, 6, 5, 5), shape2=(4, 5, 5), mode_mapping={"shared": [(0, 0), (2, 1), (3, 2)]}, rank=2,
                        n_iter_max=15, tol=-1.0, masked=(True, True)),
    "stop_abs": dict(shape1=(5, 12, 6, 6), shape2=(5, 9, 6, 6), mode_mapping=1, rank=3, n_iter_max=200, tol=2e-4),
    "stop_rec_unbalanced": dict(shape1=(5, 12, 6, 6), shape2=(5, 30, 6, 6), mode_mapping=1, rank=3, n_iter_max=200, tol=2e-4,
                                kwargs=dict(cvg_criterion="rec_error", balance_errors=False)),
    "stop_balanced_uneven": dict(shape1=(5, 12, 6, 6), shape2=(5, 30, 6, 6), mode_mapping=1, rank=3, n_iter_max=200, tol=2e-4),
}


def coupled_inputs(name, seed=11):
    """(x1, x2, mask1, mask2) of ``COUPLED_CASES[name]``: two noisy non-negative low-rank tensors whose shared modes have the
    same true factors; masked tensors carry zeros where the uint8 mask is 0 (what ``PreBuiltTensor`` hands over)."""
    case = COUPLED_CASES[name]
    s1, s2 = case["shape1"], case["shape2"]
    mm = case["mode_mapping"]
    if not isinstance(mm, dict):
        non_shared = [mm] if isinstance(mm, int) else list(mm)
        mm = {"shared": [(i, i) for i in range(len(s1)) if i not in non_shared]}
    rs = np.random.RandomState(seed)
    true_rank = 3
    f1 = [rs.gamma(1.0, 1.0, size=(s, true_rank)) for s in s1]
    f2 = [rs.gamma(1.0, 1.0, size=(s, true_rank)) for s in s2]
    for a, b in mm["shared"]:
        f2[b] = f1[a]
    out = []
    for fs in (f1, f2):
        letters = "abcdefgh"[: len(fs)]
        x = np.einsum(",".join(l + "r" for l in letters) + "->" + letters, *fs)
        x = x / x.mean() + 0.05 * rs.random_sample(x.shape)
        out.append(np.ascontiguousarray(x, dtype=np.float32))
    masks = [None, None]
    for k, flag in enumerate(case.get("masked", (False, False))):
        if flag:
            m = np.ones(out[k].shape, dtype=np.uint8)
            m[rs.rand(*out[k].shape) < 0.08] = 0
            m[1, :, 2] = 0
            masks[k] = m
            out[k] = out[k] * m
    return out[0], out[1], masks[0], masks[1]


# ----------------------------------------------------------------------------------------------------------------------
# Factorization-side API goldens (tests/golden/make_golden_ncp.py): the reference's own drivers on the oracle primitives
# ----------------------------------------------------------------------------------------------------------------------
def _lowrank32(shape, rank, seed, noise=0.05):
    rs = np.random.RandomState(seed)
    fs = [rs.gamma(1.0, 1.0, size=(s, rank)) for s in shape]
    letters = "abcdefgh"[: len(shape)]
    x = np.einsum(",".join(l + "r" for l in letters) + "->" + letters, *fs, optimize=True)
    x = x / x.mean() + noise * rs.random_sample(x.shape)
    return np.ascontiguousarray(x, dtype=np.float32)


_FIX = dict(n_iter_max=30, tol=-1.0, verbose=False)
NCP_API_CASES = {
    # BaseTensor.compute_tensor_factorization (tensor.py:216-362) through the reference's PreBuiltTensor
    "factorize_best_of_3": dict(kind="factorize", shape=(5, 40, 6, 6), masked=False,
                                kwargs=dict(rank=4, tf_type="non_negative_cp", init="random", random_state=3, runs=3, **_FIX)),
    "factorize_masked": dict(kind="factorize", shape=(5, 40, 6, 6), masked=True,
                             kwargs=dict(rank=3, tf_type="non_negative_cp", init="svd", random_state=1, runs=2, **_FIX)),
    "factorize_unordered": dict(kind="factorize", shape=(6, 30, 8, 8), masked=False,
                                kwargs=dict(rank=5, init="random", random_state=9, var_ordered_factors=False, **_FIX)),
    "factorize_raw_loadings": dict(kind="factorize", shape=(6, 30, 8, 8), masked=False,
                                   kwargs=dict(rank=3, init="random", random_state=2, normalize_loadings=False, **_FIX)),
    "factorize_hals": dict(kind="factorize", shape=(5, 40, 6, 6), masked=False,
                           kwargs=dict(rank=3, tf_type="non_negative_cp_hals", init="random", random_state=4, **_FIX)),
    "factorize_3way_stop_rule": dict(kind="factorize", shape=(30, 8, 7), masked=False,
                                     kwargs=dict(rank=3, init="random", random_state=6, n_iter_max=200, tol=1e-5, verbose=False)),
    # elbow drivers (factorization.py:186-373)
    "elbow_single": dict(kind="elbow_single", shape=(5, 40, 6, 6), masked=False,
                         kwargs=dict(upper_rank=5, init="random", random_state=5, **_FIX)),
    "elbow_single_masked": dict(kind="elbow_single", shape=(5, 40, 6, 6), masked=True,
                                kwargs=dict(upper_rank=4, init="random", random_state=2, **_FIX)),
    "elbow_multi_error": dict(kind="elbow_multi", shape=(5, 40, 6, 6), masked=False,
                              kwargs=dict(upper_rank=5, runs=3, init="random", random_state=5, metric="error", **_FIX)),
    "elbow_multi_masked": dict(kind="elbow_multi", shape=(5, 40, 6, 6), masked=True,
                               kwargs=dict(upper_rank=3, runs=2, init="random", random_state=7, metric="error", **_FIX)),
    "elbow_multi_similarity": dict(kind="elbow_multi", shape=(5, 40, 6, 6), masked=False,
                                   kwargs=dict(upper_rank=4, runs=3, init="random", random_state=5, metric="similarity", **_FIX)),
}


def ncp_api_inputs(name, seed=21):
    """(x fp32, mask uint8 or None, order_names) of ``NCP_API_CASES[name]``; a masked tensor holds zeros under its mask."""
    case = NCP_API_CASES[name]
    shape = case["shape"]
    x = _lowrank32(shape, 4, seed, noise=0.1)
    names = [["n%d_%d" % (k, i) for i in range(s)] for k, s in enumerate(shape)]
    mask = None
    if case["masked"]:
        rs = np.random.RandomState(seed + 1)
        mask = (rs.rand(*shape) > 0.1).astype(np.uint8)
        mask[1, :, 2] = 0
        x = x * mask
    return x, mask, names


NCP_FULL_CASES = {
    # BASELINE config 3 at full size: 13 x 1900 x 17 x 17, how='outer'-like mask (+ 2 % arbitrary entries), rank 10
    "config3_asd_masked": dict(shape=(13, 1900, 17, 17), true_rank=5, data_seed=3, rank=10, seed=888, n_iter=60, masked=True),
    # BASELINE config 4 at full size and its own rank: 60 x 2000 x 40 x 40, rank 20 (tensor-pipe passes)
    "config4_atlas_rank20": dict(shape=(60, 2000, 40, 40), true_rank=6, data_seed=5, rank=20, seed=888, n_iter=10, masked=False),
    # BASELINE config 5 family: K = 30 (in-plane row not a multiple of 4 floats: the VEC = 2 kernels), C and L scaled
    "config5_k30_rank12": dict(shape=(22, 870, 30, 30), true_rank=5, data_seed=7, rank=12, seed=888, n_iter=30, masked=False),
    "config5_k30_rank30": dict(shape=(22, 870, 30, 30), true_rank=5, data_seed=7, rank=30, seed=889, n_iter=15, masked=False),
}


def ncp_golden_inputs(name):
    """(x fp32, mask uint8 or None) of ``NCP_FULL_CASES[name]``."""
    case = NCP_FULL_CASES[name]
    shape = case["shape"]
    x = _lowrank32(shape, case["true_rank"], case["data_seed"])
    mask = None
    if case["masked"]:
        rs = np.random.RandomState(case["data_seed"] + 1)
        mask = np.ones(shape, dtype=np.uint8)
        mask[2::3, :, 5, :] = 0          # a cell type missing from every 3rd context (sender rows and receiver columns)
        mask[2::3, :, :, 5] = 0
        mask[rs.rand(shape[0], shape[1]) < 0.05] = 0   # gene pairs missing from a context: whole planes
        mask[rs.rand(*shape) < 0.02] = 0               # arbitrary user-masked entries
        x = x * mask
    return x, mask
