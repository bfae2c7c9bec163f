"""Host half of the tensor build: all string / set logic of the reference, resolved to integer index arrays.

The device kernel (``c2c_build_ccc``) gathers, aggregates complexes, scores and stores; this module decides *which* rows
and columns it gathers, reproducing the reference's name handling exactly:

* ``complex_table`` restates ``get_genes_from_complexes`` (cell2cell/preprocessing/ppi.py:370-444);
* ``context_rows`` restates what ``add_complexes_to_expression`` (cell2cell/preprocessing/rnaseq.py:176-196) followed by
  the upper-casing and de-duplication of ``InteractionTensor.__init__`` (cell2cell/tensor/tensor.py:812-826) leave in each
  context's index -- as *descriptors* (which source rows a name aggregates), no values are touched;
* ``select_elements`` / ``ordered`` restate the gene / cell universe and ordering of ``build_context_ccc_tensor``
  (cell2cell/tensor/tensor.py:1083-1114, find_elements.py:32-76);
* ``filter_lr_pairs`` restates ``filter_ppi_by_proteins`` / ``filter_complex_ppi_by_proteins`` (ppi.py:185-244, 447-523)
  including the ``elif`` at :511-515;
* ``resolve`` turns (context, LR pair, side) into the ``first`` / ``count`` / ``sub_rows`` arrays of the C ABI.
"""
from collections import Counter, OrderedDict

import numpy as np

__all__ = ["complex_table", "ContextRows", "context_rows", "context_tables", "plain_rows", "select_elements", "ordered",
           "filter_lr_pairs", "resolve", "group_rows"]

ZERO_ROW = "zero"       # complex with a subunit missing from the context: a row of zeros (rnaseq.py:195)


def complex_table(ppi_data, complex_sep, interaction_columns=("A", "B")):
    """dict complex name -> set of subunit names, in first-appearance order over the PPI rows (ppi.py:426-442);
    also returns the plain names and subunit unions of either column."""
    col_a, col_b = interaction_columns
    plain = (set(), set())
    subs = (set(), set())
    complexes = OrderedDict()
    for a, b in zip(ppi_data[col_a].tolist(), ppi_data[col_b].tolist()):
        for side, name in enumerate((a, b)):
            if complex_sep in name:
                parts = set(name.split(complex_sep))
                complexes[name] = parts
                subs[side].update(parts)
            else:
                plain[side].add(name)
    return plain[0], subs[0], plain[1], subs[1], complexes


class ContextRows:
    """The index of one context's (complex-augmented) expression matrix: ``names`` in order, and for every name the source
    rows of the original matrix it aggregates (``rows[name]`` = list of ints) or ``ZERO_ROW``."""

    def __init__(self, names, rows):
        self.names = names
        self.rows = rows

    def __contains__(self, name):
        return name in self.rows


def plain_rows(index):
    """Index of a matrix used as is (direct ``build_context_ccc_tensor`` call): first occurrence of every name."""
    rows = OrderedDict()
    for i, name in enumerate(index):
        if name not in rows:
            rows[name] = [i]
    return ContextRows(list(rows.keys()), rows)


def context_rows(index, complexes, upper_letter_comparison):
    """Index after ``add_complexes_to_expression`` (one appended row per complex, evaluated against the original-case
    index -- tensor.py:812-816), upper-casing (:821-823) and ``~index.duplicated(keep='first')`` (:826)."""
    names = list(index)
    descr = [[i] for i in range(len(names))]
    where = {}
    for i, n in enumerate(names):
        where.setdefault(n, []).append(i)
    if complexes:
        for cname, parts in complexes.items():
            parts = list(parts)
            if all(p in where for p in parts):
                src = []
                for p in parts:
                    for pos in where[p]:
                        d = descr[pos]
                        if d == ZERO_ROW:
                            src = ZERO_ROW      # a subunit that is itself a zero-filled complex row (cannot happen with a
                            break               # sane separator; kept for totality)
                        src = src + d
                    if src == ZERO_ROW:
                        break
                d_new = src
            else:
                d_new = ZERO_ROW
            if cname in where:                  # DataFrame.loc[existing] = values overwrites every row of that name
                for pos in where[cname]:
                    descr[pos] = d_new
            else:
                where[cname] = [len(names)]
                names.append(cname)
                descr.append(d_new)
    if upper_letter_comparison:
        names = [n.upper() for n in names]
    rows = OrderedDict()
    for n, d in zip(names, descr):
        if n not in rows:
            rows[n] = d
    return ContextRows(list(rows.keys()), rows)


def context_tables(indexes, complexes, upper_letter_comparison, keep=8):
    """``context_rows`` of every context; contexts whose indexes are equal (the usual case: one gene universe for every
    sample) share ONE ``ContextRows`` object -- the name handling is 13 ms of Python per 4000-gene context otherwise."""
    seen, tables = [], []
    for idx in indexes:
        hit = None
        for prev, tab in seen:
            if prev is idx or (len(prev) == len(idx) and prev.equals(idx)):
                hit = tab
                break
        if hit is None:
            hit = context_rows(list(idx), complexes, upper_letter_comparison)
            seen.insert(0, (idx, hit))
            del seen[keep:]
        tables.append(hit)
    return tables


def select_elements(lists, outer, fraction):
    """``inner``: intersection; ``outer``: names present in at least ``fraction`` of the contexts
    (find_elements.py:32-76: abundance = count / n_contexts, kept when >= fraction)."""
    if not outer:
        return set.intersection(*[set(l) for l in lists])
    counts = Counter()
    for l in lists:
        counts.update(set(l))
    total = len(lists)
    return set(k for k, c in counts.items() if c / total >= fraction)


def ordered(first_list, selected):
    """tensor.py:1106-1114: keep context 0's order when the selected set equals context 0's set, else ``sorted``."""
    return list(first_list) if set(first_list) == selected else sorted(selected)


def filter_lr_pairs(ppi_data, proteins, complex_sep=None, upper_letter_comparison=True, interaction_columns=("A", "B")):
    """The reference's LR filter.  Returns the filtered DataFrame (index reset, names upper-cased when asked)."""
    col_a, col_b = interaction_columns
    ppi = ppi_data.copy()
    if upper_letter_comparison:
        ppi[col_a] = [str(x).upper() for x in ppi[col_a]]
        ppi[col_b] = [str(x).upper() for x in ppi[col_b]]
        prots = set(str(p).upper() for p in proteins)
    else:
        prots = set(proteins)
    if complex_sep is None:
        keep = ppi[col_a].isin(prots) & ppi[col_b].isin(prots)
        return ppi[keep].reset_index(drop=True)
    plain_a, sub_a, plain_b, sub_b, complexes = complex_table(ppi, complex_sep, interaction_columns)
    ok_a = set(plain_a & prots)
    ok_b = set(plain_b & prots)
    sub_ok_a = sub_a & prots
    sub_ok_b = sub_b & prots
    for name, parts in complexes.items():
        if all(p in sub_ok_a for p in parts):
            ok_a.add(name)
        elif all(p in sub_ok_b for p in parts):
            ok_b.add(name)
    keep = ppi[col_a].isin(ok_a) & ppi[col_b].isin(ok_b)
    return ppi.loc[keep].reset_index(drop=True)


def resolve(tables, columns, genes, cells, ligands, receptors):
    """Index arrays of ``c2c_build_ccc``.

    tables[c]: ContextRows; columns[c]: list of that context's cell-type names; ``genes`` / ``cells``: the tensor's gene
    universe and cell order; ``ligands`` / ``receptors``: names per LR pair.  A partner that is not in ``genes`` raises
    ``KeyError`` (the reference's ``.loc`` does); one that is in ``genes`` but not in the context is NaN (count 0).
    Returns dict(cell_col, a_first, a_count, b_first, b_count, sub_rows)."""
    C, L, K = len(tables), len(ligands), len(cells)
    gene_set = set(genes)
    uniq = OrderedDict()
    for n in list(ligands) + list(receptors):
        if n not in uniq:
            if n not in gene_set:
                raise KeyError(n)
            uniq[n] = len(uniq)
    U = len(uniq)
    lig_id = np.fromiter((uniq[n] for n in ligands), dtype=np.int64, count=L)
    rec_id = np.fromiter((uniq[n] for n in receptors), dtype=np.int64, count=L)
    first_u = np.zeros((C, max(U, 1)), dtype=np.int32)
    count_u = np.zeros((C, max(U, 1)), dtype=np.int32)
    sub_rows = []
    n_sub = 0
    done = {}                                              # contexts that share a ContextRows object (same index) share its
    for c, tab in enumerate(tables):                       # first / count rows and its sub_rows (row numbers are per matrix)
        prev = done.get(id(tab))
        if prev is not None:
            first_u[c], count_u[c] = first_u[prev], count_u[prev]
            continue
        done[id(tab)] = c
        for n, u in uniq.items():
            d = tab.rows.get(n)
            if d is None:
                continue                                   # absent from this context -> NaN
            if d == ZERO_ROW:
                count_u[c, u] = -1
                continue
            first_u[c, u] = n_sub
            count_u[c, u] = len(d)
            sub_rows.extend(d)
            n_sub += len(d)
    cell_col = np.full((C, max(K, 1)), -1, dtype=np.int32)
    for c, cols in enumerate(columns):
        pos = {}
        for j, name in enumerate(cols):
            pos.setdefault(name, j)
        for k, name in enumerate(cells):
            cell_col[c, k] = pos.get(name, -1)
    return dict(cell_col=cell_col[:, :K],
                a_first=first_u[:, lig_id].reshape(-1), a_count=count_u[:, lig_id].reshape(-1),
                b_first=first_u[:, rec_id].reshape(-1), b_count=count_u[:, rec_id].reshape(-1),
                sub_rows=np.asarray(sub_rows, dtype=np.int32))


def group_rows(ppi, group_ppi_by):
    """``ppi.groupby(group_ppi_by)`` (tensor.py:1245-1252): sorted group keys and, per group, the LR row numbers."""
    keys, off, rows = [], [0], []
    for key, sub in ppi.groupby(group_ppi_by):
        keys.append(key)
        rows.extend(int(i) for i in sub.index)
        off.append(len(rows))
    return keys, np.asarray(off, dtype=np.int32), np.asarray(rows, dtype=np.int32)
