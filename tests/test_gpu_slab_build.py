"""Per-context slab build (tensor build row): ``InteractionTensor(..., context_slab=(lo, hi))`` builds
exactly rows lo..hi-1 of the context mode of the full tensor -- same genes / cells / LR pairs, bit-identical values and mask
-- and feeds the context-sharded factorization."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from tests.cases import OPTION_GRID, hard_case  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("opt", [OPTION_GRID[0], OPTION_GRID[13], OPTION_GRID[40], OPTION_GRID[75], OPTION_GRID[-1]])
def test_slabs_concatenate_to_the_full_build(opt):
    from cell2cell_b200.tensor import InteractionTensor
    mats, ppi = hard_case(3, n_contexts=7, fp32_values=True)
    kw = dict(complex_sep="&", verbose=False, **opt)
    names = ["ctx%d" % i for i in range(7)]
    full = InteractionTensor(mats, ppi, context_names=names, **kw)
    parts = [InteractionTensor(mats, ppi, context_names=names, context_slab=s, **kw) for s in ((0, 2), (2, 3), (3, 7))]
    assert torch.equal(torch.cat([p.tensor for p in parts], 0), full.tensor)
    if full.mask is None:
        assert all(p.mask is None for p in parts)
    else:
        assert torch.equal(torch.cat([p.mask for p in parts], 0), full.mask)
    for p, (lo, hi) in zip(parts, ((0, 2), (2, 3), (3, 7))):
        assert p.order_names[0] == names[lo:hi] and p.order_names[1:] == full.order_names[1:]
        assert p.genes == full.genes and p.cells == full.cells and p.full_contexts == 7 and p.context_slab == (lo, hi)
    with pytest.raises(ValueError):
        InteractionTensor(mats, ppi, context_slab=(3, 3), **kw)


def test_slab_feeds_the_sharded_factorization_world_of_one():
    """One process: the 'slab' is the whole tensor; the sharded driver must give the plain API's result."""
    from cell2cell_b200.sharded import build_sharded_interaction_tensor, compute_sharded_tensor_factorization
    from cell2cell_b200.synthetic import synthetic_case
    from cell2cell_b200.tensor import InteractionTensor
    mats, ppi, kw = synthetic_case("toy", scale=0.3)
    t = build_sharded_interaction_tensor(mats, ppi, **kw)
    errs = compute_sharded_tensor_factorization(t, 4, random_state=888, n_iter_max=30, tol=-1.0)
    ref = InteractionTensor(mats, ppi, **kw)
    ref.compute_tensor_factorization(rank=4, init="random", random_state=888, n_iter_max=30, tol=-1.0)
    assert len(errs) == 30 and list(t.factors.keys()) == list(ref.factors.keys())
    for (_, a), (_, b) in zip(t.factors.items(), ref.factors.items()):
        assert list(a.index) == list(b.index)
        assert np.linalg.norm(a.values - b.values) / np.linalg.norm(b.values) < 1e-4
    np.testing.assert_allclose(t.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-4)
