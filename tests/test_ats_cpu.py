"""CPU checks of the host side of adaptive token sampling (no kernels involved): the index stabilisation of the host mirror
against the oracle's restatement (reference blocks.py:378-391), and the constructor contract."""
import pytest
import torch

import eventful_oracle as orc
from eventful_transformer import blocks


def _random_sets(rows, n, k, gen):
    return torch.stack([torch.randperm(n, generator=gen)[:k].sort()[0] for _ in range(rows)])


@pytest.mark.parametrize("rows,n,k", [(2, 17, 12), (12, 197, 177), (3, 50, 1), (4, 40, 40)])
def test_index_stabilisation_matches_the_oracle(rows, n, k):
    gen = torch.Generator().manual_seed(rows * 1000 + n)
    blk = blocks.Block(dim=32, heads=2, input_size=(4, 4), mlp_ratio=2, ats_fraction=0.5)
    last_oracle = None
    for _ in range(6):
        new = _random_sets(rows, n, k, gen)
        got = blk._stabilize_ats_indices(new)
        want = orc.OracleBackbone._stabilize(last_oracle, new)
        assert torch.equal(got, want)
        # same token set as the new selection, and every token that stayed keeps its slot
        assert torch.equal(got.sort(dim=-1)[0], new)
        if last_oracle is not None:
            for r in range(rows):
                stay = torch.isin(last_oracle[r], new[r])
                assert torch.equal(got[r][stay], last_oracle[r][stay])
        blk.last_ats_indices = got
        last_oracle = want
    blk.reset()
    assert blk.last_ats_indices is None


def test_ats_constructor_contract():
    with pytest.raises(AssertionError):  # reference blocks.py:71-74
        blocks.Block(dim=32, heads=2, input_size=(4, 4), mlp_ratio=2, ats_fraction=0.5, window_size=(2, 2))
    with pytest.raises(AssertionError):
        blocks.Block(dim=32, heads=2, input_size=(4, 4), mlp_ratio=2, ats_fraction=1.5)
    with pytest.raises(NotImplementedError):
        blocks.Block(dim=32, heads=2, input_size=(4, 4), mlp_ratio=2, ats_fraction=0.5, relative_embedding_size=(4, 4))
    assert blocks.EventfulBlock(dim=32, heads=2, input_size=(4, 4), mlp_ratio=2, ats_fraction=0.7).ats_fraction == 0.7
