"""Host-side logic of the rulebook object (no GPU): pair lists of a table-driven SubM rulebook are built
on first access, once, and the reference-style 5-tuple stored under ``indice_key`` stays lazy."""
import torch


def test_rulebook_pair_lists_are_lazy_and_built_once():
    from ddf_b200.ops.spconv.ops import Rulebook
    calls = []

    def build():
        calls.append(1)
        return torch.zeros(27, 2, 5, dtype=torch.int32), torch.ones(27, dtype=torch.int32)

    gather = torch.full((5, 27), -1, dtype=torch.int32)
    rb = Rulebook(torch.zeros(5, 4, dtype=torch.int32), None, None, gather, gather.clone(), [8, 8, 8], kvol=27,
                  build_pairs=build)
    assert rb.kvol == 27 and not calls
    assert rb.indice_pairs.shape == (27, 2, 5) and calls == [1]
    assert int(rb.indice_pair_num.sum()) == 27 and calls == [1]      # second access: no rebuild


def test_eager_rulebook_keeps_reference_format():
    from ddf_b200.ops.spconv.ops import Rulebook
    pairs, num = torch.zeros(3, 2, 7, dtype=torch.int32), torch.zeros(3, dtype=torch.int32)
    rb = Rulebook(torch.zeros(4, 4, dtype=torch.int32), pairs, num, None, None, [4, 4, 4])
    assert rb.kvol == 3 and rb.indice_pairs is pairs and rb.indice_pair_num is num and rb.subm is False


def test_indice_tuple_is_lazy_and_indexable():
    from ddf_b200.ops.spconv.conv import _IndiceTuple
    from ddf_b200.ops.spconv.ops import Rulebook
    calls = []

    def build():
        calls.append(1)
        return torch.zeros(27, 2, 5, dtype=torch.int32), torch.ones(27, dtype=torch.int32)

    idx = torch.zeros(5, 4, dtype=torch.int32)
    rb = Rulebook(idx, None, None, torch.zeros(5, 27, dtype=torch.int32), None, [8, 8, 8], kvol=27, build_pairs=build)
    tup = _IndiceTuple(rb, idx, [8, 8, 8])
    assert len(tup) == 5 and not calls
    outids, indices, pairs, num, shape = tup              # what SparseInverseConv3d unpacks (conv.py:181-186)
    assert calls == [1] and pairs.shape == (27, 2, 5) and shape == [8, 8, 8] and tup[0] is idx
