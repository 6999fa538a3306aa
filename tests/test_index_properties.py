"""CPU property tests (hypothesis) of the integer / index host logic against the oracle restatement: these ops must be
bit-exact for ANY shape, not only the golden fixtures' shapes."""
import torch
from hypothesis import given, settings, strategies as st

import dust3r_oracle as O
import uniception_b200 as U
from uniception_b200 import dp, engine as E


@settings(max_examples=60, deadline=None)
@given(b=st.integers(1, 3), h=st.integers(1, 9), w=st.integers(1, 9))
def test_grid_positions_equal_the_reference_position_getter(b, h, w):
    ours = E.grid_positions(b, h, w, "cpu")  # int32 [(b h w), 2]
    ref = O.patch_positions(b, h, w, "cpu")  # int64 [b, h*w, 2] (patch_embed.py:25-31)
    assert ours.dtype == torch.int32 and ours.shape == (b * h * w, 2)
    assert torch.equal(ours.long().view(b, h * w, 2), ref)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 40), data=st.data())
def test_feature_take_indices_matches_oracle(n, data):
    kind = data.draw(st.sampled_from(["none", "int", "list"]))
    if kind == "none":
        idx = None
    elif kind == "int":
        idx = data.draw(st.integers(1, n))  # last-n
    else:
        idx = data.draw(st.lists(st.integers(-n, n - 1), min_size=1, max_size=min(n, 6)))  # negatives allowed
    assert U.feature_take_indices(n, idx) == O.feature_take_indices(n, idx)


@settings(max_examples=100, deadline=None)
@given(names=st.lists(st.text(alphabet="abcd", min_size=1, max_size=2), min_size=1, max_size=8), mode=st.sampled_from(["sym", "plain", "single"]))
def test_is_symmetrized_matches_oracle(names, mode):
    if mode == "sym":  # (a,b),(b,a) pairs -> True
        inst1 = [x for a, b in zip(names, names[::-1]) for x in (a, b)]
        inst2 = [x for a, b in zip(names, names[::-1]) for x in (b, a)]
    elif mode == "single":  # batch size 1 is never symmetrized (factory/dust3r.py:25-26)
        inst1, inst2 = names[:1], names[:1]
    else:  # arbitrary even-length batches
        inst1, inst2 = (names + names)[: 2 * len(names)], (names[::-1] + names)[: 2 * len(names)]
    ours = U.is_symmetrized({"instance": inst1}, {"instance": inst2})  # the reference's dict signature
    assert ours == O.is_symmetrized(inst1, inst2)
    if mode == "sym":
        assert ours
    if mode == "single":
        assert not ours


@settings(max_examples=50, deadline=None)
@given(b=st.integers(1, 5), c=st.integers(1, 4), seed=st.integers(0, 1000))
def test_interleave_is_a_bit_exact_permutation(b, c, seed):
    g = torch.Generator().manual_seed(seed)
    t1, t2 = torch.randn(b, c, generator=g), torch.randn(b, c, generator=g)
    r1, r2 = U.interleave(t1, t2)
    o1, o2 = O.interleave(t1, t2)
    assert torch.equal(r1, o1) and torch.equal(r2, o2)
    assert torch.equal(r1[0::2], t1) and torch.equal(r1[1::2], t2) and torch.equal(r2[0::2], t2) and torch.equal(r2[1::2], t1)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 200), world=st.integers(1, 16))
def test_shard_batch_partitions_the_batch(n, world):
    parts = [dp.shard_batch(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))  # contiguous, no overlap, nothing dropped
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1
    if n % 2 == 0:  # symmetrized partners (2i, 2i+1) stay on one rank
        for r in range(world):
            lo, hi = dp.shard_pairs_symmetrized(n, r, world)
            assert lo % 2 == 0 and hi % 2 == 0


@settings(max_examples=40, deadline=None)
@given(b=st.integers(1, 2), c=st.integers(1, 3), h=st.integers(1, 4), w=st.integers(1, 4), p=st.sampled_from([1, 2, 4]), seed=st.integers(0, 99))
def test_pixel_shuffle_restatement_equals_torch(b, c, h, w, p, seed):
    x = torch.randn(b, c * p * p, h, w, generator=torch.Generator().manual_seed(seed))
    assert torch.equal(O.pixel_shuffle(x, p), torch.nn.functional.pixel_shuffle(x, p))  # linear.py:81-82
