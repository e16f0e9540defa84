"""Property tests of the packed layout (hypothesis): the oracle, the host mirror and the CPU arm agree with
each other for random shapes, pack/unpack are inverses, and shards of the packed tensor are the packed shards."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import llm_awq_b200 as P
from llm_awq_b200 import tp
from oracle import cpu_path
from oracle import w4a16_oracle as O

shapes = st.tuples(st.integers(1, 12).map(lambda v: 4 * v), st.integers(1, 6).map(lambda v: 64 * v))


@settings(max_examples=40, deadline=None)
@given(shapes, st.integers(0, 2 ** 31 - 1))
def test_pack_unpack_roundtrip_and_agreement(shape, seed):
    N, K = shape
    q = np.random.default_rng(seed).integers(0, 16, (N, K)).astype(np.int32)
    po = O.pack_intweight(q)
    ph = P.pack_intweight(torch.from_numpy(q)).numpy()
    assert np.array_equal(po, ph)
    assert np.array_equal(O.unpack_intweight(po), q)
    assert np.array_equal(O.unpack_intweight_indexed(po), q)
    assert np.array_equal(P.unpack_intweight(torch.from_numpy(po)).numpy(), q)
    assert np.array_equal(cpu_path.unpack_intweight(torch.from_numpy(po)).numpy(), q)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 4).map(lambda v: 8 * v), st.integers(1, 3), st.sampled_from([1, 2, 4]), st.integers(0, 2 ** 31 - 1))
def test_packed_shards_are_shards_of_the_unpacked_matrix(n_per, g_per, world, seed):
    """Slicing the packed tensor (tp.shard_column / shard_row) == packing the slice."""
    N, K = n_per * world, 128 * g_per * world
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, (N, K)).astype(np.int32)
    packed = torch.from_numpy(O.pack_intweight(q))
    rows = O.scale_rows(K)
    s = torch.from_numpy(rng.random((rows, N)).astype(np.float16))
    z = torch.from_numpy(rng.random((rows, N)).astype(np.float16))
    for r in range(world):
        qw, sc, sz, _ = tp.shard_column(packed, s, z, None, r, world)
        assert np.array_equal(O.unpack_intweight(qw.numpy()), q[r * n_per:(r + 1) * n_per])
        assert torch.equal(sc, s[:, r * n_per:(r + 1) * n_per])
        qw, sc, sz = tp.shard_row(packed, s, z, r, world)
        k = K // world
        assert np.array_equal(O.unpack_intweight(qw.numpy()), q[:, r * k:(r + 1) * k])
        assert torch.equal(sc[:k // 128], s[r * (k // 128):(r + 1) * (k // 128)]) and sc.shape[0] == O.scale_rows(k)
        assert torch.equal(sz[:k // 128], z[r * (k // 128):(r + 1) * (k // 128)])
