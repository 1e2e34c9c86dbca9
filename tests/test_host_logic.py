"""Host-side marshalling and bookkeeping (no GPU, no compute calls)."""
import ctypes

import numpy as np
import pytest
import torch

from daliid_b200 import _lib, sharded


def _read_f32(ptr, n):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n,)).copy()


def test_as_matrix_numpy_and_torch_views():
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    b = _lib.as_matrix(a)
    assert b.shape == (3, 4) and b.ld == 4 and b.device is None and b.ptr == a.ctypes.data
    # a float64 / Fortran-ordered array is converted to a C-contiguous fp32 copy that the Buf keeps alive
    c = _lib.as_matrix(np.asfortranarray(a.astype(np.float64)))
    assert c.shape == (3, 4) and c.ld == 4 and c.keep.dtype == np.float32
    np.testing.assert_array_equal(_read_f32(c.ptr, 12), a.reshape(-1))
    # torch: a column slice keeps its row pitch (no copy), a strided or transposed view is made contiguous
    t = torch.arange(40, dtype=torch.float32).reshape(5, 8)
    v = _lib.as_matrix(t[:, :6])
    assert v.shape == (5, 6) and v.ld == 8 and v.ptr == t.data_ptr()
    s = _lib.as_matrix(t[:, ::2])
    assert s.shape == (5, 4) and s.ld == 4 and s.keep.is_contiguous()
    tr = _lib.as_matrix(t.t())
    assert tr.shape == (8, 5) and tr.ld == 5
    np.testing.assert_array_equal(_read_f32(tr.ptr, 40), t.t().contiguous().reshape(-1).numpy())
    # dtype conversion, single row, empty matrices
    assert _lib.as_matrix(t.double()).keep.dtype == torch.float32
    assert _lib.as_matrix(t[:1]).ld == 8 and _lib.as_matrix(torch.zeros(0, 4)).ld == 4
    assert _lib.as_matrix(torch.zeros(3, 0)).ld == 1
    g = torch.ones(2, 3, requires_grad=True)
    assert not _lib.as_matrix(g).keep.requires_grad
    with pytest.raises(ValueError):
        _lib.as_matrix(torch.zeros(3))
    with pytest.raises(ValueError):
        _lib.as_matrix(np.zeros((2, 2, 2)))
    i = _lib.as_matrix(np.arange(6).reshape(2, 3), np.int32)
    assert i.keep.dtype == np.int32


def test_np_ptr_is_the_array_address():
    a = np.zeros(7, dtype=np.int32)
    assert _lib.np_ptr(a) == a.ctypes.data == _lib.p_i32(a)
    assert _lib.np_ptr(a[2:]) == a.ctypes.data + 8


def test_slab_bounds_partition():
    for G in (0, 1, 7, 15913, 62956):
        for world in (1, 2, 3, 8):
            slabs = [sharded.slab_bounds(G, world, r) for r in range(world)]
            assert slabs[0][0] == 0 and sum(s for _, s in slabs) == G
            for (a, n), (b, _) in zip(slabs, slabs[1:]):
                assert a + n == b
            sizes = [s for _, s in slabs]
            assert max(sizes) - min(sizes) <= 1


def test_share_queries_without_a_process_group_is_the_identity():
    q = torch.zeros(4, 8)
    assert sharded.share_queries(q, 0) is q
    qn = np.zeros((4, 8), dtype=np.float32)
    assert sharded.share_queries(qn, 0) is qn


def test_ensemble_argument_checks_need_no_gpu():
    """ensemble_distance_matrices validates its model list before touching the device."""
    import pytest
    from daliid_b200 import metrics
    a = np.zeros((4, 8), dtype=np.float32)
    b = np.zeros((6, 8), dtype=np.float32)
    with pytest.raises(ValueError):
        metrics.ensemble_distance_matrices([], [])
    with pytest.raises(ValueError):
        metrics.ensemble_distance_matrices([a, a], [b])
    with pytest.raises(ValueError):  # the models must share the query and gallery sets
        metrics.ensemble_distance_matrices([a, np.zeros((5, 8), dtype=np.float32)], [b, b])
    with pytest.raises(ValueError):  # feature dimensions of one model must agree
        metrics.ensemble_distance_matrices([a], [np.zeros((6, 9), dtype=np.float32)])
    with pytest.raises(ValueError):
        metrics.ensemble_distance_matrices([a] * 9, [b] * 9)


def test_label_canonicalisation_keeps_int32_and_densifies_strings():
    from daliid_b200 import metrics
    q = np.array([3, 1], dtype=np.int32)
    g = np.array([1, 3, 7], dtype=np.int32)
    cq, cg = metrics.canonicalize_labels(q, g)
    assert cq.dtype == np.int32 and np.array_equal(cq, q) and np.array_equal(cg, g)
    cq, cg = metrics.canonicalize_labels(np.array(["b", "a"]), np.array(["a", "b", "zz"]))
    assert cq.dtype == np.int32 and (cq[0] == cg[1]) and (cq[1] == cg[0]) and cg[2] not in (cq[0], cq[1])
    cq, cg = metrics.canonicalize_labels(np.array([2 ** 40, 5]), np.array([5, 2 ** 40, 9]))
    assert (cq[0] == cg[1]) and (cq[1] == cg[0]) and len({int(x) for x in cg}) == 3
