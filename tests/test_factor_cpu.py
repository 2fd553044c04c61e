"""Host setup code of the direct solver on the CPU: nested-dissection ordering (csrc/ordering.cpp) + supernodal
inverse-multifrontal Cholesky (csrc/direct_factor.cpp), applied with plain loops by tests/hostcheck/factorcheck.cpp
and compared with scipy's sparse LU on the same system matrix A = M + dt^2 D^T W^2 D."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import scenes

FC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck", "libfactorcheck.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def tet_system(N, k=1e5, dt=0.04):
    x, tets = scenes.kuhn_cube(N)
    n = len(x)
    m = scenes.density_weighted_tet_mass(x, tets, 1000.0)
    v = x[tets]
    edges = np.stack([v[:, 1] - v[:, 0], v[:, 2] - v[:, 0], v[:, 3] - v[:, 0]], axis=2)
    Dm = np.array([[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    B = np.einsum("ck,tkr->tcr", Dm, np.linalg.inv(edges))
    w2 = k * scenes.tet_volumes(x, tets) * dt * dt
    K = np.einsum("t,tar,tbr->tab", w2, B, B)
    I = np.repeat(tets, 4, axis=1).reshape(-1)
    J = np.tile(tets, (1, 4)).reshape(-1)
    A = (sp.coo_matrix((K.reshape(-1), (I, J)), shape=(n, n)).tocsr() + sp.diags(m)).tocsr()
    A.sort_indices()
    return x, A


@pytest.mark.parametrize("N,leaf", [(2, 32), (3, 8), (6, 32), (12, 32), (12, 4)])
def test_supernodal_factor_solves_the_system(N, leaf):
    if not os.path.exists(FC):
        pytest.skip("tests/hostcheck not built (run __graft_entry__.build())")
    fc = C.CDLL(FC)
    fc.fc_solve.argtypes = [C.c_int, _ip, _ip, _dp, _dp, C.c_int, _dp, _dp, C.POINTER(C.c_long), _dp]
    x, A = tet_system(N)
    n = A.shape[0]
    b = np.random.default_rng(N).standard_normal((n, 3))
    xs = np.zeros((n, 3))
    info = (C.c_long * 4)()
    sec = np.zeros(2)
    rc = fc.fc_solve(n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, np.ascontiguousarray(x), leaf, b, xs, info, sec)
    assert rc == 0
    ref = spl.spsolve(A.tocsc(), b)
    err = np.linalg.norm(xs - ref) / np.linalg.norm(ref)
    print(f"N={N} leaf={leaf}: n={n} supernodes={info[0]} levels={info[1]} nnz(L)={info[2]} max width={info[3]} rel err {err:.1e}")
    assert err < 1e-12


def test_indefinite_matrix_is_rejected():
    if not os.path.exists(FC):
        pytest.skip("tests/hostcheck not built")
    fc = C.CDLL(FC)
    fc.fc_solve.argtypes = [C.c_int, _ip, _ip, _dp, _dp, C.c_int, _dp, _dp, C.POINTER(C.c_long), _dp]
    x, A = tet_system(2)
    A = (A - sp.diags(np.full(A.shape[0], 1e9))).tocsr()
    A.sort_indices()
    n = A.shape[0]
    b = np.ones((n, 3))
    xs = np.zeros((n, 3))
    rc = fc.fc_solve(n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, np.ascontiguousarray(x), 32, b, xs, None, np.zeros(2))
    assert rc == -1
