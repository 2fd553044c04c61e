"""ONE mesh over several ranks -- the host-side plans (csrc/dist_plan.cpp), checked on the CPU for world sizes 2 .. 8
(the GPU side of the same paths is tests/test_dist_gpu.py, which needs two or more devices).

* sharded direct solve: `shard_owners` maps subtrees of the supernodal elimination tree to ranks.  tests/hostcheck/
  distplan_check.cpp emulates the sharded solve with plain loops -- own subtrees forward, top rows summed over the ranks, top
  forward / backward by everyone, own subtrees backward -- and the result must equal scipy's solution of the same system; the
  mapping must be a valid cut (top closed under `parent`, every subtree below it wholly owned by one rank, no row structure
  reaching into another rank's subtree) and reasonably balanced.
* partitioned PCG rows: `plan_halo` -- what rank r sends to q is exactly what q expects from r, in the same order, and a
  matrix-vector product assembled from the owned chunk + the received slots through the remapped columns is A u.
"""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from test_factor_cpu import tet_system

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck", "libdistplancheck.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _lib():
    if not os.path.exists(LIB):
        pytest.skip("tests/hostcheck not built (run __graft_entry__.build())")
    L = C.CDLL(LIB)
    L.dp_shard_solve.argtypes = [C.c_int, _ip, _ip, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _ip, _ip, C.c_int, _dp]
    L.dp_halo.argtypes = [C.c_int, _ip, _ip, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip, C.c_int, _ip, C.c_long]
    L.dp_halo.restype = C.c_long
    return L


@pytest.mark.parametrize("N,leaf,world", [(12, 32, 2), (12, 32, 3), (12, 32, 4), (12, 32, 8), (16, 64, 8), (8, 16, 5), (3, 8, 4), (2, 32, 2)])
def test_sharded_direct_solve_plan(N, leaf, world):
    L = _lib()
    x, A = tet_system(N)
    n = A.shape[0]
    b = np.random.default_rng(100 * N + world).standard_normal((n, 3))
    xs = np.zeros((n, 3))
    cap = n
    owner = np.full(cap, -9, np.int32)
    parent = np.full(cap, -9, np.int32)
    stats = np.zeros(8)
    rc = L.dp_shard_solve(n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, np.ascontiguousarray(x), leaf, world, b, xs,
                          owner, parent, cap, stats)
    assert rc == 0
    nb = int(stats[0])
    owner, parent = owner[:nb], parent[:nb]
    ref = spl.spsolve(A.tocsc(), b)
    err = np.linalg.norm(xs - ref) / np.linalg.norm(ref)
    ranks_used = len(set(owner[owner >= 0].tolist()))
    print(f"N={N} leaf={leaf} world={world}: {nb} supernodes, top {100 * stats[1]:.1f} % of the factor, heaviest / lightest rank "
          f"{100 * stats[2]:.1f} / {100 * stats[3]:.1f} %, ranks used {ranks_used}, emulated sharded solve vs scipy {err:.1e}")
    assert err < 1e-12
    assert stats[4] == 0                                    # no row structure reaches into another rank's subtree
    assert ((owner >= -1) & (owner < world)).all()
    for J in range(nb):                                      # a valid cut of the tree
        p = parent[J]
        if owner[J] == -1:
            assert p < 0 or owner[p] == -1                   # the top is closed under `parent`
        elif p >= 0 and owner[p] != -1:
            assert owner[p] == owner[J]                      # below the cut a subtree has ONE owner
    if nb >= 8 * world:                                      # a tree with enough subtrees is really cut, and evenly
        assert ranks_used == world
        assert stats[1] < 0.7
        assert stats[2] <= 2.5 * max(stats[3], 1e-9) + 0.05
    else:                                                    # degenerate trees: everything may stay replicated
        assert ranks_used in (0, world) or ranks_used <= world


def _pattern(kind, n, rng):
    if kind == "cube":
        _, A = tet_system(6)
        return A
    # random structurally symmetric pattern with a full diagonal
    B = sp.random(n, n, density=6.0 / n, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr")
    A = (B + B.T + sp.eye(n)).tocsr()
    A.data[:] = rng.standard_normal(A.nnz)
    A.sort_indices()
    return A


@pytest.mark.parametrize("kind", ["cube", "random"])
@pytest.mark.parametrize("world", [2, 3, 5, 8])
def test_partitioned_pcg_halo_plan(kind, world):
    L = _lib()
    rng = np.random.default_rng(7 * world + len(kind))
    A = _pattern(kind, 300, rng)
    n = A.shape[0]
    Ap, Ai = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    chunk = (n + world - 1) // world
    npad = chunk * world
    plans = []
    for r in range(world):
        send_cnt = np.zeros(world, np.int32); recv_cnt = np.zeros(world, np.int32)
        send_idx = np.zeros(n, np.int32); recv_idx = np.zeros(n, np.int32)
        r0, r1 = min(n, r * chunk), min(n, r * chunk + chunk)
        nnz = int(Ap[r1] - Ap[r0])
        cols = np.zeros(max(nnz, 1), np.int32)
        e = L.dp_halo(n, Ap, Ai, chunk, world, r, send_cnt, recv_cnt, send_idx, recv_idx, n, cols, max(nnz, 1))
        assert e == nnz
        so = np.concatenate([[0], np.cumsum(send_cnt)]); ro = np.concatenate([[0], np.cumsum(recv_cnt)])
        plans.append(dict(r0=r0, r1=r1, send=[send_idx[so[q]:so[q + 1]].copy() for q in range(world)],
                          recv=[recv_idx[ro[q]:ro[q + 1]].copy() for q in range(world)], ro=ro, cols=cols[:nnz].copy()))
    moved = 0
    for r in range(world):
        assert len(plans[r]["send"][r]) == 0 and len(plans[r]["recv"][r]) == 0
        for q in range(world):
            # what r sends to q is what q expects from r, element for element
            assert np.array_equal(plans[r]["send"][q], plans[q]["recv"][r])
            s = plans[r]["send"][q]
            assert np.all((s >= plans[r]["r0"]) & (s < plans[r]["r1"])) and np.all(np.diff(s) > 0)
            moved += len(s)
    # one exchange + the SpMV through the remapped columns reproduces A u on every rank's rows
    u = rng.standard_normal(n)
    full = A @ u
    for r in range(world):
        P = plans[r]
        local = np.full(npad + int(P["ro"][-1]), np.nan)
        local[P["r0"]:P["r1"]] = u[P["r0"]:P["r1"]]                      # only the owned chunk is valid locally
        for q in range(world):
            vals = u[plans[q]["send"][r]]                                   # what q pushes to r ...
            local[npad + P["ro"][q]: npad + P["ro"][q + 1]] = vals          # ... lands in r's slots for q
        lo, hi = Ap[P["r0"]], Ap[P["r1"]]
        rows = np.repeat(np.arange(P["r0"], P["r1"]), np.diff(Ap[P["r0"]:P["r1"] + 1]))
        got = np.zeros(n)
        np.add.at(got, rows, A.data[lo:hi] * local[P["cols"]])
        assert np.allclose(got[P["r0"]:P["r1"]], full[P["r0"]:P["r1"]], rtol=1e-13, atol=1e-13)
    print(f"{kind}, world {world}: {moved} halo nodes per exchange in total against {n * (world - 1)} for an all-gather")
    assert moved < n * (world - 1)
