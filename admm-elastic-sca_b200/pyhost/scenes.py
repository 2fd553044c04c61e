"""Synthetic workload definitions (scene dictionaries) shared by tests, bench.py and the
golden generator.  A scene dictionary is the neutral description both the reference
shim (oracle/ref.py) and the CUDA path (admm_b200.py) are built from:

  name, dt, iters            solver settings        (System::Settings, A/src/system/System.hpp:35-42)
  x  [n,3] f64               rest positions (already rounded through float32 like the reference's
                             mesh loader does, M/deps/trimesh2/include/Vec.h:452-453)
  m  [n]   f64               node masses (the reference stores them x3, System.hpp:49)
  batches                    list of force batches in the order the reference would push them on
                             System::forces (System.hpp:52)
  explicit                   gravity / wind (ExplicitForce.hpp:51-68)
  x_after_init  [n,3]        optional positions written into m_x after initialize(), the way
                             singletet.cpp:40 and bunnyexpand.cpp:59-63 excite the system

Tet kinds: 0 LinearTetStrain (ARAP), 1 NeoHookean, 2 StVK, 3 TetVolume.
Tri kinds: 0 LimitedTriangleStrain, 1 TriArea, 2 FungTriangle.
Collision shape kinds: 0 sphere, 1 cylinder (axis || z), 2 floor (y plane).
"""
import numpy as np

TET_ARAP, TET_NH, TET_STVK, TET_VOLUME = 0, 1, 2, 3
TRI_STRAIN, TRI_AREA, TRI_FUNG = 0, 1, 2
SHAPE_SPHERE, SHAPE_CYLINDER, SHAPE_FLOOR = 0, 1, 2


def _f32round(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def tet_volumes(x, tets):
    v0, v1, v2, v3 = (x[tets[:, i]] for i in range(4))
    return np.abs(np.einsum("ij,ij->i", v0 - v3, np.cross(v1 - v3, v2 - v3))) / 6.0


def density_weighted_tet_mass(x, tets, total_mass):
    """src/ForceBuilder.hpp:196-231: density = mass / total volume, a quarter of each tet to its corners."""
    vol = tet_volumes(x, tets)
    density = total_mass / vol.sum()
    m = np.zeros(x.shape[0])
    np.add.at(m, tets.reshape(-1), np.repeat(density * vol / 4.0, 4))
    return m


def density_weighted_tri_mass(x, tris, total_mass):
    """src/ForceBuilder.hpp:238-283."""
    a = 0.5 * np.linalg.norm(np.cross(x[tris[:, 1]] - x[tris[:, 0]], x[tris[:, 2]] - x[tris[:, 0]]), axis=1)
    density = total_mass / a.sum()
    m = np.zeros(x.shape[0])
    np.add.at(m, tris.reshape(-1), np.repeat(density * a / 3.0, 3))
    return m


def kuhn_cube(N):
    """Kuhn 6-tet subdivision of an N^3 grid on [-1/2,1/2]^3 (SURVEY 8d input 5)."""
    g = np.linspace(-0.5, 0.5, N + 1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    x = _f32round(np.stack([X, Y, Z], axis=-1).reshape(-1, 3))
    S = N + 1

    def vid(i, j, k):
        return (i * S + j) * S + k

    I, J, K = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    I, J, K = I.reshape(-1), J.reshape(-1), K.reshape(-1)
    e = np.eye(3, dtype=np.int64)
    tets = []
    for perm in [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]:
        o1 = e[perm[0]]
        o2 = o1 + e[perm[1]]
        v0 = vid(I, J, K)
        v1 = vid(I + o1[0], J + o1[1], K + o1[2])
        v2 = vid(I + o2[0], J + o2[1], K + o2[2])
        v3 = vid(I + 1, J + 1, K + 1)
        tets.append(np.stack([v0, v1, v2, v3], axis=1))
    tets = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)
    return x, tets


def cube_scene(N, kind=TET_NH, mu=1e5, lam=1e5, maxit=5, stiffness=1e5, mass=1000.0, dt=0.04, iters=10,
               stretch=1.3, gravity=True, seed=None, name=None):
    """Synthetic tetrahedralised cube (BASELINE.json configs[4], SURVEY 8d input 5)."""
    x, tets = kuhn_cube(N)
    m = density_weighted_tet_mass(x, tets, mass)
    b = dict(type="tets", kind=kind, idx=tets)
    if kind in (TET_NH, TET_STVK):
        b.update(p0=mu, p1=lam, maxit=maxit)
    elif kind == TET_ARAP:
        b.update(p0=stiffness)
    else:
        b.update(p0=stiffness, p1=0.95, p2=1.05)
    sc = dict(name=name or f"cube{N}", dt=dt, iters=iters, x=x, m=m, batches=[b], explicit=[])
    if gravity:
        sc["explicit"].append(dict(type="gravity", dir=np.array([0.0, -9.8, 0.0])))
    xa = x * stretch if stretch is not None else None
    if seed is not None:
        rng = np.random.default_rng(seed)
        xa = (xa if xa is not None else x) + rng.uniform(-0.3, 0.3, size=x.shape) / N
    if xa is not None:
        sc["x_after_init"] = xa
    return sc


def grid_cloth(W, H, sx=1.5, sy=1.0):
    """(W+1)x(H+1) vertex sheet in the xy plane, two triangles per quad, plus interior hinges.

    Hinge vertex order follows the reference's BendForce convention (idx0, idx1 = the two
    wing vertices, idx2, idx3 = the shared edge; BendForce.cpp:26-55, ForceBuilder.cpp:140-218).
    """
    gx = np.linspace(-sx / 2, sx / 2, W + 1)
    gy = np.linspace(-sy / 2, sy / 2, H + 1)
    X, Y = np.meshgrid(gx, gy, indexing="ij")
    x = _f32round(np.stack([X, Y, np.zeros_like(X)], axis=-1).reshape(-1, 3))

    def vid(i, j):
        return i * (H + 1) + j

    tris = []
    for i in range(W):
        for j in range(H):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
            tris.append((a, b, c))
            tris.append((a, c, d))
    tris = np.array(tris, dtype=np.int32)
    # hinges: for every interior edge shared by two triangles
    edge_map = {}
    for t, tri in enumerate(tris):
        for k in range(3):
            e = tuple(sorted((int(tri[k]), int(tri[(k + 1) % 3]))))
            edge_map.setdefault(e, []).append((t, int(tri[(k + 2) % 3])))
    hinges, springs = [], []
    for e, lst in sorted(edge_map.items()):
        springs.append(e)
        if len(lst) == 2:
            hinges.append((lst[0][1], lst[1][1], e[0], e[1]))
    return x, tris, np.array(hinges, dtype=np.int32), np.array(springs, dtype=np.int32)


def cloth_scene(W=12, H=8, stiffness=100.0, bend=20.0, limit=(0.95, 1.05), mass=0.5, dt=0.04, iters=30,
                wind=None, springs=False, anchors=(0, None), seed=3, name=None):
    """windyflag-shaped workload: LimitedTriangleStrain + BendForce (+ Spring), two StaticAnchors, gravity, wind."""
    x, tris, hinges, edges = grid_cloth(W, H)
    m = density_weighted_tri_mass(x, tris, mass)
    limit = _f32round(limit)  # XML limits pass through float (ForceBuilder.cpp:109-110)
    batches = [dict(type="tris", kind=TRI_STRAIN, idx=tris, stiffness=stiffness, lmin=limit[0], lmax=limit[1], flag=1),
               dict(type="bends", idx=hinges, stiffness=bend)]
    if springs:
        batches.append(dict(type="springs", idx=edges, stiffness=stiffness))
    a = [anchors[0], H if anchors[1] is None else anchors[1]]
    batches.append(dict(type="static_anchors", idx=np.array(a, dtype=np.int32), weight=-1.0))
    sc = dict(name=name or f"cloth{W}x{H}", dt=dt, iters=iters, x=x, m=m, batches=batches,
              explicit=[dict(type="gravity", dir=np.array([0.0, -9.8, 0.0]))])
    if wind is not None:
        sc["explicit"].append(dict(type="wind", tris=tris, dir=np.asarray(wind, dtype=np.float64)))
    if seed is not None:
        rng = np.random.default_rng(seed)
        sc["x_after_init"] = x + rng.uniform(-0.01, 0.01, size=x.shape)
    return sc


def singletet_scene():
    """A/samples/singletet.cpp:27-111: 3 StaticAnchors + 1 LinearTetStrain(k=1), dt=1, 20 its, node 3 x := 200."""
    x = np.zeros((4, 3))
    x[0, 1] = 1
    x[2, 2] = 1
    x[3, 0] = 1
    xa = x.copy()
    xa[3, 0] = 200.0
    return dict(name="singletet", dt=1.0, iters=20, x=x, m=np.ones(4),
                batches=[dict(type="static_anchors", idx=np.array([0, 1, 2], dtype=np.int32), weight=-1.0),
                         dict(type="tets", kind=TET_ARAP, idx=np.array([[0, 1, 2, 3]], dtype=np.int32), p0=1.0)],
                explicit=[], x_after_init=xa)


def singlenode_scene():
    """A/samples/singlenode.cpp:25-73: one node, gravity only, dt=1, 20 its."""
    return dict(name="singlenode", dt=1.0, iters=20, x=np.zeros((1, 3)), m=np.ones(1), batches=[],
                explicit=[dict(type="gravity", dir=_f32round([0.0, -9.8, 0.0]))])


def save_scene(path, sc):
    """Flatten a scene dictionary into an .npz (golden input fixture)."""
    import json
    arrays, meta = {}, dict(name=sc["name"], dt=sc["dt"], iters=sc["iters"], batches=[], explicit=[])
    arrays["x"] = sc["x"]
    arrays["m"] = sc["m"]
    if "x_after_init" in sc:
        arrays["x_after_init"] = sc["x_after_init"]
    for i, b in enumerate(sc["batches"]):
        mb = {}
        for k, v in b.items():
            if isinstance(v, np.ndarray):
                arrays[f"b{i}_{k}"] = v
                mb[k] = f"@b{i}_{k}"
            else:
                mb[k] = v.item() if isinstance(v, np.generic) else v
        meta["batches"].append(mb)
    for i, e in enumerate(sc.get("explicit", [])):
        me = {}
        for k, v in e.items():
            if isinstance(v, np.ndarray):
                arrays[f"e{i}_{k}"] = v
                me[k] = f"@e{i}_{k}"
            else:
                me[k] = v
        meta["explicit"].append(me)
    for k in sc.get("extra", {}):
        arrays[f"x_{k}"] = np.asarray(sc["extra"][k])
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


def load_scene(path):
    import json
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    sc = dict(name=meta["name"], dt=meta["dt"], iters=meta["iters"], x=z["x"], m=z["m"], batches=[], explicit=[])
    if "x_after_init" in z:
        sc["x_after_init"] = z["x_after_init"]

    def res(d):
        return {k: (z[v[1:]] if isinstance(v, str) and v.startswith("@") else v) for k, v in d.items()}

    sc["batches"] = [res(b) for b in meta["batches"]]
    sc["explicit"] = [res(e) for e in meta["explicit"]]
    sc["extra"] = {k[2:]: z[k] for k in z.files if k.startswith("x_") and k != "x_after_init"}
    return sc


# ---- trajectory hand-off: TetGen .node / .ele as the reference writes them -------------------------------------
def save_tetmesh(filename, x, tets=None):
    """TetMesh::save (M/src/TetMesh.cpp:306-352): `<filename>.node` (and `<filename>.ele` when tets are given) in
    TetGen format.  Coordinates go through float and are printed at the default stream precision, as trimesh's
    Vec::str does (Vec.h:229-237) -- i.e. '%g'."""
    xf = np.asarray(x, dtype=np.float32).reshape(-1, 3)
    with open(filename + ".node", "w") as f:
        f.write(f"{xf.shape[0]} 3 0 0\n")
        for i, p in enumerate(xf):
            f.write("\t%d %g %g %g\n" % (i, p[0], p[1], p[2]))
        f.write("# Generated by mclscene (www.mattoverby.net)")
    if tets is not None:
        t = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
        with open(filename + ".ele", "w") as f:
            f.write(f"{t.shape[0]} 4 0\n")
            for i, e in enumerate(t):
                f.write("\t%d %d %d %d %d\n" % (i, e[0], e[1], e[2], e[3]))
            f.write("# Generated by mclscene (www.mattoverby.net)")


def load_tetmesh(filename):
    """Reads `<filename>.node` (and `.ele` if present) back: (x float32 [n, 3], tets int32 [T, 4] or None)."""
    import os

    def rows(path):
        with open(path) as f:
            return [ln.split() for ln in f if ln.strip() and not ln.lstrip().startswith("#")]
    r = rows(filename + ".node")
    n = int(r[0][0])
    x = np.array([[float(v) for v in q[1:4]] for q in r[1:1 + n]], dtype=np.float32)
    tets = None
    if os.path.exists(filename + ".ele"):
        r = rows(filename + ".ele")
        tets = np.array([[int(v) for v in q[1:5]] for q in r[1:1 + int(r[0][0])]], dtype=np.int32)
    return x, tets
