"""Synthetic meshes for the benchmark configurations: geodesic icosphere of
frequency nu (20 nu^2 panels, 30 nu^2 edges; nu = 2^L reproduces recursive 4-way subdivision)
written as GMSH v2.2 ASCII + a .scuffgeo wrapper.  Deterministic (no RNG).  Every line stays
below 100 characters because the reference parser reads with fgets(Line,100,...)
(libs/libscuff/ParseMeshFiles.cc:77-81)."""
import os
import numpy as np


def icosphere(nu, radius=1.0):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    V0 = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                   [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    F0 = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
          (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
          (8, 6, 7), (9, 8, 1)]
    verts = {}
    out = []

    def vid(key, p):
        if key not in verts:
            verts[key] = len(out)
            out.append(p)
        return verts[key]

    tris = []
    for (a, b, c) in F0:
        idx = {}
        for i in range(nu + 1):
            for j in range(nu + 1 - i):
                k = nu - i - j
                # canonical key: sorted (corner, weight) pairs with non-zero weight -> shared along edges/corners
                key = tuple(sorted((cc, w) for cc, w in ((a, k), (b, i), (c, j)) if w > 0))
                p = (k * V0[a] + i * V0[b] + j * V0[c]) / nu
                idx[(i, j)] = vid(key, p)
        for i in range(nu):
            for j in range(nu - i):
                tris.append((idx[(i, j)], idx[(i + 1, j)], idx[(i, j + 1)]))
                if i + j < nu - 1:
                    tris.append((idx[(i + 1, j)], idx[(i + 1, j + 1)], idx[(i, j + 1)]))
    V = np.array(out)
    V *= radius / np.linalg.norm(V, axis=1)[:, None]
    return V, np.array(tris, dtype=np.int64)


def write_msh(path, V, T):
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(V))
        for i, v in enumerate(V):
            f.write("%d %.17g %.17g %.17g\n" % (i + 1, v[0], v[1], v[2]))
        f.write("$EndNodes\n$Elements\n%d\n" % len(T))
        for i, t in enumerate(T):
            f.write("%d 2 2 1 1 %d %d %d\n" % (i + 1, t[0] + 1, t[1] + 1, t[2] + 1))
        f.write("$EndElements\n")


def write_sphere_geo(outdir, nu, material=None, name=None):
    """returns the .scuffgeo path of a unit icosphere with 30 nu^2 RWG edges (PEC unless material given)"""
    os.makedirs(outdir, exist_ok=True)
    V, T = icosphere(nu)
    name = name or ("IcoSphere_%d%s" % (30 * nu * nu, "" if material is None else "_" + material))
    msh = os.path.join(outdir, "IcoSphere_%d.msh" % (30 * nu * nu))
    if not os.path.exists(msh):
        write_msh(msh, V, T)
    geo = os.path.join(outdir, name + ".scuffgeo")
    with open(geo, "w") as f:
        f.write("OBJECT TheSphere\n\tMESHFILE %s\n" % os.path.basename(msh))
        if material:
            f.write("\tMATERIAL %s\n" % material)
        f.write("ENDOBJECT\n")
    return geo
