"""Convenience constructors above the C ABI for the headline workload (two-phase linear elasticity / conduction).
Host-side restatement of what the reference's LinearElasticIsotropic / LinearThermalIsotropic constructors and
MaterialManager::compute_reference_stiffness do (include/material_models/LinearElastic.h:7-75,
LinearThermal.h:7-46, MaterialManager.h:177-205) — parameter bookkeeping only, all numerics run in libfans_gpu."""
import os

import numpy as np

from . import _lib as L


def sphere_microstructure(n, radius_frac=0.4):
    """Synthetic config-2/-headline image: phase 1 inside a centred sphere of radius 0.4 n."""
    c = (n - 1) / 2.0
    i = (np.arange(n) - c) ** 2
    ms = np.empty((n, n, n), dtype=np.uint16)
    r2 = (radius_frac * n) ** 2
    for x in range(n):  # plane by plane keeps the temporary small at 512^3
        ms[x] = (i[x] + i[:, None] + i[None, :]) <= r2
    return ms


def ellipsoid_microstructure(dims, x0=0, n0=None):
    """Slab [x0, x0+n0) of the bench image on a dims[0] x dims[1] x dims[2] grid: phase 1 inside the centred ellipsoid with
    semi-axes 0.4 n_x, 0.4 n_y, 0.4 n_z (the sphere of sphere_microstructure when the grid is a cube)."""
    nx, ny, nz = dims
    n0 = nx if n0 is None else n0
    fx = ((np.arange(nx) - (nx - 1) / 2.0) / (0.4 * nx)) ** 2
    fy = ((np.arange(ny) - (ny - 1) / 2.0) / (0.4 * ny)) ** 2
    fz = ((np.arange(nz) - (nz - 1) / 2.0) / (0.4 * nz)) ** 2
    ms = np.empty((n0, ny, nz), dtype=np.uint16)
    for i in range(n0):
        ms[i] = (fx[x0 + i] + fy[:, None] + fz[None, :]) <= 1.0
    return ms


def elastic_tangent(lam, mu):
    k = np.zeros((6, 6))
    k[:3, :3] = lam
    k += 2.0 * mu * np.eye(6)
    return k


def elastic_phase_descs(bulk, shear):
    """phase descriptors of LinearElasticIsotropic phases 0..len(bulk)-1 (LinearElastic.h:43-53: lambda = K - 2 mu / 3)"""
    bulk = np.asarray(bulk, dtype=np.float64)
    mu = np.asarray(shear, dtype=np.float64)
    lam = bulk - (2.0 / 3.0) * mu
    descs = []
    for i in range(len(bulk)):
        d = L.PhaseDesc()
        d.model, d.local_mat, d.group_n_mat = L.MAT_LINEAR, i, len(bulk)
        for k, v in enumerate(elastic_tangent(lam[i], mu[i]).reshape(-1)):
            d.params[k] = v
        descs.append(d)
    return descs


def linear_elastic_context(ms, Lbox, bulk, shear, fe_type="HEX8", device=-1, gdims=None, comm=None):
    """LinearElasticIsotropic on phases 0..len(bulk)-1; reference stiffness = (max+min)/2 of lambda and mu.
    ms is this rank's slab; gdims the global grid (defaults to ms.shape)."""
    bulk = np.asarray(bulk, dtype=np.float64)
    mu = np.asarray(shear, dtype=np.float64)
    lam = bulk - (2.0 / 3.0) * mu
    ctx = L.Context(gdims if gdims is not None else ms.shape, Lbox, 3, 6, fe_type, device, comm)
    ctx.set_materials(elastic_phase_descs(bulk, mu))
    ctx.set_microstructure(ms)
    ctx.set_reference_stiffness(elastic_tangent((lam.max() + lam.min()) / 2, (mu.max() + mu.min()) / 2))
    return ctx


def linear_elastic_tensor_context(ms, Lbox, tangents, fe_type="HEX8", device=-1, gdims=None, comm=None):
    """LinearElasticTriclinic-style phases: one full 6x6 Mandel tangent per phase id (LinearElastic.h:77-159); reference stiffness =
    their mean (LinearElastic.h:146-150).  A polycrystal with one tensor per grain."""
    tangents = np.asarray(tangents, dtype=np.float64)
    ctx = L.Context(gdims if gdims is not None else ms.shape, Lbox, 3, 6, fe_type, device, comm)
    descs = []
    for i in range(len(tangents)):
        d = L.PhaseDesc()
        d.model, d.local_mat, d.group_n_mat = L.MAT_LINEAR, i, len(tangents)
        for k, v in enumerate(tangents[i].reshape(-1)):
            d.params[k] = v
        descs.append(d)
    ctx.set_materials(descs)
    ctx.set_microstructure(ms)
    ctx.set_reference_stiffness(tangents.mean(0))
    return ctx


def rotated_cubic_tangents(n, c11=168.0, c12=121.0, c44=75.0, seed=7):
    """n randomly rotated cubic stiffness tensors (copper-like constants) in Mandel notation: a synthetic polycrystal"""
    rng = np.random.default_rng(seed)
    C = np.zeros((3, 3, 3, 3))
    for i in range(3):
        for j in range(3):
            C[i, i, j, j] += c12
            C[i, j, i, j] += c44
            C[i, j, j, i] += c44
        C[i, i, i, i] += c11 - c12 - 2.0 * c44
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
    out = []
    for _ in range(n):
        Q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        Cr = np.einsum("ia,jb,kc,ld,abcd->ijkl", Q, Q, Q, Q, C)
        M = np.zeros((6, 6))
        for a, (i, j) in enumerate(pairs):
            for b, (k, l) in enumerate(pairs):
                M[a, b] = Cr[i, j, k, l] * (np.sqrt(2.0) if a >= 3 else 1.0) * (np.sqrt(2.0) if b >= 3 else 1.0)
        out.append(0.5 * (M + M.T))
    return np.array(out)


def voronoi_labels(dims, n_seeds, seed=2024):
    """periodic Voronoi grain labels 0..n_seeds-1 (the generator of voronoi_microstructure without the mod 2)"""
    from scipy.spatial import cKDTree
    nx, ny, nz = dims
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.0, 1.0, size=(n_seeds, 3)) * np.array([nx, ny, nz], dtype=np.float64)
    tree = cKDTree(pts, boxsize=[nx, ny, nz])
    yy, zz = np.meshgrid(np.arange(ny) + 0.5, np.arange(nz) + 0.5, indexing="ij")
    q = np.empty((ny * nz, 3))
    q[:, 1], q[:, 2] = yy.ravel(), zz.ravel()
    ms = np.empty((nx, ny, nz), dtype=np.uint16)
    for i in range(nx):
        q[:, 0] = i + 0.5
        _, lab = tree.query(q, workers=-1)
        ms[i] = lab.reshape(ny, nz)
    return ms


def linear_thermal_context(ms, Lbox, conductivity, fe_type="HEX8", device=-1):
    k = np.asarray(conductivity, dtype=np.float64)
    ctx = L.Context(ms.shape, Lbox, 1, 3, fe_type, device)
    descs = []
    for i in range(len(k)):
        d = L.PhaseDesc()
        d.model, d.local_mat, d.group_n_mat = L.MAT_LINEAR, i, len(k)
        for q, v in enumerate((k[i] * np.eye(3)).reshape(-1)):
            d.params[q] = v
        descs.append(d)
    ctx.set_materials(descs)
    ctx.set_microstructure(ms)
    ctx.set_reference_stiffness(np.eye(3) * k.mean())
    return ctx


def fiber_microstructure(n, n_fibers=64, vf=0.4, seed=1234):
    """Synthetic config-3 image: unidirectional fibres along z — n_fibers non-overlapping periodic discs in the x-y
    plane, radius n*sqrt(vf/(n_fibers*pi)), centres by random sequential addition; phase 1 = fibre, phase 0 = matrix."""
    rng = np.random.default_rng(seed)
    r = n * np.sqrt(vf / (n_fibers * np.pi))
    centres = []
    while len(centres) < n_fibers:
        c = rng.uniform(0.0, n, size=2)
        ok = True
        for q in centres:
            d = np.abs(c - q)
            d = np.minimum(d, n - d)
            if d[0] ** 2 + d[1] ** 2 < (2.0 * r) ** 2:
                ok = False
                break
        if ok:
            centres.append(c)
    x = np.arange(n) + 0.5
    plane = np.zeros((n, n), dtype=np.uint16)
    for q in centres:
        dx = np.abs(x - q[0])
        dx = np.minimum(dx, n - dx)
        dy = np.abs(x - q[1])
        dy = np.minimum(dy, n - dy)
        plane |= ((dx[:, None] ** 2 + dy[None, :] ** 2) <= r * r).astype(np.uint16)
    return np.ascontiguousarray(np.broadcast_to(plane[:, :, None], (n, n, n)))


def j2_fiber_context(ms, Lbox, fe_type="HEX8", device=-1, gdims=None, comm=None):
    """Config 3: phase 0 = J2ViscoPlastic_NonLinearIsotropicHardening with the parameters of test/input_files/test_J2Plasticity.json,
    phase 1 = LinearElasticIsotropic fibres (K = 222.222, G = 166.6667).  Reference stiffness = mean over the two material groups of
    their elastic tangents (MaterialManager.h:177-205)."""
    K0, G0, K1, G1 = 62.5, 28.8462, 222.222, 166.6667
    ctx = L.Context(gdims if gdims is not None else ms.shape, Lbox, 3, 6, fe_type, device, comm)
    d0 = L.PhaseDesc()
    d0.model, d0.local_mat, d0.group_n_mat = L.MAT_J2_NONLIN, 0, 1
    for k, v in enumerate([K0, G0, 0.1, 0.0, 0.0, 1.0, 0.01, 0.15, 1000.0]):   # K, G, sigma_y, K_iso, H, eta, dt, sigma_inf, delta
        d0.params[k] = v
    d1 = L.PhaseDesc()
    d1.model, d1.local_mat, d1.group_n_mat = L.MAT_LINEAR, 0, 1
    for k, v in enumerate(elastic_tangent(K1 - 2.0 / 3.0 * G1, G1).reshape(-1)):
        d1.params[k] = v
    ctx.set_materials([d0, d1])
    ctx.set_microstructure(ms)
    ctx.set_reference_stiffness(0.5 * (elastic_tangent(K0 - 2.0 / 3.0 * G0, G0) + elastic_tangent(K1 - 2.0 / 3.0 * G1, G1)))
    return ctx


def spatial_tangent_at_identity(lam, mu):
    """Reference medium of the finite-strain models in closed form: dP/dF of LargeStrainMechModel.h:105-180 at F = I, S = 0 for the
    isotropic tangent (lam, mu).  The reference sums dE_PQ/dF_kL over P <= Q only, so shear pairs carry mu / 2 (kept on purpose: it
    fixes the Green operator and the iteration counts).  Same formula as fans_b200/host/matmodel.hpp."""
    A = np.zeros((9, 9))
    for i in range(3):
        for k in range(3):
            A[3 * i + i, 3 * k + k] = lam + (2.0 * mu if i == k else 0.0)
            if i != k:
                A[3 * i + k, 3 * i + k] = A[3 * i + k, 3 * k + i] = 0.5 * mu
    return A


def neohooke_context(ms, Lbox, bulk, shear, fe_type="HEX8", device=-1, gdims=None, comm=None, model=None):
    """Config 5: CompressibleNeoHookean (or SaintVenantKirchhoff with model=L.MAT_SVK) on phases 0..len(bulk)-1, one material
    group; reference stiffness = mean over the group's materials of the spatial tangent at F = I (LargeStrainMechModel.h:66-86)."""
    bulk = np.asarray(bulk, dtype=np.float64)
    mu = np.asarray(shear, dtype=np.float64)
    lam = bulk - (2.0 / 3.0) * mu
    ctx = L.Context(gdims if gdims is not None else ms.shape, Lbox, 3, 9, fe_type, device, comm)
    descs = []
    for i in range(len(bulk)):
        d = L.PhaseDesc()
        d.model, d.local_mat, d.group_n_mat = (L.MAT_NEOHOOKE if model is None else model), i, len(bulk)
        d.params[0], d.params[1] = lam[i], mu[i]
        descs.append(d)
    ctx.set_materials(descs)
    ctx.set_microstructure(ms)
    ctx.set_reference_stiffness(np.mean([spatial_tangent_at_identity(lam[i], mu[i]) for i in range(len(bulk))], axis=0))
    return ctx


def voronoi_microstructure(dims, n_seeds=None, seed=2024, x0=0, n0=None):
    """Synthetic config-4 image: "polycrystal-like" periodic Voronoi tessellation, phase = grain label mod 2.
    n_seeds defaults to 512 scaled with the volume relative to 1024^3 (grain size ~128 voxels).  Returns the slab [x0, x0 + n0)."""
    from scipy.spatial import cKDTree
    nx, ny, nz = dims
    n0 = nx if n0 is None else n0
    if n_seeds is None:
        n_seeds = max(8, int(round(512.0 * nx * ny * nz / 1024.0 ** 3)))
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.0, 1.0, size=(n_seeds, 3)) * np.array([nx, ny, nz], dtype=np.float64)
    tree = cKDTree(pts, boxsize=[nx, ny, nz])   # periodic nearest neighbour
    # Voronoi cells are convex: when two voxels of a z line lie in the same cell, so does everything between them.  Query every
    # STRIDE-th voxel and only fill in the segments whose ends disagree (a few per cent at ~128-voxel grains): the same image as the
    # voxel-by-voxel query at about a sixth of its cost (a 1024^3 image is 10^9 nearest-neighbour queries otherwise).
    stride = 8 if nz % 8 == 0 and nz >= 64 else 1
    workers = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
    zc = np.arange(0, nz, stride)
    yy, zz = np.meshgrid(np.arange(ny) + 0.5, zc + 0.5, indexing="ij")
    q = np.empty((ny * len(zc), 3))
    q[:, 1], q[:, 2] = yy.ravel(), zz.ravel()
    ms = np.empty((n0, ny, nz), dtype=np.uint16)
    for i in range(n0):
        q[:, 0] = x0 + i + 0.5
        _, lab = tree.query(q, workers=workers)
        lab = lab.reshape(ny, len(zc))
        if stride == 1:
            ms[i] = lab % 2
            continue
        full = np.repeat(lab, stride, axis=1)                       # segment [z, z + stride) takes the label of its left end ...
        differ = lab != np.roll(lab, -1, axis=1)                    # ... unless the next sample (periodic) sits in another cell
        iy, iz = np.nonzero(differ)
        if len(iy):
            off = np.arange(1, stride)
            qy = np.repeat(iy, stride - 1)
            qz = np.repeat(zc[iz], stride - 1) + np.tile(off, len(iy))
            qq = np.stack([np.full(len(qy), x0 + i + 0.5), qy + 0.5, qz + 0.5], axis=1)
            _, l2 = tree.query(qq, workers=workers)
            full[qy, qz] = l2
        ms[i] = full % 2
    return ms
