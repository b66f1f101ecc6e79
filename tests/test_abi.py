"""CPU tests of the drop-in boundary: libfans_gpu.so loads, exports every symbol include/fans_gpu.h declares, its
struct layouts match the ctypes mirror, and the compute entry points fail LOUDLY (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fans_gpu.h")


@pytest.fixture(scope="module")
def lib():
    from fans_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fans_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from fans_b200 import _lib
    names = declared_functions()
    assert len(names) >= 29
    for n in names:
        assert hasattr(lib, n), "libfans_gpu.so does not export " + n
    assert sorted(_lib.EXPORTS) == names, (set(names) ^ set(_lib.EXPORTS))


def test_struct_layouts_match_header(lib, tmp_path):
    from fans_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fans_gpu.h"\nint main(void){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(fans_phase_desc), sizeof(fans_config), sizeof(fans_mixed_bc),'
                    'sizeof(fans_solve_params), sizeof(fans_solve_result), offsetof(fans_config, nccl_comm), offsetof(fans_mixed_bc, M),'
                    'offsetof(fans_solve_params, ls_tol)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)  # header is plain C
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.PhaseDesc), C.sizeof(_lib.Config), C.sizeof(_lib.MixedBCDesc), C.sizeof(_lib.SolveParams), C.sizeof(_lib.SolveResult),
            _lib.Config.nccl_comm.offset, _lib.MixedBCDesc.M.offset, _lib.SolveParams.ls_tol.offset]
    assert got == want


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_fallback(lib):
    from fans_b200 import _lib
    cfg = _lib.Config()
    cfg.dims[:] = [8, 8, 8]
    cfg.L[:] = [1.0, 1.0, 1.0]
    cfg.howmany, cfg.n_str, cfg.fe_type, cfg.world_size = 1, 3, 0, 1
    cfg.local_n0, cfg.local_n1, cfg.device = 8, 8, -1
    ptr = C.c_void_p()
    rc = lib.fans_create(C.byref(ptr), C.byref(cfg))
    assert rc == 2 and not ptr.value  # FANS_ERR_CUDA
    assert b"no CPU fallback" in lib.fans_last_error(None)
    with pytest.raises(_lib.FansError):
        _lib.Context((8, 8, 8), [1, 1, 1], 1, 3)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may touch oracle/."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "fans_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"fans_oracle|oracle/|import oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_fibre_image_generator():
    """fans_b200.simple.fiber_microstructure (BASELINE config 3 image): deterministic, z-invariant, 64 periodic non-overlapping discs
    at the requested volume fraction."""
    import numpy as np
    from fans_b200 import simple
    ms = simple.fiber_microstructure(64)
    assert ms.dtype == np.uint16 and ms.shape == (64, 64, 64) and set(np.unique(ms)) == {0, 1}
    assert (ms == ms[:, :, :1]).all()                       # fibres run along z, the reference's fastest axis
    assert abs(float(ms.mean()) - 0.4) < 0.03               # discs are rasterised on 64^2 pixels
    assert (simple.fiber_microstructure(64) == ms).all()    # seeded
    # separate discs: many periodic connected components of the fibre phase, never more than the number of fibres
    plane = ms[:, :, 0].astype(bool)
    lab = -np.ones(plane.shape, dtype=int)
    n = 0
    for i, j in zip(*np.nonzero(plane)):
        if lab[i, j] >= 0:
            continue
        stack = [(i, j)]
        lab[i, j] = n
        while stack:
            a, b = stack.pop()
            for da, db in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                c, d = (a + da) % 64, (b + db) % 64
                if plane[c, d] and lab[c, d] < 0:
                    lab[c, d] = n
                    stack.append((c, d))
        n += 1
    assert 20 <= n <= 64    # discs at distance ~2r touch on the raster and merge; a broken generator gives 1 blob or > 64 specks


def test_voronoi_image_generator():
    """fans_b200.simple.voronoi_microstructure (BASELINE config 4 image): deterministic, two phases, slab-wise generation equals the
    full image (what every rank of the multi-GPU bench relies on), periodic in every direction."""
    import numpy as np
    from fans_b200 import simple
    dims = (32, 16, 24)
    ms = simple.voronoi_microstructure(dims, n_seeds=12)
    assert ms.dtype == np.uint16 and ms.shape == dims and set(np.unique(ms)) == {0, 1}
    assert 0.2 < float(ms.mean()) < 0.8
    lo, hi = simple.voronoi_microstructure(dims, n_seeds=12, x0=0, n0=16), simple.voronoi_microstructure(dims, n_seeds=12, x0=16, n0=16)
    assert np.array_equal(np.concatenate([lo, hi]), ms)
    # periodic: the image of the seeds shifted by one period is the same image
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(2024)
    pts = rng.uniform(0.0, 1.0, size=(12, 3)) * np.array(dims, dtype=np.float64)
    img = np.concatenate([pts + np.array(s) * np.array(dims) for s in np.ndindex(3, 3, 3)]) - np.array(dims)   # 27 periodic copies
    x, y, z = np.meshgrid(*(np.arange(n) + 0.5 for n in dims), indexing="ij")
    _, lab = cKDTree(img).query(np.stack([x.ravel(), y.ravel(), z.ravel()], 1))
    assert np.array_equal(((lab % 12) % 2).reshape(dims).astype(np.uint16), ms)
