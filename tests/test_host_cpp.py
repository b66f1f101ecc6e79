"""CPU tests of the C++ host layer (fans_b200/host/*.hpp): the mirror of Reader / Matmodel / MaterialManager / MixedBC that sits
above the C ABI.  `FANS_gpu --describe` prints what the host derives from a reference input file (phase descriptors, reference
stiffness, mixed-BC matrices) without touching the GPU; it is compared with the oracle's parse of the same input."""
import json
import os
import subprocess

import numpy as np
import pytest

import fans_oracle as fo
import golden_util as gu
import cpp_host
import util

SCENARIOS = ["LinearThermal", "LinearElastic", "PseudoPlastic", "J2Plasticity", "CompressibleNeoHookean", "MixedBCs", "MixedBCs_LargeStrain"]


@pytest.fixture(scope="module")
def exe():
    return cpp_host.build()


@pytest.fixture(scope="module")
def ms_file(tmp_path_factory):
    p = tmp_path_factory.mktemp("ms") / "sphere32.u16"
    gu.sphere32().tofile(p)
    return str(p)


@pytest.mark.parametrize("name", SCENARIOS)
def test_describe_matches_oracle(exe, ms_file, name, tmp_path):
    cfg = gu.reference_input(name)
    inp = tmp_path / "in.json"
    inp.write_text(json.dumps(cfg))
    out = subprocess.run([exe, "--describe", str(inp), ms_file, "32", "32", "32"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout)
    strain_type = cfg.get("strain_type", "small")
    par = fo.OracleSolver(gu.sphere32(), cfg["microstructure"]["L"], cfg["problem_type"], cfg["materials"], cfg.get("FE_type", "HEX8"),
                          cfg["method"], strain_type, cfg["error_parameters"], 0)
    assert (d["howmany"], d["n_str"], d["n_phases"], d["all_linear"]) == (par.h, par.n_str, par.n_phases, par.all_linear)
    assert np.allclose(np.array(d["kapparef"]).reshape(par.n_str, par.n_str), par.kapparef, rtol=1e-13, atol=1e-13)
    assert np.allclose(d["volume_fractions"], [1 - 8744 / 32768, 8744 / 32768])
    for got, want in zip(d["phases"], util.phase_descs_from_oracle(par)):
        assert (got["model"], got["local_mat"], got["group_n_mat"]) == (want.model, want.local_mat, want.group_n_mat)
        assert np.allclose(got["params"], list(want.params), rtol=1e-13, atol=1e-15)
    mixed = [lc for lc in cfg["macroscale_loading"] if isinstance(lc, dict)]
    assert len(d["mixed_M"]) == len(mixed)
    for got, lc in zip(d["mixed_M"], mixed):
        mbc = fo.MixedBC(lc["strain_indices"], lc["stress_indices"], lc.get("strain", []), lc.get("stress", []), par.n_str)
        mbc.finalize(par.kapparef)
        assert np.allclose(np.array(got).reshape(mbc.M.shape), mbc.M, rtol=1e-10, atol=1e-14)


def test_bad_inputs_raise_reference_errors(exe, ms_file, tmp_path):
    cfg = gu.reference_input("LinearElastic")
    for mutate, msg in [(lambda c: c.update(FE_type="HEX27"), "FE_type must be one of"),
                        (lambda c: c.update(problem_type="acoustic"), "not a valid problem type"),
                        (lambda c: c["materials"][0].update(matmodel="Nope"), "Nope"),
                        (lambda c: c["materials"][0].update(phases=[0, 2]), "not assigned")]:
        c = json.loads(json.dumps(cfg))
        mutate(c)
        inp = tmp_path / "bad.json"
        inp.write_text(json.dumps(c))
        out = subprocess.run([exe, "--describe", str(inp), ms_file, "32", "32", "32"], capture_output=True, text=True)
        assert out.returncode == 10 and msg in out.stderr, (out.returncode, out.stderr)


@pytest.mark.skipif(not os.path.exists("/root/reference/test/microstructures/sphere32.h5"), reason="reference checkout not present")
def test_h5mini_reads_the_reference_microstructure(exe, tmp_path):
    """The built-in HDF5 reader on the reference's own file (deflate chunk, zyx order) == the committed fixture."""
    cfg = gu.reference_input("LinearThermal")
    inp = tmp_path / "in.json"
    inp.write_text(json.dumps(cfg))
    out = subprocess.run([exe, "--describe", str(inp)], capture_output=True, text=True, cwd="/root/reference/test")
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout)
    assert d["dims"] == [32, 32, 32]
    assert np.allclose(d["volume_fractions"], [1 - 8744 / 32768, 8744 / 32768])


def test_h5_results_writer(exe, tmp_path):
    """fans_b200/host/h5write.hpp: `FANS_gpu --h5selftest` writes an HDF5 results file with the reference's layout (nested groups, 40
    time steps, f64 / f32 / u16 / i32, fields transposed to [Z][Y][X][extra] + permute_order = "zyx", include/reader.h:173-351);
    tests/h5_minireader.py — an independent reader written from the format specification, which also walks the reference's own
    sphere32.h5 — parses it back."""
    import h5_minireader as h5
    out = tmp_path / "selftest.h5"
    r = subprocess.run([exe, "--h5selftest", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    f = h5.H5File(str(out))
    w = f.walk()
    base = "/img/4x3x2/ms_results/run1"
    assert len(w) == 45
    for t in range(40):
        a, attrs = w["%s/load0/time_step%d/stress_average" % (base, t)]
        assert a.dtype == np.float64 and np.array_equal(a, t + 0.125 * np.arange(6)) and attrs == {}
    x, y, z, c = np.meshgrid(np.arange(4), np.arange(3), np.arange(2), np.arange(3), indexing="ij")
    field_xyz = 1000.0 * x + 100.0 * y + 10.0 * z + c + 0.5
    d, attrs = w[base + "/load0/time_step0/displacement"]
    assert attrs == {"permute_order": "zyx"} and d.shape == (2, 3, 4, 3)
    assert np.array_equal(d, np.transpose(field_xyz, (2, 1, 0, 3)))         # [Z][Y][X][extra]
    m, attrs = w[base + "/load0/time_step0/microstructure"]
    assert m.dtype == np.uint16 and m.shape == (2, 3, 4, 1) and attrs == {"permute_order": "zyx"}
    assert np.array_equal(m[..., 0], np.transpose(100 * x[..., 0] + 10 * y[..., 0] + z[..., 0], (2, 1, 0)))
    assert np.array_equal(w[base + "/load1/time_step0/some_floats"][0], np.array([1.5, -2.25, 3.0], dtype=np.float32))
    assert np.array_equal(w[base + "/load1/time_step0/plastic_flag_like"][0], np.array([-7, 0, 123456], dtype=np.int32))
    assert np.array_equal(w[base + "/load1/time_step0/homogenized_tangent"][0], (np.arange(36.0) ** 2).reshape(6, 6))
    # the C++ microstructure reader (h5mini.hpp, validated on the reference's h5py-written fixture) resolves the symbol-table groups of
    # the written file and reads the [Z][Y][X][1] field back as an image (permute_order = "zyx" -> dims X, Y, Z)
    cfg = gu.reference_input("LinearElastic")
    cfg["microstructure"] = {"filepath": str(out), "datasetname": base + "/load0/time_step0/microstructure", "L": [1.0, 1.0, 1.0]}
    inp = tmp_path / "in_h5.json"
    inp.write_text(json.dumps(cfg))
    r = subprocess.run([exe, "--describe", str(inp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout)["dims"] == [4, 3, 2]
    # the same reader walks a file written by the real HDF5 library (h5py): the reference's fixture
    ref = "/root/reference/test/microstructures/sphere32.h5"
    if os.path.exists(ref):
        assert list(h5.H5File(ref).walk()) == ["/sphere/32x32x32/ms"]


def gb_expected():
    normals = {2: np.array([1.0, 0.0, 0.0]), 3: np.array([0.6, 0.8, 0.0])}
    D_bulk, D_par, D_perp = 1.5, 4.0, 0.25
    kap = [D_bulk * np.eye(3), D_bulk * np.eye(3)]
    for t in (2, 3):
        n = normals[t]
        kap.append(D_par * (np.eye(3) - np.outer(n, n)) + D_perp * np.outer(n, n))
    return kap


def gb_cfg(path, props_extra=None):
    props = {"GB_unformity": True, "D_bulk": 1.5, "D_par": 4.0, "D_perp": 0.25}
    props.update(props_extra or {})
    return {"microstructure": {"filepath": path, "datasetname": "/gb/image", "L": [2.0, 1.0, 1.0]}, "problem_type": "thermal",
            "materials": [{"phases": [0, 1, 2, 3], "matmodel": "GBDiffusion", "material_properties": props}], "FE_type": "HEX8",
            "method": "cg", "error_parameters": {"measure": "Linfinity", "type": "absolute", "tolerance": 1e-10}, "n_it": 100,
            "macroscale_loading": [[[0.01, 0.02, -0.01]]], "results": ["stress_average", "GBnormals"]}


def test_gbdiffusion_reads_dataset_attributes(exe, tmp_path):
    """GBDiffusion (GBDiffusion.h:45-135): num_crystals / num_GB / GBVoxelInfo come from the attributes of the microstructure dataset
    (written here by the repo's own HDF5 writer, read back by the independent mini reader and by the C++ front end); the phase
    tangents are D_bulk I in the crystals and D_par (I - N N^T) + D_perp N N^T in the boundary phases, kapparef their mean."""
    import h5_minireader as h5
    f = tmp_path / "gb.h5"
    assert subprocess.run([exe, "--gbselftest", str(f)]).returncode == 0
    img, attrs = h5.H5File(str(f)).walk()["/gb/image"]
    assert img.shape == (4, 4, 8) and attrs["num_crystals"] == 2 and attrs["num_GB"] == 2 and "GB_normal" in attrs["GBVoxelInfo"]
    inp = tmp_path / "in.json"
    inp.write_text(json.dumps(gb_cfg(str(f))))
    out = subprocess.run([exe, "--describe", str(inp)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout)
    kap = gb_expected()
    assert (d["howmany"], d["n_str"], d["n_phases"], d["all_linear"], d["dims"]) == (1, 3, 4, True, [8, 4, 4])
    for got, want in zip(d["phases"], kap):
        assert got["model"] == 0 and np.allclose(np.array(got["params"][:9]).reshape(3, 3), want, rtol=1e-14, atol=1e-15)
    assert np.allclose(np.array(d["kapparef"]).reshape(3, 3), sum(kap) / 4, rtol=1e-14)
    # the same through material_properties for images without HDF5 attributes (.npy / raw)
    ms = np.zeros((8, 4, 4), dtype=np.uint16)
    ms[3], ms[4:7], ms[7] = 2, 1, 3
    raw = tmp_path / "gb.u16"
    ms.tofile(raw)
    info = {"a": {"GB_tag": 2, "GB_normal": [1.0, 0.0, 0.0]}, "b": {"GB_tag": 3, "GB_normal": [0.6, 0.8, 0.0]}}
    inp.write_text(json.dumps(gb_cfg("unused.h5", {"num_crystals": 2, "num_GB": 2, "GBVoxelInfo": info})))
    out = subprocess.run([exe, "--describe", str(inp), str(raw), "8", "4", "4"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    d2 = json.loads(out.stdout)
    assert d2["phases"] == d["phases"] and d2["kapparef"] == d["kapparef"]
    # no attributes, no fallback entries: the reference's error text
    inp.write_text(json.dumps(gb_cfg("unused.h5")))
    out = subprocess.run([exe, "--describe", str(inp), str(raw), "8", "4", "4"], capture_output=True, text=True)
    assert out.returncode != 0 and "Error in GBDiffusion initialization" in out.stderr
