"""The C++ gen-pk host end to end on a GPU: BASELINE config 1 (bundled test_g2_snap),
the -s / -c / -j modes, a multi-species bigfile snapshot (config 4, reduced size) and the
synthetic generator, each against the CPU oracle on the bytes the reference's readers hand over."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import Oracle, have_reference, padded_shape, rfftn_padded
from tests.bigfile_writer import write_snapshot

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "genpk_b200", "bin", "gen-pk")
SNAP = os.path.join(ROOT, "oracle", "_ref", "test_g2_snap")
GOLD = os.path.join(ROOT, "tests", "golden")
TYPE_STR = {0: "by", 1: "DM", 2: "nu", 4: "st"}


@pytest.fixture(scope="module")
def orc(port):
    return Oracle("reference") if have_reference() else port


def gen_pk(*args):
    r = subprocess.run([BIN, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    return r


def read_pk(path):
    rows = np.loadtxt(path, ndmin=2)
    return rows[:, 0], rows[:, 1], rows[:, 2].astype(np.int64)


def assert_file_matches(path, power, count, keffs):
    k, p, c = read_pk(path)
    nz = count > 0
    assert np.array_equal(c, count[nz]), "mode counts differ"
    np.testing.assert_allclose(p, power[nz], rtol=1.2e-5)         # 1e-5 + the 7-digit %e of utils.cpp:17
    np.testing.assert_allclose(k, keffs[nz], rtol=1.2e-5)


def oracle_pk(orc, box, dims, parts_a, tm_a, parts_b=None, tm_b=None):
    """parts: list of (pos, masses, mass) deposited into one field."""
    def field_of(parts):
        f = np.zeros(padded_shape(dims))
        for pos, masses, mass in parts:
            orc.fieldize(box, dims, f, np.ascontiguousarray(pos, np.float32), masses, mass, 1)
        return rfftn_padded(f, dims)
    a = field_of(parts_a)
    b = None if parts_b is None else field_of(parts_b)
    _, p, c, k = orc.powerspectrum(dims, a, b, dims, tm_a, tm_a if tm_b is None else tm_b)
    return p, c, k


needs_snap = pytest.mark.skipif(not os.path.exists(SNAP + ".0"), reason="snapshot fixture not built (oracle/_ref)")


@needs_snap
def test_config1_per_type_spectra(tmp_path):
    """./gen-pk -i test_g2_snap -o out (BASELINE configs[0]) vs the reference's own objects."""
    gold = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    r = gen_pk("-i", SNAP, "-o", str(tmp_path), "--json", str(tmp_path / "t.json"))
    assert "FFT grid dimension: 32" in r.stdout
    assert "total_mass in type 0 = 717639" in r.stdout and "total_mass in type 1 = 167.363" in r.stdout
    for t in (0, 1, 4):
        path = tmp_path / f"PK-{TYPE_STR[t]}-test_g2_snap"
        assert path.exists()
        assert_file_matches(path, gold[f"power{t}"], gold[f"count{t}"], gold[f"keffs{t}"])
        assert len(open(path).read().splitlines()) == 29
        if have_reference():
            # SURVEY 7 protocol (4): byte for byte the file the reference's own print_pk() writes for its own spectrum
            # (our P(k) agrees with it to ~1e-12, far inside the seven digits of "%e")
            want = tmp_path / f"ref-PK-{TYPE_STR[t]}"
            Oracle("reference").print_pk(want, len(gold[f"power{t}"]), gold[f"keffs{t}"], gold[f"power{t}"], gold[f"count{t}"])
            assert open(path, "rb").read() == open(want, "rb").read(), f"PK file of type {t} differs from the reference's bytes"
    rep = json.load(open(tmp_path / "t.json"))
    assert rep["grid"] == 32 and [s["type"] for s in rep["spectra"]] == ["by", "DM", "st"]
    # deterministic mode gives the same files to print precision
    fx = tmp_path / "fixed"
    fx.mkdir()
    gen_pk("-i", SNAP, "-o", str(fx), "--fixed")
    assert_file_matches(fx / "PK-DM-test_g2_snap", gold["power1"], gold["count1"], gold["keffs1"])


@needs_snap
@pytest.mark.parametrize("gpus", [2, 4])
def test_config1_on_several_gpus_equals_one_gpu(tmp_path, gpus):
    """gen-pk --gpus N (x-slabs over N GPUs of the box; with fewer GPUs the slabs share devices): in the
    deterministic mode the N-GPU grid equals the single-GPU grid bit for bit, so the printed PK- files are the same
    to the last digit the reference prints (7 significant digits, utils.cpp:17); and they match the golden
    spectra the reference objects produced."""
    gold = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    one, many = tmp_path / "one", tmp_path / "many"
    one.mkdir()
    many.mkdir()
    gen_pk("-i", SNAP, "-o", str(one), "--fixed")
    gen_pk("-i", SNAP, "-o", str(many), "--fixed", "--gpus", str(gpus))
    for t in (0, 1, 4):
        name = f"PK-{TYPE_STR[t]}-test_g2_snap"
        a, b = open(one / name).read(), open(many / name).read()
        ka, pa, ca = read_pk(one / name)
        kb, pb, cb = read_pk(many / name)
        assert np.array_equal(ca, cb)
        np.testing.assert_allclose(pb, pa, rtol=2e-6)            # at most the last printed digit (summation order of the bins)
        np.testing.assert_allclose(kb, ka, rtol=2e-6)
        # the scale of the integer sums follows each type's mass unit (stellar masses of 1e-8 here), so every type
        # matches the reference's double-precision spectra
        assert_file_matches(many / name, gold[f"power{t}"], gold[f"count{t}"], gold[f"keffs{t}"])
        assert len(b.splitlines()) == len(a.splitlines()) == 29


@needs_snap
def test_stars_are_baryons_and_cross_modes(tmp_path, orc):
    gold = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    box, dims = float(gold["box"]), 32
    by = (gold["pos0"], gold["masses0"], 0.0)
    st = (gold["pos4"], gold["masses4"], 0.0)
    dm = (gold["pos1"], None, float(gold["mass"][1]))
    tm = {t: float(gold[f"total_mass{t}"]) for t in (0, 1, 4)}
    # -s 1: stars deposited into the baryon field, total_mass carries both "+1" (gen-pk.cpp:228-230)
    s_dir = tmp_path / "s"
    s_dir.mkdir()
    gen_pk("-i", SNAP, "-o", str(s_dir), "-s", "1")
    p, c, k = oracle_pk(orc, box, dims, [by, st], tm[0] + tm[4])
    assert_file_matches(s_dir / "PK-by-test_g2_snap", p, c, k)
    # -c 0: DM x baryons inside the snapshot (gen-pk.cpp:304-356)
    c_dir = tmp_path / "c"
    c_dir.mkdir()
    gen_pk("-i", SNAP, "-o", str(c_dir), "-c", "0")
    p, c, k = oracle_pk(orc, box, dims, [dm], tm[1], [by], tm[0])
    kk, pp, cc = read_pk(c_dir / "PK-DMxby-test_g2_snap")
    nz = c > 0
    assert np.array_equal(cc, c[nz])
    np.testing.assert_allclose(pp, p[nz], rtol=1e-4, atol=1e-7 * np.abs(p).max())     # cross power changes sign
    # -j: the same snapshot against itself is the auto spectrum (gen-pk.cpp:240-303)
    j_dir = tmp_path / "j"
    j_dir.mkdir()
    gen_pk("-i", SNAP, "-j", SNAP, "-o", str(j_dir))
    assert_file_matches(j_dir / "PX-DM-test_g2_snap", gold["power1"], gold["count1"], gold["keffs1"])
    assert not (j_dir / "PX-nu-test_g2_snap").exists()


def test_bigfile_multi_species(tmp_path, orc):
    """BASELINE configs[3] at reduced size: DM + gas (Mass block) + neutrinos from a bigfile."""
    rng = np.random.default_rng(12)
    box = 250.0
    n = 20 ** 3
    species = {1: (rng.random((n, 3)) * box, None, 0.8), 0: (rng.random((n, 3)) * box, 10.0 ** rng.uniform(-1, 0, n), 0.0),
               2: (rng.random((n // 2, 3)) * box, None, 0.05)}
    root = str(tmp_path / "PART_005")
    write_snapshot(root, species, box, nfile=3)
    out = tmp_path / "out"
    out.mkdir()
    r = gen_pk("-i", root, "-o", str(out))
    dims = 64                                                    # 8000 -> 20 -> 32 -> 64
    assert f"FFT grid dimension: {dims}" in r.stdout
    for t, (pos, masses, m) in species.items():
        pos32 = pos.astype(np.float32)
        m32 = None if masses is None else masses.astype(np.float32)
        tm = float(m32.astype(np.float64).sum()) if m32 is not None else m * len(pos)     # no "+1" on this path
        p, c, k = oracle_pk(orc, box, dims, [(pos32, m32, 0.0 if m32 is not None else m)], tm)
        assert_file_matches(out / f"PK-{TYPE_STR[t]}-PART_005", p, c, k)


def test_synthetic_input_matches_the_library_pipeline(tmp_path):
    import torch
    import genpk_b200 as gp
    from genpk_b200 import api
    n_side, dims, box = 32, 64, 1000.0
    gen_pk("--synthetic", f"clustered:{n_side}", "-g", str(dims), "-o", str(tmp_path), "--fixed")
    n = n_side ** 3
    d = torch.empty(3 * n, dtype=torch.float32, device="cuda")
    api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, box, dims, d.data_ptr())
    torch.cuda.synchronize()
    with gp.Context(dims, flags=api.FLAG_FIXED_POINT) as ctx:
        ctx.grid_zero()
        ctx.deposit_dev(d.data_ptr(), n, 0, 1.0, box)
        ctx.fft()
        p, c, k = ctx.power(dims, float(n), float(n))
        ctx.synchronize()
    assert_file_matches(tmp_path / f"PK-DM-synthetic-clustered-{n_side}", p, c.astype(np.int64), k)


@needs_snap
def test_fold_and_min_modes_extensions(tmp_path, orc):
    """Two items of the reference's own to-do list (gen-pk.cpp:27-31).  --fold F: the box folded F times onto itself is
    the deposit with a box F times smaller (every position is wrapped periodically, fieldize.cpp:70-75), k_eff printed
    in modes of the full box.  --min-modes N: neighbouring bins merged to at least N modes, mode-weighted means."""
    gold = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    box, dims, fold = float(gold["box"]), 32, 2
    out = tmp_path / "fold"
    out.mkdir()
    gen_pk("-i", SNAP, "-o", str(out), "--fold", str(fold))
    for t in (0, 1):
        masses = gold[f"masses{t}"] if f"masses{t}" in gold else None
        p, c, k = oracle_pk(orc, box / fold, dims, [(gold[f"pos{t}"], masses, float(gold["mass"][t]))], float(gold[f"total_mass{t}"]))
        assert_file_matches(out / f"PK-{TYPE_STR[t]}-test_g2_snap", p, c, k * fold)
    # merged bins: same modes, same mode-weighted sums as the unmerged file
    plain, merged = tmp_path / "plain", tmp_path / "merged"
    plain.mkdir()
    merged.mkdir()
    gen_pk("-i", SNAP, "-o", str(plain), "--fixed")
    gen_pk("-i", SNAP, "-o", str(merged), "--fixed", "--min-modes", "500")
    k0, p0, c0 = read_pk(plain / "PK-DM-test_g2_snap")
    k1, p1, c1 = read_pk(merged / "PK-DM-test_g2_snap")
    assert c1.sum() == c0.sum() == dims ** 3 - 1 and len(c1) < len(c0) and (c1 >= 500).all()
    np.testing.assert_allclose((p1 * c1).sum(), (p0 * c0).sum(), rtol=1e-5)
    np.testing.assert_allclose((k1 * c1).sum(), (k0 * c0).sum(), rtol=1e-5)
