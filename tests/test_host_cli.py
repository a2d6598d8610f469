"""CPU-only checks of the C++ gen-pk host (genpk_b200/host): flags, snapshot readers.
The readers must hand the deposit the same float32 bytes the reference's adapters do."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import padded_shape
from tests.bigfile_writer import write_snapshot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "genpk_b200", "bin", "gen-pk")
SNAP = os.path.join(ROOT, "oracle", "_ref", "test_g2_snap")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module", autouse=True)
def binary():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", os.path.join(ROOT, "genpk_b200", "csrc"), "-j8"], check=True, stdout=subprocess.DEVNULL)
        subprocess.run(["make", "-C", os.path.join(ROOT, "genpk_b200", "host")], check=True, stdout=subprocess.DEVNULL)
    return BIN


def run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True)


def test_help_and_missing_arguments():
    r = run("-h")
    assert r.returncode == 0 and "Usage: ./gen-pk -i filenames" in r.stderr      # utils.cpp:44-55
    r = run("-i", "nothing")                                                      # no outdir: help, gen-pk.cpp:153-156
    assert r.returncode == 0 and "Usage" in r.stderr


@pytest.mark.skipif(not os.path.exists(SNAP + ".0"), reason="snapshot fixture not built (oracle/_ref)")
def test_gadget_header_and_grid_rule():
    r = run("-i", SNAP, "--info")
    assert r.returncode == 0
    out = r.stdout.splitlines()
    assert out[0] == "Boxsize=3000, NPart=(15.9254,16,0,0,3.8485,0)**3"            # gen-pk.cpp:163-166
    assert out[1] == "Masses=[0 0.0406161 0 ]"
    assert out[2].startswith("redshift=2.6")
    assert out[3] == "FFT grid dimension: 32"                                       # 4039 -> 15 -> 16 -> 32
    assert run("-i", SNAP, "--info", "-g", "48").stdout.splitlines()[3] == "FFT grid dimension: 48"


@pytest.mark.skipif(not os.path.exists(SNAP + ".0"), reason="snapshot fixture not built (oracle/_ref)")
@pytest.mark.parametrize("ptype", [0, 1, 4])
def test_gadget_reader_hands_over_the_reference_bytes(tmp_path, ptype):
    """positions / masses / total_mass equal what the reference's GadgetReader +
    read_fieldize produce (golden vectors), incl. the MASS offset quirk and the +1."""
    gold = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    dump = str(tmp_path / "p")
    r = run("-i", SNAP, "--dump", str(ptype), dump)
    assert r.returncode == 0, r.stderr
    pos = np.fromfile(dump, np.float32).reshape(-1, 3)
    assert np.array_equal(pos.view(np.uint32), gold[f"pos{ptype}"].view(np.uint32))
    if f"masses{ptype}" in gold:
        m = np.fromfile(dump + ".mass", np.float32)
        assert np.array_equal(m.view(np.uint32), gold[f"masses{ptype}"].view(np.uint32))
    tm = float(r.stdout.strip().splitlines()[-1].split("=")[1])
    assert tm == float(gold[f"total_mass{ptype}"])
    assert run("-i", SNAP, "--dump", "2", dump).returncode != 0                     # type not present


def test_bigfile_reader_matches_reference_library(tmp_path, ref, port):
    """A hand-written bigfile snapshot (f8 positions in 2 files, one species with a Mass
    block): our reader's float32 output equals the narrowed arrays, and the reference's own
    read_fieldize_bigfile() on the same directory deposits the same grid."""
    rng = np.random.default_rng(4)
    box, dims = 100.0, 16
    species = {1: (rng.random((5000, 3)) * box, None, 0.8),
               0: (rng.random((3001, 3)) * box, 10.0 ** rng.uniform(-2, 0, 3001), 0.0)}
    root = str(tmp_path / "snap_000")
    write_snapshot(root, species, box)
    info = run("-i", root, "--info")
    assert info.returncode == 0, info.stderr
    assert info.stdout.splitlines()[0].startswith("NumPart=[3001,5000,0,0,0,0], Masses=[0 0.8 0 0 0 0]")
    assert "FFT grid dimension: 64" in info.stdout                   # 5000 -> 17 -> 32 -> 64
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgenpk_ref.so"))
    lib.read_fieldize_bigfile.restype = C.c_int
    npart = (C.c_int64 * 6)(3001, 5000, 0, 0, 0, 0)
    mass = (C.c_double * 6)(0.0, 0.8, 0, 0, 0, 0)
    for t, (pos, masses, m) in species.items():
        dump = str(tmp_path / f"d{t}")
        r = run("-i", root, "--dump", str(t), dump)
        assert r.returncode == 0, r.stderr
        got = np.fromfile(dump, np.float32).reshape(-1, 3)
        assert np.array_equal(got, pos.astype(np.float32))
        gm = None
        if masses is not None:
            gm = np.fromfile(dump + ".mass", np.float32)
            assert np.array_equal(gm, masses.astype(np.float32))
        tm_ours = float(r.stdout.strip().splitlines()[-1].split("=")[1])
        field = np.zeros(padded_shape(dims))
        tm = C.c_double(0.0)
        rc = lib.read_fieldize_bigfile(field.ctypes.data_as(C.c_void_p), root.encode(), C.c_int(t), C.c_double(box),
                                       C.c_int(dims), C.byref(tm), npart, mass, C.c_double(0.3))
        assert rc == 0
        assert tm_ours == tm.value
        want = np.zeros(padded_shape(dims))
        port.fieldize(box, dims, want, got, gm, m, 1)
        np.testing.assert_allclose(field, want, rtol=1e-12, atol=1e-14)


@pytest.mark.skipif(not os.path.exists(SNAP + ".0"), reason="snapshot fixture not built (oracle/_ref)")
@pytest.mark.parametrize("env", [{"GENPK_READ_CHUNK": "1000"}, {"GENPK_READ_CHUNK": "97", "GENPK_READ_THREADS": "1"},
                                 {"GENPK_READ_THREADS": "5"}])
def test_read_pipeline_is_independent_of_chunking_and_threads(tmp_path, env):
    """The double-buffered reader (a thread one chunk ahead, pieces of a chunk in different files read in parallel) hands
    over the same bytes and the same total_mass whatever the chunk size and the number of read threads."""
    gold = np.load(os.path.join(GOLD, "test_g2_snap.npz"))
    for ptype in (0, 1, 4):
        dump = str(tmp_path / f"p{ptype}")
        r = subprocess.run([BIN, "-i", SNAP, "--dump", str(ptype), dump], capture_output=True, text=True,
                           env={**os.environ, **env})
        assert r.returncode == 0, r.stderr
        pos = np.fromfile(dump, np.float32).reshape(-1, 3)
        assert np.array_equal(pos.view(np.uint32), gold[f"pos{ptype}"].view(np.uint32))
        if f"masses{ptype}" in gold:
            m = np.fromfile(dump + ".mass", np.float32)
            assert np.array_equal(m.view(np.uint32), gold[f"masses{ptype}"].view(np.uint32))
        tm = float(r.stdout.strip().splitlines()[-1].split("=")[1])
        assert tm == float(gold[f"total_mass{ptype}"])


@pytest.mark.skipif(not os.path.exists(SNAP + ".0"), reason="snapshot fixture not built (oracle/_ref)")
def test_a_truncated_snapshot_file_is_an_error_not_a_crash(tmp_path):
    """The parallel read falls back to the reference's file-by-file semantics when a piece is short; the host then
    reports the read error (read_fieldize.cpp:56-59) instead of depositing garbage."""
    import shutil
    base = str(tmp_path / "snap")
    for i in (0, 1):
        shutil.copy(f"{SNAP}.{i}", f"{base}.{i}")
    with open(base + ".1", "r+b") as f:
        f.truncate(1000)                                             # header and the start of the POS block only
    r = run("-i", base, "--dump", "1", str(tmp_path / "d"))
    assert r.returncode != 0
    assert "Error reading particle data for type 1" in r.stderr
