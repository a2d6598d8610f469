"""Regenerates the golden vectors in this directory from the REFERENCE's own
object code (oracle/_ref/libgenpk_ref.so, built by oracle/Makefile from the
unmodified sources under /root/reference).  Run in the build container:

    python tests/golden/make_golden.py

The FFT between deposit and binning is pocketfft (FFTW3, the reference's
transform, is absent from the image; see oracle/oracle.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.oracle import Oracle, RefSnapshot, build, padded_shape, rfftn_padded  # noqa: E402


def fieldize_cases(ref):
    rng = np.random.default_rng(20261017)
    out = {}
    for name, dims, n, box, lo, hi in [("a", 8, 500, 25.0, 0.0, 1.0), ("b", 16, 3000, 3000.0, -0.3, 1.4),
                                       ("c", 5, 64, 10.0, 0.0, 1.0)]:
        pos = ((rng.random((n, 3)) * (hi - lo) + lo) * box).astype(np.float32)
        masses = (10.0 ** rng.uniform(-3, 2, n)).astype(np.float32)
        g_const = np.zeros(padded_shape(dims))
        ref.fieldize(box, dims, g_const, pos, None, 0.75, 1)
        g_var = np.zeros(padded_shape(dims))
        ref.fieldize(box, dims, g_var, pos, masses, 0.0, 1)
        out.update({f"{name}_dims": dims, f"{name}_box": box, f"{name}_pos": pos, f"{name}_masses": masses,
                    f"{name}_grid_const": g_const, f"{name}_grid_var": g_var})
    return out


def powerspectrum_cases(ref):
    rng = np.random.default_rng(7)
    out = {}
    for name, dims, nrbins in [("s16", 16, 16), ("s12", 12, 7), ("s32", 32, 32)]:
        shape = (dims, dims, dims // 2 + 1)
        a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        b = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        _, p, c, k = ref.powerspectrum(dims, a, None, nrbins, 3.0, 3.0)
        _, px, cx, kx = ref.powerspectrum(dims, a, b, nrbins, 3.0, 5.0)
        out.update({f"{name}_dims": dims, f"{name}_nrbins": nrbins, f"{name}_a": a, f"{name}_b": b, f"{name}_power": p,
                    f"{name}_count": c, f"{name}_keffs": k, f"{name}_xpower": px, f"{name}_xcount": cx,
                    f"{name}_xkeffs": kx})
    return out


def snapshot_case(ref):
    """gen-pk -i test_g2_snap -o out (gen-pk.cpp:202-239) through the reference's
    reader, deposit and binning; positions/masses are stored exactly as the
    reference's GadgetReader hands them to fieldize() (incl. SURVEY App. D-2)."""
    snap = RefSnapshot()
    assert snap.numfiles() == 2
    npart = [snap.npart(t) for t in range(6)]
    mass = [snap.mass(t) for t in range(6)]
    box = snap.box()
    out = {"npart": np.array(npart), "mass": np.array(mass), "box": box, "redshift": snap.redshift(),
           "omega0": snap.omega0()}
    dims = 32                       # gen-pk.cpp:169-172 on these particle counts
    for t in range(6):
        if npart[t] == 0:
            continue
        skip = (1 << 6) - 1 - (1 << t)                              # read_fieldize.cpp:27,37
        got, pos = snap.get_block("POS ", npart[t], 0, skip, 3)
        assert got == npart[t]
        out[f"pos{t}"] = pos.reshape(-1, 3)
        if mass[t] == 0:
            got, m = snap.get_block("MASS", npart[t], 0, skip, 1)
            assert got == npart[t]
            out[f"masses{t}"] = m
        field = np.zeros(padded_shape(dims))
        rc, total_mass = snap.read_fieldize(field, t, box, dims, 0.0)
        assert rc == 0
        out[f"total_mass{t}"] = total_mass
        out[f"grid{t}"] = field
        spec = rfftn_padded(field, dims)
        _, p, c, k = ref.powerspectrum(dims, spec, None, dims, total_mass, total_mass)
        out[f"power{t}"], out[f"count{t}"], out[f"keffs{t}"] = p, c, k
    # check_read_fieldize, test.cpp:88-100: baryons then stars into one 4^3 field
    f4 = np.zeros(padded_shape(4))
    rc, tm = snap.read_fieldize(f4, 0, 3000.0, 4, 0.0)
    rc2, tm = snap.read_fieldize(f4, 4, 3000.0, 4, tm)
    assert rc == 0 and rc2 == 0
    out["rf4_grid"], out["rf4_total_mass"] = f4, tm
    snap.close()
    return out


def main():
    build(ref=True)
    ref, port = Oracle("reference"), Oracle("port")
    np.savez_compressed(os.path.join(HERE, "fieldize_cases.npz"), **fieldize_cases(ref))
    np.savez_compressed(os.path.join(HERE, "powerspectrum_cases.npz"), **powerspectrum_cases(ref))
    np.savez_compressed(os.path.join(HERE, "test_g2_snap.npz"), **snapshot_case(ref))
    # Mode counts of the large BASELINE grids (nrbins = dims, gen-pk.cpp:173).  The
    # reference's powerspectrum() on a zero field gives the same integers up to 512^3
    # (tests/test_oracle.py); above that the count-only restatement is used.
    counts = {}
    for d in (512, 1024, 2048):
        c, ks = port.mode_counts(d, d)
        counts[f"count{d}"] = c
        counts[f"ksum{d}"] = ks
    np.savez_compressed(os.path.join(HERE, "mode_counts.npz"), **counts)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
