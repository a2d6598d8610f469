"""Writes the bigfile on-disk format by hand (test helper; bigfile/src/bigfile.c:506-528
header, :1452-1513 attr-v2, raw little-endian data files %06X)."""
import os

import numpy as np


def _dtype_str(a):
    return a.dtype.newbyteorder("<").str if a.dtype.byteorder != "|" else a.dtype.str


def write_block(root, name, data=None, nfile=1, attrs=None):
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    if data is None:
        dtype, nmemb, parts = "<i8", 0, []
    else:
        data = np.ascontiguousarray(data)
        dtype = _dtype_str(data)
        nmemb = 1 if data.ndim == 1 else data.shape[1]
        parts = np.array_split(data, nfile)
    with open(os.path.join(d, "header"), "w") as f:
        f.write(f"DTYPE: {dtype}\nNMEMB: {nmemb}\nNFILE: {len(parts)}\n")
        for i, p in enumerate(parts):
            raw = p.astype(p.dtype.newbyteorder("<")).tobytes()
            f.write(f"{i:06X}: {len(p)} : {sum(raw) & 0xFFFFFFFF} : 0\n")
            with open(os.path.join(d, f"{i:06X}"), "wb") as g:
                g.write(raw)
    with open(os.path.join(d, "attr-v2"), "w") as f:
        for key, val in (attrs or {}).items():
            val = np.atleast_1d(np.asarray(val))
            raw = val.astype(val.dtype.newbyteorder("<")).tobytes()
            f.write(f"{key} {_dtype_str(val)} {val.size} {raw.hex().upper()} #HUMANE [ {' '.join(str(v) for v in val)} ]\n")


def write_snapshot(root, species, box, time=0.5, hubble=0.7, omega0=0.3, nfile=2, pos_dtype=np.float64):
    """species: {type: (positions[n,3], masses[n] or None, table_mass)}"""
    tot = np.zeros(6, np.uint64)
    table = np.zeros(6, np.float64)
    for t, (pos, masses, m) in species.items():
        tot[t] = len(pos)
        table[t] = 0.0 if masses is not None else m
    write_block(root, "Header", None, attrs={"TotNumPart": tot, "MassTable": table, "Time": np.float64(time),
                                             "HubbleParam": np.float64(hubble), "Omega0": np.float64(omega0),
                                             "BoxSize": np.float64(box)})
    for t, (pos, masses, m) in species.items():
        write_block(root, f"{t}/Position", np.asarray(pos, pos_dtype), nfile)
        if masses is not None:
            write_block(root, f"{t}/Mass", np.asarray(masses, np.float32), nfile)
