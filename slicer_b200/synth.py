"""Synthetic GADGET format-2 snapshots, as the reference reads them.

Layout follows SURVEY.md App. B, i.e. what `readHeader` (gadget2io.cpp:8-31), `Block` / `Header`
(data.h:59-95), `fastforwardToBlock` (gadget2io.cpp:133-165), `readPos` (gadget2io.cpp:189-202) and the
mass reads in `mapParticles` (densitymaps.cpp:358-370) expect:

  every block = tag  [i32 8]["NAME"][i32 size+8][i32 8]  +  payload [i32 size][bytes][i32 size]
  HEAD payload = 256-byte header;  "POS " = float32 xyz triplets, the six types concatenated in type order;
  "MASS" = float32 masses of the types with massarr == 0, in type order;
  "BHMA" = float32 x npart[5] (read instead of the type-5 part of MASS).

Used by tests, bench.py and the C++ driver's examples to make inputs; not part of the hot path.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional, Sequence

import numpy as np

HEADER_FMT = "<6i6d2d2i6I2i4d2i6i1i"
HEADER_BYTES = 256


def pack_header(npart, massarr, redshift, npart_total, numfiles, boxsize, om0, oml, h) -> bytes:
    a = 1.0 / (1.0 + redshift)
    raw = struct.pack(
        HEADER_FMT,
        *[int(v) for v in npart],
        *[float(v) for v in massarr],
        a,
        float(redshift),
        0,
        0,
        *[int(v) & 0xFFFFFFFF for v in npart_total],
        0,
        int(numfiles),
        float(boxsize),
        float(om0),
        float(oml),
        float(h),
        0,
        0,
        *[int(v) >> 32 for v in npart_total],
        0,
    )
    return raw + b"\0" * (HEADER_BYTES - len(raw))


def _write_block(f, name: str, payload_bytes: int, writer) -> None:
    assert len(name) == 4
    f.write(struct.pack("<i", 8))
    f.write(name.encode("ascii"))
    f.write(struct.pack("<i", payload_bytes + 8))
    f.write(struct.pack("<i", 8))
    f.write(struct.pack("<i", payload_bytes))
    writer(f)
    f.write(struct.pack("<i", payload_bytes))


def split_counts(n: int, numfiles: int):
    """Contiguous split of n particles over sub-files (remainder spread over the first files)."""
    base, rem = divmod(n, numfiles)
    counts = [base + (1 if i < rem else 0) for i in range(numfiles)]
    offs = np.concatenate([[0], np.cumsum(counts)])
    return counts, offs


def write_snapshot(
    path_base: str,
    pos: Dict[int, np.ndarray],
    massarr: Sequence[float],
    redshift: float,
    boxsize: float,
    om0: float = 0.3,
    oml: float = 0.7,
    h: float = 0.7,
    numfiles: int = 1,
    masses: Optional[Dict[int, np.ndarray]] = None,
    bh_masses: Optional[np.ndarray] = None,
    with_vel_id: bool = True,
) -> list:
    """Write `<path_base>.<ff>` for ff in range(numfiles).

    pos[t]     float32 [n_t, 3] in the units of `boxsize` (kpc/h with POS_U 1.0, gadget2io.h:14)
    massarr[t] header mass of type t; 0 means per-particle masses from `masses[t]` (MASS block)
    bh_masses  type-5 masses for the "BHMA" block (the MASS block then carries `masses[5]` too,
               which the reference skips: densitymaps.cpp:361-365)
    Returns the list of files written.
    """
    masses = masses or {}
    os.makedirs(os.path.dirname(os.path.abspath(path_base)), exist_ok=True)
    ntot = [int(pos[t].shape[0]) if t in pos else 0 for t in range(6)]
    splits = {t: split_counts(ntot[t], numfiles) for t in range(6)}
    files = []
    for ff in range(numfiles):
        npart = [splits[t][0][ff] for t in range(6)]
        nall = sum(npart)
        name = f"{path_base}.{ff}"
        with open(name, "wb") as f:
            hdr = pack_header(npart, massarr, redshift, ntot, numfiles, boxsize, om0, oml, h)
            _write_block(f, "HEAD", HEADER_BYTES, lambda fh: fh.write(hdr))

            def wpos(fh):
                for t in range(6):
                    if npart[t]:
                        o = splits[t][1][ff]
                        np.ascontiguousarray(pos[t][o : o + npart[t]], dtype="<f4").tofile(fh)

            _write_block(f, "POS ", nall * 12, wpos)
            if with_vel_id:
                _write_block(f, "VEL ", nall * 12, lambda fh: fh.write(b"\0" * (nall * 12)))
                _write_block(f, "ID  ", nall * 4, lambda fh: np.arange(nall, dtype="<u4").tofile(fh))
            mtypes = [t for t in range(6) if npart[t] and float(massarr[t]) == 0.0]
            if mtypes:
                nm = sum(npart[t] for t in mtypes)

                def wmass(fh):
                    for t in mtypes:
                        o = splits[t][1][ff]
                        np.ascontiguousarray(masses[t][o : o + npart[t]], dtype="<f4").tofile(fh)

                _write_block(f, "MASS", nm * 4, wmass)
            if npart[5] and float(massarr[5]) == 0.0:
                o = splits[5][1][ff]
                bh = bh_masses if bh_masses is not None else masses[5]
                _write_block(
                    f, "BHMA", npart[5] * 4, lambda fh: np.ascontiguousarray(bh[o : o + npart[5]], dtype="<f4").tofile(fh)
                )
        files.append(name)
    return files


def uniform_positions(n: int, boxsize: float, seed: int) -> np.ndarray:
    """U[0, L) float32 positions (SURVEY.md App. C generator)."""
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * np.float32(boxsize)).astype(np.float32)


def clustered_positions(n: int, boxsize: float, seed: int, nclumps: int = 64, sigma_frac: float = 0.01) -> np.ndarray:
    """Sum-of-Gaussians positions wrapped into the box: stresses atomic contention in the deposit."""
    rng = np.random.default_rng(seed)
    centres = rng.random((nclumps, 3)) * boxsize
    which = rng.integers(0, nclumps, size=n)
    p = centres[which] + rng.normal(0.0, sigma_frac * boxsize, size=(n, 3))
    p = np.mod(p, boxsize).astype(np.float32)
    p[p >= np.float32(boxsize)] = 0.0
    return p


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_positions(n: int, boxsize: float, seed: int, start: int = 0) -> np.ndarray:
    """Host restatement of the device generator (csrc/aux_kernels.cuh: synth_mix / synth_positions_kernel):
    coordinate k of particle i = float(splitmix64(seed, 3 i + k) >> 40) * 2^-24 * float(boxsize).
    Any chunk [start, start+n) can be regenerated without storing the snapshot."""
    idx = (np.arange(start * 3, (start + n) * 3, dtype=np.uint64) + np.uint64(1))
    with np.errstate(over="ignore"):
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (u * np.float32(boxsize)).astype(np.float32).reshape(n, 3)
