#!/usr/bin/env python3
"""Convert the reference's sampler tables into one binary blob.

Reads (read-only) /root/reference/src/core/sobolmatrices.rs and emits
pbrt-rust_b200/tables/sobol_tables.npz with

  sobol32   u32[1024*52]   SOBOL_MATRICES_32            (sobolmatrices.rs:5)
  vdc       u64[25*52]     VD_C_SOBOL_MATRICES, rows M1..M25 zero-padded to 52 (:26636-26842)
  vdc_inv   u64[26*52]     VD_C_SOBOL_MATRICES_INV, rows MI1..MI26 padded to 52 (:26846-27534)

The reference's Rust host owns these tables; across the C ABI they are *inputs*
(pbrt_b200_sampler_desc), so neither the library nor the oracle embeds them.
This script only runs in the build container (the GPU box has no /root/reference);
its output is committed.  Zero padding is safe: the reference would panic, not
read, past the end of a jagged row.
"""
import re
import sys
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/src/core/sobolmatrices.rs")
OUT = Path(__file__).resolve().parent.parent / "pbrt-rust_b200" / "tables" / "sobol_tables.npz"

NUM = re.compile(r"0x[0-9a-fA-F_]+|\b\d[\d_]*\b")


def parse_array(text, name):
    m = re.search(r"const\s+%s\s*:\s*\[[^\]]*\]\s*=\s*\[(.*?)\];" % re.escape(name), text, re.S)
    if not m:
        raise SystemExit("array %s not found" % name)
    body = re.sub(r"_u64|_u32", "", m.group(1))
    vals = []
    for tok in NUM.findall(body):
        tok = tok.replace("_", "")
        vals.append(int(tok, 16) if tok.startswith("0x") else int(tok))
    return vals


def main():
    text = SRC.read_text()
    s32 = parse_array(text, "SOBOL_MATRICES_32")
    assert len(s32) == 1024 * 52, len(s32)
    vdc = np.zeros((25, 52), dtype=np.uint64)
    for m in range(1, 26):
        row = parse_array(text, "M%d" % m)
        assert len(row) == 52 - 2 * m, (m, len(row))
        vdc[m - 1, : len(row)] = row
    inv = np.zeros((26, 52), dtype=np.uint64)
    for m in range(1, 27):
        row = parse_array(text, "MI%d" % m)
        assert len(row) == 2 * m, (m, len(row))
        inv[m - 1, : len(row)] = row
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, sobol32=np.array(s32, dtype=np.uint32), vdc=vdc.reshape(-1), vdc_inv=inv.reshape(-1))
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    sys.exit(main())
