"""The sampler tables shipped in pbrt-rust_b200/tables/ are the reference crate's own constants
(src/core/sobolmatrices.rs, converted by tools/convert_tables.py).  When the reference tree is present (the
build container) the conversion is re-done and compared; everywhere, structural known answers are checked."""
import re
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference/src/core/sobolmatrices.rs")


def test_table_shapes_and_known_answers(pkg):
    t = pkg.host.sampler_tables()
    assert t["sobol32"].shape == (1024 * 52,) and t["sobol32"].dtype == np.uint32
    assert t["vdc"].shape == (25 * 52,) and t["vdc_inv"].shape == (26 * 52,) and t["vdc"].dtype == np.uint64
    # dimension 0 of the Sobol' sequence is the van der Corput sequence: matrix column i = bit (31 - i)
    assert t["sobol32"][:32].tolist() == [0x80000000 >> i for i in range(32)]
    # jagged rows: M_m has 52 - 2m entries, MI_m has 2m entries; the rest is zero padding
    vdc, vdci = t["vdc"].reshape(25, 52), t["vdc_inv"].reshape(26, 52)
    for m in range(1, 26):
        assert np.all(vdc[m - 1, 52 - 2 * m:] == 0) and np.all(vdc[m - 1, :52 - 2 * m] != 0)
    for m in range(1, 27):
        assert np.all(vdci[m - 1, 2 * m:] == 0) and np.all(vdci[m - 1, :2 * m] != 0)


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the build container")
def test_tables_equal_the_reference_source(pkg):
    text = REF.read_text()
    m = re.search(r"SOBOL_MATRICES_32\s*:\s*\[[^\]]*\]\s*=\s*\[(.*?)\];", text, re.S)
    vals = [int(x.replace("_", ""), 0) for x in re.findall(r"0x[0-9a-fA-F_]+|\b\d[\d_]*\b", re.sub(r"_?u32", "", m.group(1)))]
    assert len(vals) == 1024 * 52
    assert np.array_equal(np.array(vals, np.uint64).astype(np.uint32), pkg.host.sampler_tables()["sobol32"])
