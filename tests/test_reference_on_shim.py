"""The reference's OWN model / encoder / decoder code, unmodified, on the product's MinkowskiEngine + torchac drop-in
modules (the L2->L1 boundary).  Needs a reference checkout (LINR_REFERENCE_DIR, default /root/reference):
in the build container (no GPU) the device entry points under the shim are oracle doubles; on a box with BOTH a GPU and a
checkout the driver replaces nothing and the reference runs on the real kernels.  The GPU boxes of this project have no
checkout, so there the same shim is checked op by op against the oracle in tests/test_gpu_me_shim.py."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("LINR_REFERENCE_DIR", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="needs the read-only reference checkout")
def test_reference_modules_run_unmodified_on_the_shim():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_on_shim_driver.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    r = json.loads(line[len("RESULT "):])
    assert r["n_params"] == 53112                                   # 3 scales (54,712 with the 7 scales of loot)
    tol = 1e-4 if r["on_gpu"] else 1e-5                             # fp32 on the device: north_star tolerance
    assert abs(r["bits"] - r["bits_fixture"]) <= tol * abs(r["bits_fixture"])
    assert r["grad_max_abs_err"] <= tol * r["grad_max_abs"] + 1e-7
    assert r["probs_max_abs_err"] <= (1e-4 if r["on_gpu"] else 1e-6)
    if r["on_gpu"]:
        assert abs(r["all_bit"] - r["all_bit_fixture"]) <= 0.005 * r["all_bit_fixture"]   # bpp within 0.5 %
    else:
        assert r["bytes_equal"] and r["all_bit"] == r["all_bit_fixture"]   # same bitstreams as the recorded reference flow
    assert r["lossless"]                                            # decoder.py:140
    assert r["tables_built"] <= 64                                  # kernel maps are cached per coordinate tensor
