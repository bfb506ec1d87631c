"""The committed goldens are outputs of the REFERENCE: where the reference tree is available (the build
container), re-run tests/golden/make_golden.py for a few small cases into a scratch directory and require the
regenerated arrays to equal the committed ones.  Skipped on the GPU box, which has no /root/reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _golden as G

REFERENCE = os.environ.get("PR_REFERENCE", "/root/reference")
SCRIPT = os.path.join(G.GOLDEN_DIR, "make_golden.py")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "passiveRadar")),
                                reason="reference tree not available")


@pytest.mark.parametrize("name", ["xambg_small_kaiser", "ls_small", "nlms_peek0", "toep_small", "multi_fs_odd",
                                  "cfar_odd", "direct_small", "front_f32_short", "resample_c128"])
def test_committed_golden_is_what_the_reference_produces(name, tmp_path):
    env = dict(os.environ, PR_GOLDEN_OUT=str(tmp_path), PR_REFERENCE=REFERENCE)
    subprocess.run([sys.executable, SCRIPT, "--only", name], check=True, env=env, capture_output=True, timeout=600)
    fresh_path = tmp_path / (name + ".npz")
    assert fresh_path.exists(), f"make_golden.py --only {name} wrote nothing"
    committed = G.load(name)
    with np.load(fresh_path, allow_pickle=False) as z:
        fresh = {k: z[k] for k in z.files}
    assert set(fresh) == set(committed)
    for k in committed:
        if k == "seconds":
            continue                                   # wall time of the reference call
        np.testing.assert_array_equal(fresh[k], committed[k], err_msg=f"{name}[{k}]")
