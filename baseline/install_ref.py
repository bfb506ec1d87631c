"""Place the UNMODIFIED reference modules of the hot path under baseline/_ref/ (git-ignored, travels to the GPU box).

The reference (Max-Manning/passiveRadar) is plain Python without setup.py / pyproject.toml, so
``pip install --target baseline/_ref /root/reference`` has nothing to build; this recipe is the equivalent: it copies
the package directory byte for byte.  ``bench.py --impl reference`` and ``cpu_baseline`` import
``passiveRadar.clutter_removal.LS_Filter`` / ``NLMS_filter`` and ``passiveRadar.range_doppler_processing.fast_xambg``
from there (kind "reference"); when baseline/_ref is absent they time the oracle port instead (kind "port").

    python baseline/install_ref.py [/root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["__init__.py", "clutter_removal.py", "range_doppler_processing.py", "signal_utils.py"]


def install(src_root="/root/reference"):
    src = os.path.join(src_root, "passiveRadar")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DEST, "passiveRadar")
    os.makedirs(dst, exist_ok=True)
    digest = {}
    for f in FILES:
        a = os.path.join(src, f)
        if not os.path.exists(a):
            if f == "__init__.py":
                open(os.path.join(dst, f), "w").close()
            continue
        shutil.copyfile(a, os.path.join(dst, f))
        with open(a, "rb") as fh:
            digest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "sha256": digest}, fh, indent=1)
    return dst


if __name__ == "__main__":
    out = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print(out if out else "reference tree not found; nothing installed")
