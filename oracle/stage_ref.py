"""TEST / BASELINE INFRASTRUCTURE — stages the UNMODIFIED reference modules of the hot path into oracle/_ref/.

The reference (yukyeongleee/pggan) is a pure-Python script tree: nothing to compile.  This recipe copies, byte for
byte, the eight files its PGGAN path imports from where they lie under /root/reference into oracle/_ref/ (git-ignored,
NOT gpurun-ignored: like a built .so it travels to the GPU box, where /root/reference does not exist), and writes a
manifest of their sha256 sums.  `bench.py --impl reference` and bench.py's `cpu_baseline` leg then time the REFERENCE'S
OWN Generator / Discriminator / WGANGPLoss (kind "reference") through oracle/ref_step.py; without oracle/_ref they
fall back to the oracle port (kind "port").  Nothing under oracle/ is imported by the product package.

    python oracle/stage_ref.py            # run in the build container (needs /root/reference)

No reference source enters the repository's history: oracle/_ref/ is listed in .gitignore.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("PGGAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

# what `from pggan.nets import ...`, `from pggan.loss import WGANGPLoss`, `from lib.loss import Loss` pull in
FILES = [
    "lib/layers.py", "lib/custom_layers.py", "lib/blocks.py", "lib/utils.py", "lib/loss.py",
    "pggan/nets.py", "pggan/loss.py",
    "submodel/lpips.py",          # imported (never instantiated) by lib/loss.py:2
]


def stage(verbose: bool = True) -> bool:
    """returns True if oracle/_ref is complete afterwards"""
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "yukyeongleee/pggan (unmodified copies)", "files": manifest}, fh, indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files into {DST}")
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json")) and all(
        os.path.exists(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    ok = stage()
    sys.exit(0 if ok else 1)
