"""Recipe for ``oracle/_ref``: a byte-for-byte copy of the reference's Python package, so that the UNMODIFIED reference
can be timed on the GPU box's host cores (``bench.py --impl reference``) and re-checked against the oracle there.

    python oracle/make_ref.py            # build container only: needs /root/reference

Nothing is built or patched: ``multimodal_flows/**/*.py`` is copied as is and a SHA-256 manifest is written next to it.
``oracle/_ref/`` is git-ignored (reference sources never enter the history) but not gpurun-ignored, so it travels with the
snapshot like the built ``.so`` files.  The reference's own ``setup.py`` is of no use here: ``pip install --no-deps --target``
succeeds (with VERSION set) but installs only a dist-info, because ``find_packages`` sees no ``__init__.py`` under
``multimodal_flows``; the scripts import the sub-packages by putting that directory on ``sys.path``, and so does
``oracle/ref_loader.py``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("MMF_REFERENCE_SRC", "/root/reference"), "multimodal_flows")
DST = os.path.join(HERE, "_ref", "multimodal_flows")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not found: oracle/_ref left as it is")
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    lines = []
    for dirpath, _, files in sorted(os.walk(SRC)):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            src = os.path.join(dirpath, f)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            lines.append(f"{hashlib.sha256(open(src, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(HERE, "_ref", "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print(f"copied {len(lines)} files to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
