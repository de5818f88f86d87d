"""ORACLE (test infrastructure, not product code).

Recipe for oracle/_ref: a verbatim, UNMODIFIED copy of the reference's own Python
modules for the hot path (package `sedt`, `utilities`, `config.py` and the metadata
tsv files `config.py` reads at import time), taken from /root/reference.

    python oracle/make_ref.py            # in the build container (where /root/reference exists)

oracle/_ref/ is listed in .gitignore (reference sources never enter this repo's
history) but NOT in .gpurunignore, so the copy travels to the GPU box with the
snapshot, where /root/reference does not exist.  Users: bench.py's `--impl reference`
arm and its `gpu_eager_baseline` leg (the reference itself instead of the oracle
port), tests that compare the oracle against the live reference.  The product
package never imports it.  __graft_entry__.build() runs this when the reference is
present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SEDT_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
WHAT = ("sedt", "utilities", "config.py", "data")     # data/: metadata tsv only (5.7 MB); config.py:63-65 reads two of them


def make(force: bool = False) -> str | None:
    """Copies the reference modules; returns the destination or None when the reference is not available."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, "sedt")) else None
    stamp = os.path.join(REF_DST, ".copied_from")
    if not force and os.path.exists(stamp) and open(stamp).read().strip() == REF_SRC:
        return REF_DST
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    for name in WHAT:
        src, dst = os.path.join(REF_SRC, name), os.path.join(REF_DST, name)
        if os.path.isdir(src):
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.wav", "*.npy"))
        elif os.path.exists(src):
            shutil.copy2(src, dst)
    for root, dirs, files in os.walk(REF_DST):        # the reference tree is read-only: make the copy removable
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    with open(stamp, "w") as f:
        f.write(REF_SRC + "\n")
    return REF_DST


if __name__ == "__main__":
    out = make(force="--force" in sys.argv)
    print(out or f"{REF_SRC} not found and no previous copy under {REF_DST}")
