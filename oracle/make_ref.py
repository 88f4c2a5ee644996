#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: a git-ignored working copy of the reference's Python package.

    python oracle/make_ref.py            (build container only: /root/reference must exist)

The reference is pure Python, so "building" it for the GPU box means copying the package files
it imports (``codes/**/*.py``, the per-dataset ``surrogates_config.py`` files and the two run
configs) from where they lie under /root/reference into ``oracle/_ref/``.  ``oracle/_ref/`` is
listed in .gitignore (reference sources never enter the history) but not in .gpurunignore, so the
copy travels with the gpurun snapshot exactly like a compiled ``.so``.  It is imported through
``oracle/ref_shim.py`` by ``bench.py --impl reference`` (the reference's own
``MultiONet.predict`` on the host cores) and by the ``reference_gpu`` comparator leg.
TEST / BASELINE INFRASTRUCTURE ONLY.
"""
import os
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def make(verbose: bool = True) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "codes")):
        if verbose:
            print(f"{SRC}/codes not present: keeping the existing {DST} (if any)")
        return DST if os.path.isdir(os.path.join(DST, "codes")) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for sub, pattern in (("codes", (".py",)), ("datasets", ("surrogates_config.py", "data_sources.yaml"))):
        for dirpath, _, files in os.walk(os.path.join(SRC, sub)):
            rel = os.path.relpath(dirpath, SRC)
            if "__pycache__" in rel or "_data_generation" in rel or "_data_analysis" in rel:
                continue
            for f in files:
                if f.endswith(pattern):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copy2(os.path.join(dirpath, f), os.path.join(DST, rel, f))
                    n += 1
    for f in ("config.yaml", "config_full.yaml", "LICENSE"):
        if os.path.isfile(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
            n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write("Unmodified files copied from /root/reference (robin-janssen/CODES-Benchmark) by oracle/make_ref.py.\n"
                 "Git-ignored; baseline/test infrastructure only.\n")
    if verbose:
        print(f"oracle/_ref: {n} files copied from {SRC}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
