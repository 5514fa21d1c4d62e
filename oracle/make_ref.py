"""Make the UNMODIFIED reference travel to the GPU box: copy its Python modules from /root/reference into the
git-ignored ``baseline/_ref/`` (listed in .gitignore, NOT in .gpurunignore, so a ``gpurun`` snapshot carries it).

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing is vendored into the repository history and nothing under
``timetuning_b200/`` reads this directory.  Users: ``oracle/ref_loader.py`` (falls back to this copy when
/root/reference is absent), ``tests/test_gpu_reference_dropin.py`` (the drop-in boundary exercised against the real
reference modules on the GPU) and ``bench.py --impl reference`` / ``cpu_baseline`` (the reference's own CPU path,
``kind: "reference"``).  Called by ``__graft_entry__.build()`` whenever /root/reference is present.

    python oracle/make_ref.py
"""
from __future__ import annotations

import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("TIMET_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def make(verbose: bool = True) -> bool:
    """Copy the reference's flat *.py modules (+ LICENSE).  Returns False when /root/reference is absent."""
    if not os.path.isfile(os.path.join(SRC, "mask_propagation.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(SRC)):
        if name.endswith(".py") or name == "LICENSE":
            s, d = os.path.join(SRC, name), os.path.join(DST, name)
            if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
                n += 1
    if verbose:
        print(f"baseline/_ref: {n} file(s) refreshed from {SRC}")
    return True


if __name__ == "__main__":
    if not make():
        raise SystemExit(f"{SRC} not found")
