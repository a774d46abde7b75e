#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's own hot-path modules, byte for byte, so that they travel to the GPU box.

    python oracle/build_ref.py            (also run by __graft_entry__.build() whenever /root/reference exists)

The reference's hot path is pure Python (S2VTModel.py, utils.py, attention_baseline.py: torch library calls only), so
"building" it is a verbatim copy of those three files from where they lie under /root/reference into oracle/_ref/, which is
git-ignored (the sources never enter this repository's history) but not gpurun-ignored.  bench.py's reference arm and CPU /
GPU incumbent legs import the UNMODIFIED modules from there (cpu_baseline.kind = "reference"); without the directory they fall
back to the port in oracle/torch_port.py (kind = "port").  Test / baseline infrastructure only: the product package never
imports anything under oracle/.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("S2VT_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("S2VTModel.py", "utils.py", "attention_baseline.py")


def build(quiet: bool = False) -> bool:
    if not all(os.path.exists(os.path.join(SRC, f)) for f in FILES):
        if not quiet:
            print("reference sources not found under %s: oracle/_ref left as it is" % SRC)
        return os.path.exists(os.path.join(DST, FILES[0]))
    os.makedirs(DST, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        lines.append("%s  %s" % (hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest(), f))
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if not quiet:
        print("oracle/_ref:", ", ".join(FILES))
    return True


def import_reference():
    """(S2VT, MaskCriterion, Att_Baseline) classes of the unmodified reference, or None when oracle/_ref is absent."""
    if not os.path.exists(os.path.join(DST, FILES[0])):
        return None
    sys.dont_write_bytecode = True
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from S2VTModel import S2VT                      # noqa: E402  (reference)
    from utils import MaskCriterion                 # noqa: E402  (reference)
    from attention_baseline import Att_Baseline     # noqa: E402  (reference)
    return S2VT, MaskCriterion, Att_Baseline


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
