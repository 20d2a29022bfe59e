"""Recipe for oracle/_ref: the reference's OWN hot-path modules, unmodified, importable without the repo.

    python oracle/build_ref.py            # needs /root/reference (this container only)

The reference (HumaticsLAB/SEAM-Match-RCNN) is a flat script repository without setup.py, so it cannot be
pip-installed into baseline/_ref.  Its hot path lives in two pure-Python files, models/nlb.py and
models/match_head.py (TemporalAggregationNLB, MatchPredictor, NONLocalBlock1D); this script copies those two
files byte for byte from /root/reference into oracle/_ref/seam_ref/ (git-ignored: reference sources never
enter the history; NOT gpurun-ignored: the directory travels to the GPU box like the built .so) and adds
  * an empty ``pycocotools`` shim: match_head.py imports ``pycocotools.mask`` for filter_proposals (training
    only, models/match_head.py:451); pycocotools is not installed here and the hot path never calls it;
  * ``MANIFEST.json`` with the sha256 of each copied file.
oracle/_ref is TEST / BASELINE infrastructure: only tests/, bench.py --impl reference and bench.py's
cpu_baseline leg import it (kind = "reference"); when it is absent they fall back to the oracle port.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SEAM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["models/nlb.py", "models/match_head.py"]

SHIM = '''"""Empty stand-in: models/match_head.py imports pycocotools.mask at module level but only
filter_proposals (training) uses it."""
'''


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(REF_ROOT):
        if verbose:
            print(f"[build_ref] {REF_ROOT} not present: keeping whatever oracle/_ref holds")
        return os.path.isdir(os.path.join(OUT, "seam_ref"))
    pkg = os.path.join(OUT, "seam_ref")
    os.makedirs(pkg, exist_ok=True)
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(pkg, os.path.basename(rel))
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(pkg, "__init__.py"), "w") as f:
        f.write('"""The reference\'s models/nlb.py and models/match_head.py, copied unmodified by oracle/build_ref.py."""\n')
    shim = os.path.join(OUT, "pycocotools")
    os.makedirs(shim, exist_ok=True)
    for name in ("__init__.py", "mask.py"):
        with open(os.path.join(shim, name), "w") as f:
            f.write(SHIM)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "files": manifest}, f, indent=1)
    if verbose:
        print(f"[build_ref] {len(FILES)} reference files -> {pkg}")
    return True


def load():
    """Import the vendored reference modules; returns the match_head module or None when oracle/_ref is absent."""
    pkg = os.path.join(OUT, "seam_ref")
    if not os.path.exists(os.path.join(pkg, "match_head.py")):
        return None
    if OUT not in sys.path:
        sys.path.append(OUT)            # appended: a real pycocotools, if installed, wins over the shim
    import importlib
    return importlib.import_module("seam_ref.match_head")


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok else 1)
