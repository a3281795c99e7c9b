"""Recipe for `oracle/_ref/`: an installed copy of the UNMODIFIED reference package (ravest v0.4.0).

TEST / BASELINE INFRASTRUCTURE ONLY.  `oracle/_ref/` is git-ignored (never in history) but NOT gpurun-ignored,
so the reference's own `LogPosterior.log_probability` can be timed on the GPU box's host cores
(`bench.py --impl reference`, `cpu_baseline.kind = "reference"`) and checked live against the CUDA path.

ravest is a pure-Python package with a poetry-core build backend.  `pip install --no-index --no-build-isolation
--no-deps --target oracle/_ref <copy of /root/reference>` is tried first; poetry-core is not in this image's
wheelhouse, so the recipe falls back to what that install would have produced for a pure-Python project: the
package directory `src/ravest/*.py`, byte for byte.  Nothing is edited; PROVENANCE.txt records file hashes.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def make(verbose: bool = False) -> str | None:
    """Returns the path of the installed copy, or None when /root/reference is absent (GPU box: use what travelled)."""
    pkg_src = os.path.join(SRC, "src", "ravest")
    if not os.path.isdir(pkg_src):
        return DST if os.path.isfile(os.path.join(DST, "ravest", "fit.py")) else None
    os.makedirs(DST, exist_ok=True)
    how = "copy of src/ravest (pure-Python package; pip --target failed: no poetry-core backend offline)"
    installed = False
    if os.environ.get("RVLP_REF_TRY_PIP"):
        with tempfile.TemporaryDirectory() as tmp:
            work = os.path.join(tmp, "ravest_src")
            shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git"))
            r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                                "--find-links", "/opt/wheelhouse", "--target", DST, "--upgrade", work],
                               capture_output=True, text=True)
            installed = r.returncode == 0
            if installed:
                how = "pip install --no-index --no-build-isolation --no-deps --target oracle/_ref"
    if not installed:
        dst_pkg = os.path.join(DST, "ravest")
        if os.path.isdir(dst_pkg):
            shutil.rmtree(dst_pkg)
        shutil.copytree(pkg_src, dst_pkg, ignore=shutil.ignore_patterns("__pycache__"))
    lines = [f"ravest reference package, unmodified; made by oracle/make_ref.py: {how}"]
    for root, _, files in sorted(os.walk(os.path.join(DST, "ravest"))):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(root, f)
                lines.append(f"{hashlib.sha256(open(p, 'rb').read()).hexdigest()}  {os.path.relpath(p, DST)}")
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print(f"oracle/_ref: {how}")
    return DST


if __name__ == "__main__":
    print(make(verbose=True))
