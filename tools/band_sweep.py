#!/usr/bin/env python
"""Build (here) / time (GPU box) variants of the percentile-band kernel: python tools/band_sweep.py build|run"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "build_variants")
VARIANTS = {"band_u2_mb4": ["-DRVLP_BAND_UNROLL=2", "-DRVLP_BAND_MINB1=4"],
            "band_u4_mb3": ["-DRVLP_BAND_UNROLL=4", "-DRVLP_BAND_MINB1=3"],
            "band_u8_mb3": ["-DRVLP_BAND_UNROLL=8", "-DRVLP_BAND_MINB1=3"],
            "band_u8_mb2": ["-DRVLP_BAND_UNROLL=8", "-DRVLP_BAND_MINB1=2"],
            "band_u16_mb2": ["-DRVLP_BAND_UNROLL=16", "-DRVLP_BAND_MINB1=2"]}
sys.path.insert(0, ROOT)
if sys.argv[1] == "build":
    from ravest_b200 import _lib
    os.makedirs(VDIR, exist_ok=True)
    for name, flags in VARIANTS.items():
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = ["nvcc"] + _lib.NVCC_FLAGS + flags + ["-o", out, os.path.join(_lib.CSRC, "rvlp_capi.cu")]
        subprocess.run(cmd, check=True)
        print("built", out)
else:
    for name in VARIANTS:
        env = dict(os.environ, RVLP_LIB=os.path.join(VDIR, f"lib_{name}.so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "rows_time.py"), "bands"], env=env, capture_output=True, text=True)
        print(name, "\n".join(l for l in r.stdout.splitlines() if "percentile" in l), r.stderr[-300:] if r.returncode else "")
