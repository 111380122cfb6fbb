"""Stage the UNMODIFIED reference for the CPU arm of bench.py (`--impl reference`) and for §8(d) timings.

Test / measurement infrastructure, not part of the product path.  The reference is a script directory without packaging
(no setup.py / pyproject), so "installing" it means copying the files the hot path imports -- byte for byte --
from /root/reference into baseline/_ref/, which is git-ignored (never in history) but travels to the GPU box with gpurun:
    centerface.py, centerface_ext.py, eval_widerface.py, demo.py, model/, utils/, weight/model_epoch_100.pt, imgs/*.jpg
Runs only where /root/reference exists (the build container); __graft_entry__.build() calls it.

    python oracle/make_ref.py
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["centerface.py", "centerface_ext.py", "eval_widerface.py", "demo.py", "weight/model_epoch_100.pt"]
DIRS = ["model", "utils", "imgs"]


def stage():
    if not os.path.isdir(REF):
        return False
    manifest = {}
    for rel in list(FILES) + [os.path.join(d, f) for d in DIRS for f in sorted(os.listdir(os.path.join(REF, d)))
                              if os.path.isfile(os.path.join(REF, d, f))]:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else "no /root/reference here: nothing staged", DST)
    sys.exit(0)
