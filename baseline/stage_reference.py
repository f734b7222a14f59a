"""Stage the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, ships with gpurun).

    python baseline/stage_reference.py            # copies /root/reference/{utils,trainer}/*.py byte for byte

The reference (yl-jiang/YOLOSeries) has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference`
has nothing to build; it is importable only from a checkout.  `bench.py --impl reference` and the mode-B parity tests
 run on the GPU box where /root/reference does not exist, so the two packages the path lives in --
`utils/` (nms.py, bbox_tools.py, anchor.py, weighted_fusion_bbox.py, mAP.py and the modules utils/__init__.py
star-imports) and `trainer/` (eval_*.py) -- are copied verbatim into baseline/_ref/.  Nothing under baseline/_ref is
tracked by git and no file is edited: a SHA-256 manifest (baseline/_ref/MANIFEST.json) records what was copied so
that a run can prove it executed the unmodified sources.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("YSB_REFERENCE_ROOT", "/root/reference")
PACKAGES = ("utils", "trainer")


def stage(src=SRC, dest=DEST, verbose=False):
    """Returns the manifest {relative path: sha256}; raises if the checkout is missing."""
    if not os.path.isdir(os.path.join(src, "trainer")):
        raise RuntimeError(f"reference checkout not found at {src}")
    manifest = {}
    for pkg in PACKAGES:
        out_dir = os.path.join(dest, pkg)
        os.makedirs(out_dir, exist_ok=True)
        for name in sorted(os.listdir(os.path.join(src, pkg))):
            if not name.endswith(".py"):
                continue
            s, d = os.path.join(src, pkg, name), os.path.join(out_dir, name)
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
            with open(d, "rb") as f:
                manifest[f"{pkg}/{name}"] = hashlib.sha256(f.read()).hexdigest()
            if verbose:
                print("staged", f"{pkg}/{name}")
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def staged() -> bool:
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


if __name__ == "__main__":
    m = stage(verbose="-v" in sys.argv)
    print(f"staged {len(m)} files under {DEST}")
