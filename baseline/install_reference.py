#!/usr/bin/env python
"""Place the UNMODIFIED reference (donydchen/matchnerf) under baseline/_ref/ so that it travels to the GPU box.

The task's recipe (`pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference`) fails here with "Directory '/root/reference' is not installable. Neither 'setup.py' nor 'pyproject.toml'
found": the reference is a plain script tree, not a package.  "Installing" it therefore means copying the tree verbatim;
baseline/_ref/ is git-ignored (never part of this repo's history, never imported by the product package) but not
gpurun-ignored.  A manifest with the sha256 of every file is written next to it so that bench.py can state which
reference it timed.  Only bench.py's reference legs and the dev-container drop-in tests import it, through
oracle/reference_shim.py.

    python baseline/install_reference.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
KEEP = ("models", "misc", "configs", "datasets", "docs", "options.py", "coach.py", "test.py", "train.py", "LICENSE", "README.md",
        "requirements.txt")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("MATCHNERF_REFERENCE", "/root/reference"))
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(args.src, "models")):
        raise SystemExit(f"{args.src} does not look like the reference checkout")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for name in KEEP:
        s = os.path.join(args.src, name)
        if not os.path.exists(s):
            continue
        d = os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    for root, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=args.src, files=manifest), f, indent=1, sort_keys=True)
    print(f"copied {len(manifest)} files of the unmodified reference to {DST}")


if __name__ == "__main__":
    main()
