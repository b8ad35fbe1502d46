"""Recipe for oracle/_ref/: the UNMODIFIED reference, made importable on a box that has no
/root/reference (the GPU box).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

    python oracle/make_ref.py            (also called by __graft_entry__.build())

The reference is a flat directory of Python scripts with no build system, so "building" it means
placing its nine source files, byte for byte, next to three stub modules for imports this image
does not have (ipdb, tensorboardX, matplotlib: imported at module top in layers.py:12, utils.py:5,
18-19, main.py:13, 17 -- never on the timed path).  Outputs go ONLY into oracle/_ref/, which is
git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so it
travels to the GPU box like the built libedis.so.  MANIFEST.json records a sha256 per file.
Consumers: bench.py's `--impl reference` arm and `cpu_baseline` leg (oracle/ref_arm.py), tests.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
STUBS = os.path.join(os.path.dirname(HERE), "tests", "golden", "_stubs")


def make(ref="/root/reference", dest=DEST, quiet=False):
    """Returns True if oracle/_ref/ is usable afterwards (freshly made or already there)."""
    if not os.path.isdir(ref):
        return os.path.exists(os.path.join(dest, "MANIFEST.json"))
    os.makedirs(dest, exist_ok=True)
    manifest = {}
    for f in sorted(os.listdir(ref)):
        if not f.endswith(".py"):
            continue
        src, dst = os.path.join(ref, f), os.path.join(dest, f)
        shutil.copyfile(src, dst)
        manifest[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    sdst = os.path.join(dest, "_stubs")
    if os.path.isdir(sdst):
        shutil.rmtree(sdst)
    shutil.copytree(STUBS, sdst, ignore=shutil.ignore_patterns("__pycache__"))
    json.dump({"source": ref, "files": manifest}, open(os.path.join(dest, "MANIFEST.json"), "w"), indent=1,
              sort_keys=True)
    if not quiet:
        print("oracle/_ref: %d reference files + stubs" % len(manifest))
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
