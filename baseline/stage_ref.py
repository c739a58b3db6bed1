"""Stage the UNMODIFIED reference Python of the hot path under baseline/_ref/ so it travels to the GPU box.

baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so the
files ride along with the snapshot exactly like the in-tree .so files do.  What is staged: nerf/internal/*.py and
nerf/gridencoder/{__init__,grid,backend}.py, byte for byte, plus a MANIFEST.json with their sha256 so a test can assert
that what ran on the GPU box is what lies under /root/reference.  Used by
  * tests/test_gpu_reference_modules.py  - the reference's own grid.py / Model / render_image on the drop-in kernels;
  * bench.py (`gpu_reference` key, `--impl reference`) - the reference timed on the same box.
`__graft_entry__.build()` calls stage() whenever /root/reference is present; on the GPU box it is a no-op."""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/nerf"
DST = os.path.join(HERE, "_ref", "nerf")
GROUPS = {"internal": None, "gridencoder": ("__init__.py", "grid.py", "backend.py")}


def staged_root():
    """Directory to put on sys.path so `import internal.models` / `import gridencoder` resolve to the reference:
    the live tree when present (build container), else the staged copy (GPU box), else None."""
    if os.path.isdir(os.path.join(SRC, "internal")):
        return SRC
    if os.path.isdir(os.path.join(DST, "internal")):
        return DST
    return None


def stage(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "internal")):
        return DST if os.path.isdir(os.path.join(DST, "internal")) else None
    manifest = {}
    for sub, names in GROUPS.items():
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        for f in sorted(names or (x for x in os.listdir(os.path.join(SRC, sub)) if x.endswith(".py"))):
            s, d = os.path.join(SRC, sub, f), os.path.join(DST, sub, f)
            data = open(s, "rb").read()
            manifest[f"{sub}/{f}"] = hashlib.sha256(data).hexdigest()
            if not os.path.exists(d) or open(d, "rb").read() != data:
                shutil.copyfile(s, d)
    json.dump(manifest, open(os.path.join(DST, "..", "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"[baseline] staged {len(manifest)} reference files under {DST}")
    return DST


if __name__ == "__main__":
    print(stage(verbose=True))
