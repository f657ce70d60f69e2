"""Recipe for oracle/_ref: the reference's OWN hot-path module files, staged for the CPU baseline.  TEST INFRASTRUCTURE.

    python oracle/build_ref.py          (dev container only: needs /root/reference)

The reference is pure Python (no build system, no setup.py); its hot path is two module files that import only
detectron2.layers / detectron2.utils.registry / fvcore.nn.weight_init.  "Building" it therefore means staging those two files,
unmodified and byte-for-byte, where the GPU box can load them: oracle/_ref/ is git-ignored (never part of the history) but
NOT gpurun-ignored, so it travels like the built .so files do.  oracle/ref_runner.py loads them through oracle/_ref_stubs
(40-line stand-ins for the detectron2 / fvcore names) and runs the stage-1 step of
afigan/engine/stage1_trainer.py:334-433 on them: bench.py's `cpu_baseline` (kind "reference") and `--impl reference`.
__graft_entry__.build() runs this when /root/reference is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("AFIGAN_REFERENCE_ROOT", "/root/reference")
FILES = ("afigan/modeling/feat_interpol/generator_rdb.py", "afigan/modeling/feat_interpol/feature_patch_discriminator.py")
OUT = os.path.join(HERE, "_ref")


def build(verbose=True):
    if not os.path.isdir(REF_ROOT):
        if verbose:
            print(f"oracle/build_ref.py: {REF_ROOT} is absent; keeping the staged files" if os.path.isdir(OUT) else
                  f"oracle/build_ref.py: {REF_ROOT} is absent and nothing is staged (bench.py falls back to the oracle port)")
        return os.path.isdir(OUT)
    os.makedirs(OUT, exist_ok=True)
    manifest = {"source_root": REF_ROOT, "files": {}}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, os.path.basename(rel))
        shutil.copyfile(src, dst)
        manifest["files"][os.path.basename(rel)] = {"from": rel, "sha256": hashlib.sha256(open(src, "rb").read()).hexdigest()}
    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("staged", ", ".join(manifest["files"]), "->", OUT)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
