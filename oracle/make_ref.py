"""Recipe for oracle/_ref/: the UNMODIFIED reference files of the hot path, staged where they can travel to the GPU box.

TEST INFRASTRUCTURE, not product.  Nothing under ullsam_b200/ imports this; only tests/, __graft_entry__.smoke()/build()
and bench.py's reference / cpu_baseline legs may use what it produces.

The reference (ieellee/uLLSAM) is a directory of Python scripts with no build system, so "building" it means staging
its own source files, byte for byte, from where they lie under /root/reference (read-only, build container only) into
oracle/_ref/ — git-ignored (never in the history), NOT gpurun-ignored (it travels to the GPU box with the snapshot,
like the built .so files).  With it present

  * `bench.py --impl reference` and the `cpu_baseline` leg time the reference's own modules (kind "reference") instead
    of the oracle port (kind "port");
  * the GPU tests can swap the CUDA path into the REAL reference classes (build_sam._build_sam, InternVLSAMModel) and
    compare against them in the same process.

    python -m oracle.make_ref            # stages the files, prints what it copied

Files staged: build_sam.py and modeling/*.py (image_encoder, mask_decoder, transformer, common,
prompt_encoder, sam, and the InternVL bridge: modeling_internvl_sam, modeling_internlm2, configuration_*, conversation).
A MANIFEST.json with the SHA-256 of every file is written next to them so a stale copy is detectable.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("ULLSAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

FILES = ["build_sam.py"] + [os.path.join("modeling", f) for f in (
    "__init__.py", "common.py", "image_encoder.py", "mask_decoder.py", "transformer.py", "prompt_encoder.py", "sam.py",
    "modeling_internvl_sam.py", "modeling_internlm2.py", "configuration_intern_vit.py", "configuration_internlm2.py",
    "configuration_internvl_chat.py", "conversation.py")]


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "MANIFEST.json"))


def stage(verbose: bool = True) -> bool:
    """Copy the reference's files into oracle/_ref/.  Returns False (and does nothing) when /root/reference is absent."""
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/_ref: {REF} not present (GPU box?) — using whatever was staged before: available={available()}")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: staged {len(manifest)} unmodified reference files from {REF}")
    return True


def import_reference():
    """Import the staged reference: returns (modeling package, build_sam module).  Raises if it has not been staged.

    The reference's top-level package is called `modeling` and build_sam.py does `from modeling import ...`, so
    oracle/_ref goes on sys.path (first, so nothing else named `modeling` shadows it)."""
    if not available():
        raise RuntimeError("oracle/_ref has not been staged: run `python -m oracle.make_ref` where /root/reference exists")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import importlib

    modeling = importlib.import_module("modeling")
    if not os.path.abspath(modeling.__file__).startswith(DST):
        raise RuntimeError(f"`modeling` resolved to {modeling.__file__}, not to oracle/_ref")
    build_sam = importlib.import_module("build_sam")
    return modeling, build_sam


if __name__ == "__main__":
    ok = stage()
    sys.exit(0 if ok or available() else 1)
