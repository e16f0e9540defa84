#!/usr/bin/env python
"""Stage the reference's OWN Python call sites into the git-ignored baseline/_ref/ so that the GPU box (which has no
/root/reference) can run them UNMODIFIED on top of this repo's `awq_inference_engine` plugin:

    awq/quantize/qmodule.py            WQLinear.forward                      (qmodule.py:201-224)
    tinychat/modules/fused_mlp.py      QuantLlamaMLP.our_llama_mlp           (fused_mlp.py:36-83)
    tinychat/modules/fused_attn.py     make_quant_attn's QKV concatenation   (fused_attn.py:566-594)
  + whatever those modules import from the reference's own packages.

Nothing here enters the repository history: baseline/_ref/ is listed in .gitignore (it still travels to the GPU box
with gpurun).  Only *.py files of the `awq` and `tinychat` packages are copied (no kernels, no assets); the copy is
byte-for-byte, and tests/test_reference_callsites_gpu.py checks the staged files against the manifest written here.

    python scripts/stage_reference.py          (build container only; a no-op when /root/reference is absent)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("AWQ_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ["awq", "tinychat"]
SKIP_DIRS = {"kernels", "serve", "__pycache__", "scripts"}


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "awq", "quantize"))


def stage() -> dict:
    manifest = {}
    for pkg in PACKAGES:
        src_root = os.path.join(REF, pkg)
        for d, dirs, files in os.walk(src_root):
            dirs[:] = [x for x in dirs if x not in SKIP_DIRS]
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(d, f)
                rel = os.path.relpath(src, REF)
                dst = os.path.join(OUT, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "files": manifest}, fh, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    if not available():
        print("reference checkout not found at", REF, "- nothing staged")
        sys.exit(0)
    m = stage()
    print("staged", len(m), "reference python files into", OUT)
