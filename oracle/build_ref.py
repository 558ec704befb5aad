"""ORACLE tooling — recipe that stages the UNMODIFIED reference under oracle/_ref/ (git-ignored,
NOT gpurun-ignored: it travels to the GPU box the way built .so files do).

The reference is pure Python, so "building" it is copying the files of the hot path byte for byte
from /root/reference/src (read-only) — model.py, loss.py, generate.py and what they import
(configs.py, utils.py, MyDataset.py, config/model_config.json) — plus a SHA-256 manifest.
Nothing under oracle/_ref/ is committed or shipped with the product: it is the checker and the
`bench.py --impl reference` / `cpu_baseline` arm (kind "reference"), exactly like a compiled
oracle/_ref binary would be for a C reference. `__graft_entry__.build()` runs this when
/root/reference is present; on the GPU box the staged copy is used as it arrived.

    python oracle/build_ref.py            # stage / refresh
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref", "src")
FILES = ["model.py", "loss.py", "generate.py", "configs.py", "utils.py", "MyDataset.py",
         os.path.join("config", "model_config.json")]


def build(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref/src holds a complete copy (fresh or already there)."""
    if not os.path.isfile(os.path.join(SRC, "model.py")):
        return os.path.isfile(os.path.join(DST, "model.py"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print("staged", len(FILES), "reference files under", DST)
    return True


if __name__ == "__main__":
    build(verbose=True)
