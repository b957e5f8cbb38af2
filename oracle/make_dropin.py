"""Test infrastructure: a working copy of the reference's entry point with the GPU patch applied.

Copies /root/reference/{vclust.py, test.py, example/} to oracle/_ref/dropin/ (build output, git-ignored, travels to the GPU
box with the reference binaries), applies vclust_b200.integration.apply() to vclust.py and fills bin/ with the reference
binaries oracle/build_ref.sh built.  tests/test_dropin.py then runs the reference's OWN test.py there.  Nothing of this is
committed; without /root/reference the function does nothing."""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REF = Path("/root/reference")
OUT = HERE / "_ref" / "dropin"


def make(force: bool = False) -> Path | None:
    if not (REF / "vclust.py").exists() or not (HERE / "_ref" / "kmer-db").exists():
        return OUT if (OUT / "vclust.py").exists() else None
    sys.path.insert(0, str(ROOT))
    from vclust_b200 import integration
    patched = integration.apply((REF / "vclust.py").read_text())
    if not force and (OUT / "vclust.py").exists() and (OUT / "vclust.py").read_text() == patched and (OUT / "example" / "multifasta.fna").exists():
        return OUT
    if OUT.exists():
        shutil.rmtree(OUT)
    (OUT / "bin").mkdir(parents=True)
    (OUT / "vclust.py").write_text(patched)
    (OUT / "vclust.py").chmod(0o755)
    shutil.copy(REF / "test.py", OUT / "test.py")
    shutil.copytree(REF / "example", OUT / "example")
    for b in ("kmer-db", "lz-ani"):
        shutil.copy(HERE / "_ref" / b, OUT / "bin" / b)
    return OUT


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
