"""The drop-in, proven with the reference's own test suite: oracle/_ref/dropin/ holds a copy of the reference's vclust.py
with vclust_b200/integration.py's patch applied (made by oracle/make_dropin.py where /root/reference exists; build
output, it travels to the GPU box with the reference binaries), next to the reference's unmodified test.py and example/.
The prefilter / align / workflow tests of test.py (test.py:336-589) must pass
  * with the GPU library in place (the handlers call libvclust_b200.so), and
  * with it disabled (same wrapper, the subprocess path to kmer-db / lz-ani: BASELINE config c1, plumbing only)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
DROPIN = ROOT / "oracle" / "_ref" / "dropin"
SELECT = "prefilter or align"


def _run(env_extra, deselect=()):
    env = dict(os.environ)
    env.update(env_extra)
    env["PYTHONPATH"] = str(DROPIN) + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "pytest", "test.py", "-q", "-x", "-p", "no:cacheprovider", "-k", SELECT]
    for d in deselect:
        cmd += ["--deselect", d]
    return subprocess.run(cmd, cwd=DROPIN, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def _need_dropin():
    if not (DROPIN / "vclust.py").exists():
        pytest.skip("oracle/_ref/dropin not built (needs /root/reference at build time)")


def test_patch_applies_only_to_the_known_vclust():
    from vclust_b200 import integration
    with pytest.raises(ValueError):
        integration.apply("print('not vclust')\n")
    if (DROPIN / "vclust.py").exists():
        txt = (DROPIN / "vclust.py").read_text()
        assert txt.count("_b200_prefilter(args, logger)") == 2 and txt.count("_b200_align(args, logger)") == 2
        with pytest.raises(ValueError):
            integration.apply(txt)                       # already patched


def test_reference_suite_through_the_subprocess_fallback():
    """No GPU library: the patched wrapper must behave exactly like the reference (kmer-db / lz-ani subprocesses)."""
    _need_dropin()
    # (--batch-size needs the reference's mfasta-tool on this path, which oracle/build_ref.sh does not build; the GPU path
    # below runs those cases too)
    p = _run({"VCLUST_B200_DISABLE": "1"}, deselect=["test.py::test_prefilter_default[input2-params2]",
                                                      "test.py::test_workflow_prefilter_align[input2-params2]"])
    assert p.returncode == 0, p.stdout[-3000:]
    assert " passed" in p.stdout and "failed" not in p.stdout


@pytest.mark.gpu
def test_reference_suite_through_the_gpu_library(tmp_path):
    """The reference's prefilter / align / workflow tests with handle_prefilter / handle_align calling libvclust_b200.so."""
    _need_dropin()
    lib = ROOT / "vclust_b200" / "libvclust_b200.so"
    # a marker file proves that the library path ran: vb_* calls are made in the wrapper process, which logs "libvclust_b200"
    p = _run({"VCLUST_B200_LIB": str(lib)})
    assert p.returncode == 0, p.stdout[-3000:]
    assert " passed" in p.stdout and "failed" not in p.stdout
    out = tmp_path / "f.txt"
    q = subprocess.run([sys.executable, str(DROPIN / "vclust.py"), "prefilter", "-i", str(DROPIN / "example" / "multifasta.fna"), "-o", str(out)],
                       env=dict(os.environ, VCLUST_B200_LIB=str(lib)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert q.returncode == 0 and "Running: libvclust_b200 prefilter" in q.stderr and "kmer-db" not in q.stderr
    assert out.read_bytes() == (DROPIN / "example" / "output" / "fltr.txt").read_bytes()
