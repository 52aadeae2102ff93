"""lightzero_b200/csrc/lz_exact_math.h restates glibc's expf; check it on the host against libm (the
function the reference tree calls at cnode.cpp:129).  The full sweep (every float in [-104, +0],
1.12e9 inputs, 0 mismatches) takes ~20 s: `check_expf 1`.  CI runs a strided sweep + the one input
where the fused and unfused range reductions differ."""
import os
import subprocess

from conftest import ROOT


def test_expf_matches_libm(tmp_path):
    exe = str(tmp_path / "check_expf")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tests", "csrc", "check_expf.c"), "-lm"])
    stride = "1" if os.environ.get("LZ_EXHAUSTIVE") else "61"
    out = subprocess.run([exe, stride], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0" in out.stdout
