"""The hardware-decode binding against a SOFTWARE TEST DOUBLE of libnvcuvid (tests/mock_nvcuvid): on boxes whose GPU is
exposed for compute only (this round's: every real cuvid call answers CUDA_ERROR_NO_DEVICE) the hardware tests of
tests/test_nvdec_gpu.py skip; here they run in a child process that loads the double through GVL_NVCUVID_LIB, so the
binding's callback flow, surface mapping, sampling rule, NV12 -> RGB kernel, batch ring and pipeline integration execute
on the GPU.  The double shares csrc/cuvid_abi.h with the binding: it says nothing about the real driver's struct layouts."""
import os
import re
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_binding_end_to_end_against_the_software_double(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++ to build the test double")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    lib = str(tmp_path / "libmock_nvcuvid.so")
    build = subprocess.run(["g++", "-shared", "-fPIC", "-O2", "-std=c++17", f"-I{cuda}/include",
                            f"-I{ROOT}/gameplay_vision_llm_b200/csrc", f"{ROOT}/tests/mock_nvcuvid/mock_nvcuvid.cpp", "-o", lib,
                            f"-L{cuda}/lib64", "-lcudart"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    env = dict(os.environ, GVL_NVCUVID_LIB=lib)
    run = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_nvdec_gpu.py"), "-q", "-m", "gpu",
                          "-rs", "-s", "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=ROOT)
    print(run.stdout[-3000:])
    assert run.returncode == 0, run.stdout[-4000:] + run.stderr[-2000:]
    assert "NVDEC unusable" not in run.stdout, "the child process did not pick up the test double"
    m = re.search(r"(\d+) passed", run.stdout)
    assert m and int(m.group(1)) >= 10, run.stdout[-2000:]
