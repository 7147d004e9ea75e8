"""-m gpu: the C++ look-alike adapters (include/pf_dropin.hpp: CKMCFile / SeqAlign shapes over the C ABI),
compiled with g++ against libpfgpu.so and run as a maintainer's code would."""
import os
import subprocess

import pytest

from tests import gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_dropin_test(outdir):
    exe = os.path.join(str(outdir), "dropin_test")
    pkg = os.path.join(ROOT, "ploidyfrost_b200")
    subprocess.run(["g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "dropin", "dropin_test.cpp"),
                    "-o", exe, "-L", pkg, "-lpfgpu", "-Wl,-rpath," + pkg], check=True)
    return exe


def test_dropin_adapters_compile(tmp_path):
    """CPU: the adapter header is valid C++ against the C ABI and links (no compute call is made)."""
    from ploidyfrost_b200 import build
    build.build_library()
    assert os.path.exists(build_dropin_test(tmp_path))


@pytest.mark.gpu
def test_dropin_adapters_run(tmp_path):
    from ploidyfrost_b200 import build
    build.build_library()
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=9, genome_len=5000, k=25, version=0x200, p=5)
    exe = build_dropin_test(tmp_path)
    absent = "ACGTACGTACGTACGTACGTACGTA"
    assert absent not in g
    r = subprocess.run([exe, prefix, g[100:125], absent], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
