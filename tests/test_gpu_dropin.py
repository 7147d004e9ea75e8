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


def build_caller_test(outdir):
    exe = os.path.join(str(outdir), "caller_test")
    pkg = os.path.join(ROOT, "ploidyfrost_b200")
    subprocess.run(["g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "dropin", "caller_test.cpp"),
                    "-o", exe, "-L", pkg, "-lpfgpu", "-Wl,-rpath," + pkg, "-pthread"], check=True)
    return exe


def test_dropin_adapters_compile(tmp_path):
    """CPU: the adapter header is valid C++ against the C ABI and links (no compute call is made)."""
    from ploidyfrost_b200 import build
    build.build_library()
    assert os.path.exists(build_dropin_test(tmp_path))
    assert os.path.exists(build_caller_test(tmp_path))


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


@pytest.mark.gpu
def test_batched_caller_writes_the_reference_files(tmp_path):
    """include/pf_caller.hpp (C++, over the C ABI) fed with the bubbles of tests/golden/e2e writes P_alignseq.txt and the coverage /
    frequency files of the unmodified reference's `-t 1` run, byte for byte."""
    from ploidyfrost_b200 import build
    build.build_library()
    exe = build_caller_test(tmp_path)
    fx = os.path.join(ROOT, "tests", "golden", "e2e")
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([exe, fx, str(out), "2", "1000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "called 276" in r.stdout, r.stdout
    for name in ["P_alignseq.txt"] + [f"P_{a}{b}.txt" for a in ("bi", "tri", "tetra", "penta") for b in ("cov", "fre")]:
        want = open(os.path.join(fx, name), "rb").read()
        got = open(out / name, "rb").read()
        assert got == want, f"{name} differs from the reference's file"
