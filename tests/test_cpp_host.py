"""The compiled host side above the C ABI: include/imagepipe_b200.hpp (C++17 mirror of the reference's Pipeline /
ImageOp surface) driven by tests/cpp/host_demo.cpp.  CPU: it compiles with g++, links libipb200.so and its host-only
calls work without a GPU.  GPU: the same binary runs a frame through Pipeline::output_8bit and through the eight
ImageOp::run calls, and both equal the oracle's bytes."""
import os
import subprocess

import numpy as np
import pytest

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_demo(tmp_path_factory, ip):
    out = str(tmp_path_factory.mktemp("cpp") / "host_demo")
    libdir = os.path.join(ROOT, "imagepipe_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_demo.cpp"), "-o", out, "-L", libdir, "-lipb200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_cpp_mirror_compiles_links_and_negotiates_sizes(host_demo):
    r = subprocess.run([host_demo, "sizes"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,maxwidth", [(640, 360, 0), (800, 600, 200)])
def test_cpp_pipeline_and_ops_match_the_oracle(host_demo, orc, tmp_path, w, h, maxwidth):
    out = str(tmp_path / "out.bin")
    r = subprocess.run([host_demo, "run", str(w), str(h), str(common.SEED), out, str(maxwidth)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    ow, oh = (int(v) for v in lines[0].split())
    data = common.synth_cfa(w, h)
    st = {"maxwidth": maxwidth} if maxwidth else None
    want = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", common.raw_params(), st))
    assert (ow, oh) == (want.shape[1], want.shape[0])
    got = np.fromfile(out, np.uint8).reshape(oh, ow, 3)
    common.assert_bit_exact(got, want, "C++ Pipeline::output_8bit")
    ops = np.fromfile(out + ".ops", np.uint8).reshape(oh, ow, 3)
    common.assert_bit_exact(ops, want, "C++ ImageOp::run chain")
