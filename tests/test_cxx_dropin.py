"""A C++ caller written against the reference's headers/namespaces (tests/cxx_dropin.cpp) builds with plain g++
against include/ + libopengjk_b200.so (CPU check) and prints the README's documented outputs on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "opengjk-gpu_b200", "lib")


def _build(flag):
    out = os.path.join(ROOT, "tests", "_build", "cxx_dropin" + ("_f64" if flag else "_f32"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx_dropin.cpp"),
           "-L", LIBDIR, "-lopengjk_b200", f"-Wl,-rpath,{LIBDIR}", "-o", out]
    if flag:
        cmd.insert(1, flag)
    subprocess.check_call(cmd)
    return out


@pytest.mark.parametrize("flag", ["", "-DOGJK_USE_64BITS"])
def test_reference_style_caller_compiles_and_links(pkg, flag):
    pkg.load_library()
    assert os.path.exists(_build(flag))


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["", "0,0"])  # "0,0": the fan-out path of GJK::GPU::compute* (device 0 listed twice)
@pytest.mark.parametrize("flag", ["", "-DOGJK_USE_64BITS"])
def test_reference_style_caller_prints_readme_outputs(pkg, flag, devices):
    pkg.load_library()
    exe = _build(flag)
    env = dict(os.environ)
    env.pop("OGJK_DEVICES", None)
    if devices:
        env["OGJK_DEVICES"] = devices
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    text = out.stdout
    # reference README.md:111-115
    assert "Distance between bodies 3.653650" in text
    assert "Witnesses: (1.025173, 1.490318, 0.255463) and (-1.025173, -1.490318, -0.255463)" in text
    # reference README.md:131-137
    assert "Penetration depth: 1.500000" in text
    assert "Witness point on cube 1: (1.000000, 0.500000, 0.707107)" in text
    assert "Witness point on cube 2: (-0.500000, 0.500000, 0.707107)" in text
    assert "Contact normal (from cube 1 to cube 2): (1.000000, -0.000000, 0.000000)" in text
    assert "README-API depth 1.500000 w1 (1.000000, 0.500000, 0.707107)" in text
    assert "indexed depth 1.500000" in text
