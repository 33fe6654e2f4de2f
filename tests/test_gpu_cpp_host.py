"""GPU (-m gpu): builds tests/cpp/fc_tests.cpp against the C++ host mirror + libkolibrie_b200.so and runs it.
The CPU half (-m "not gpu") only checks that the host mirror compiles and links against the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "fc_tests.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "fc_tests")
LIBDIR = os.path.join(ROOT, "kolibrie_b200")


def build():
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-o", BIN, SRC, f"-L{LIBDIR}", "-lkolibrie_b200", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return BIN


def test_cpp_host_mirror_compiles_and_links():
    build()


@pytest.mark.gpu
def test_cpp_fc_tests_pass_on_device():
    exe = build()
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "0 failed" in r.stdout
