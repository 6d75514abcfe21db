"""The third-generation small-level kernel (csrc/kernels_coarse.cu, k_coarse_local<N0>) executed WITHOUT a GPU: its
source is compiled by g++ against tests/cpp/emu/host_emulation.h (one OS thread per CUDA thread, counting barriers that
abort on a missing or surplus arrival) and every top size / cycle form is compared bit for bit with the oracle.  This
pins the template recursion over the level sizes, the per-level thread groups with their named barriers, the buffer
parity of the warps that sit levels out, and the arithmetic order on CPU; the GPU suite
(test_gpu_parity.py::test_small_level_kernel_generations_agree_bitwise and every cycle test, which run it by default)
then checks the compiled kernel."""
import os
import subprocess

import cpu_checkers as cc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "parallel-geometric-multigrid-for-poisson-problem_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_coarse_kernel_emu")


def test_coarse_kernel_source_matches_oracle_under_host_emulation():
    cc.load("orc")  # builds oracle/liboracle.so if needed
    src = os.path.join(ROOT, "tests", "cpp", "test_coarse_kernel_emu.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-pthread", "-DPMG_HOST_EMULATION",
                    "-I" + os.path.join(ROOT, "tests", "cpp", "emu"), "-I" + os.path.join(PKG, "csrc"),
                    "-I" + os.path.join(ROOT, "include"), src, "-o", EXE, "-L" + cc.ORACLE_DIR, "-loracle",
                    "-Wl,-rpath," + cc.ORACLE_DIR], check=True)
    p = subprocess.run([EXE, "full"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all bit-identical" in p.stdout and "MISMATCH" not in p.stdout
    assert p.stdout.count("bit-identical") >= 12
