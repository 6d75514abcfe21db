"""The fused streaming kernels (csrc/kernels_fused.cu: Pass A / Pass B, with their launch and geometry code) executed
WITHOUT a GPU: the kernel source is compiled by g++ against tests/cpp/emu/host_emulation.h (32 fibers per warp, warp
shuffles as collectives, cp.async / shared memory emulated) and every output is compared bit for bit with the CPU
oracle -- whole levels (all sweep counts, variants, chunk geometries) and row slabs with 2-4 ranks in one process
reading each other's halo rows in place, exactly as the multi-GPU path does over NVLink."""
import os
import subprocess

import cpu_checkers as cc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "parallel-geometric-multigrid-for-poisson-problem_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_fused_kernel_emu")


def test_fused_kernel_source_matches_oracle_under_host_emulation():
    cc.load("orc")  # builds oracle/liboracle.so if needed
    src = os.path.join(ROOT, "tests", "cpp", "test_fused_kernel_emu.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-pthread", "-DPMG_HOST_EMULATION",
                    "-I" + os.path.join(ROOT, "tests", "cpp", "emu"), "-I" + os.path.join(PKG, "csrc"),
                    "-I" + os.path.join(ROOT, "include"), src, "-o", EXE, "-L" + cc.ORACLE_DIR, "-loracle",
                    "-Wl,-rpath," + cc.ORACLE_DIR], check=True)
    p = subprocess.run([EXE, "full"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-2000:]
    assert "all bit-identical" in p.stdout and "MISMATCH" not in p.stdout
    assert p.stdout.count(": ok") >= 20
