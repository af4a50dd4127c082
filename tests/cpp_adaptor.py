"""Helpers for the C++ adaptor test program (tests/cpp/adaptor_test.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adaptor_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "adaptor_test")


def build():
    """g++ the adaptor test against libchinium_fock.so (rpath relative to the binary, so it travels to the GPU box)."""
    deps = [SRC, os.path.join(ROOT, "chinium_b200", "cpp", "Int4C2E_b200.hpp"), os.path.join(ROOT, "include", "chinium_fock.h")]
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(d) for d in deps):
        return EXE
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", EXE, SRC, "-L" + os.path.join(ROOT, "chinium_b200"),
                           "-lchinium_fock", "-Wl,-rpath,$ORIGIN/../../chinium_b200"])
    return EXE


def write_input(path, fb, D):
    with open(path, "w") as f:
        f.write("%d %d\n" % (fb.nshell, fb.nbf))
        for s in range(fb.nshell):
            x = fb.center_xyz[s]
            f.write("%d %d %d %.17g %.17g %.17g\n" % (fb.type[s], fb.nprim[s], fb.shell2atom[s], x[0], x[1], x[2]))
            for k in range(fb.nprim[s]):
                o = fb.prim_offset[s] + k
                f.write("%.17g %.17g\n" % (fb.exps[o], fb.coefs_normalized[o]))
        for v in D.reshape(-1, order="F"):
            f.write("%.17g\n" % v)
