"""Build helper for the C++ host front end (fans_b200/host/main.cpp -> tests/_build/FANS_gpu)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "FANS_gpu")


def build():
    src = os.path.join(ROOT, "fans_b200", "host", "main.cpp")
    lib = os.path.join(ROOT, "fans_b200", "lib")
    deps = [src] + [os.path.join(ROOT, "fans_b200", "host", f) for f in os.listdir(os.path.join(ROOT, "fans_b200", "host"))]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE, "-L", lib, "-lfans_gpu", "-lz",
                    "-Wl,-rpath," + lib], check=True)
    return EXE
