"""Builds A/B variants of libsurfel_b200.so: the default objects with some translation units recompiled
under extra -D flags.   python profiles/build_variants.py NAME=file.cu:-DFLAG=1[,file2.cu:-DX=2] ...
Output: 2d-gaussian-splatting_b200/lib/variants/NAME.so (git-ignored; travels to the GPU box)."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "2d-gaussian-splatting_b200")
spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)


def main():
    b.build()
    vdir = os.path.join(PKG, "lib", "variants"); os.makedirs(vdir, exist_ok=True)
    odir = os.path.join(PKG, "build", "variants"); os.makedirs(odir, exist_ok=True)
    procs = []
    for spec_ in sys.argv[1:]:
        name, rest = spec_.split("=", 1)
        over = {}
        for item in rest.split(","):
            f, flags = item.split(":", 1)
            over.setdefault(f, []).extend(flags.split())
        objs = []
        for src, extra in b.SOURCES.items():
            if src in over:
                o = os.path.join(odir, f"{name}_{src.replace('.cu', '.o')}")
                cmd = [b._nvcc()] + b.ARCH + [f for f in b.COMMON if f != "--use_fast_math=false"] + extra + over[src] + \
                      ["-c", os.path.join(b.CSRC, src), "-o", o]
                procs.append((name, subprocess.Popen(cmd)))
                objs.append(o)
            else:
                objs.append(os.path.join(b.OBJ_DIR, src.replace(".cu", ".o")))
        procs.append((name, ("link", [b._nvcc()] + b.ARCH + ["-shared", "-o", os.path.join(vdir, name + ".so")] + objs + ["-lcudart"])))
    for name, p in procs:
        if isinstance(p, tuple):
            continue
        if p.wait() != 0:
            raise SystemExit(f"nvcc failed for {name}")
    for name, p in procs:
        if isinstance(p, tuple):
            subprocess.check_call(p[1])
            print("built", name)


if __name__ == "__main__":
    main()
