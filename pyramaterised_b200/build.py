"""In-tree build of the CUDA engine: csrc/*.cu -> pyramaterised_b200/libpqc_b200.so.

sm_100a only (B200); nvcc cross-compiles without a GPU.  Run as
``python -m pyramaterised_b200.build`` or through ``__graft_entry__.build()``.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpqc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--use_fast_math=false"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")]
    procs = []
    for src in sources():
        obj = os.path.join(CSRC, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stdout.write(out.decode())
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


def build_variant(name, defines, verbose=False):
    """Developer A/B builds: pyramaterised_b200/variants/lib<name>.so with extra -D flags."""
    vdir = os.path.join(HERE, "variants")
    os.makedirs(os.path.join(vdir, name), exist_ok=True)
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")] + ["-D" + d for d in defines]
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(vdir, name, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stdout.write(out.decode())
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    lib = os.path.join(vdir, "lib%s.so" % name)
    subprocess.check_call([NVCC, "-shared", "-o", lib] + objs +
                          ["-gencode", "arch=compute_100a,code=sm_100a"])
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [d for d in sys.argv[i + 2:] if not d.startswith("-")],
                            verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
