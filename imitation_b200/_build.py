"""Compile csrc/*.cu into imitation_b200/libimb.so for sm_100a (in-tree, no JIT cache)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
VARIANT = os.environ.get("IMB_VARIANT", "")  # e.g. "_timing" with IMB_EXTRA_NVCC_FLAGS=-DIMB_PPO_TIMING
LIB = os.path.join(HERE, "libimb" + VARIANT + ".so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--use_fast_math=false" if False else "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "imb.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "_obj" + VARIANT), exist_ok=True)
    procs = []
    hdr_t = max(os.path.getmtime(p) for p in glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "imb.h")])
    for src in sources():
        obj = os.path.join(HERE, "_obj" + VARIANT, os.path.basename(src) + ".o")
        objs.append(obj)
        if (not force and not verbose and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t)):
            continue  # object is up to date
        cmd = [nvcc, "-c", src, "-o", obj, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
               "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-DIMB_BUILDING"] + os.environ.get("IMB_EXTRA_NVCC_FLAGS", "").split()
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see stderr)")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
