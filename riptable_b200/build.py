"""Builds libriptable_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the repo."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libriptable_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-DRTB_BUILD",
] + os.environ.get("RTB_NVCC_FLAGS", "").split()  # measurement builds only (e.g. -DRTB_FUSE_MINB=4)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "include", "riptable_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("riptable_b200: CUDA build failed")
    subprocess.check_call([NVCC, "-Wno-deprecated-gpu-targets", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


NCCL_SRC = os.path.join(HERE, "csrc_nccl", "sharded_nccl.cu")
NCCL_LIB = os.path.join(HERE, "libriptable_b200_nccl.so")


def build_nccl(force=False):
    """libriptable_b200_nccl.so: the multi-GPU C ABI (include/riptable_b200_nccl.h).  A library of its own, linked against
    libriptable_b200.so and libnccl.so.2, so that the single-GPU library carries no NCCL dependency."""
    build(force=force)
    hdr = os.path.join(HERE, "..", "include", "riptable_b200_nccl.h")
    if not force and os.path.exists(NCCL_LIB) and all(os.path.getmtime(NCCL_LIB) >= os.path.getmtime(d) for d in (NCCL_SRC, hdr, LIB)):
        return NCCL_LIB
    cmd = [NVCC, "-Wno-deprecated-gpu-targets", *FLAGS, "-shared", NCCL_SRC, "-o", NCCL_LIB, "-L" + HERE, "-lriptable_b200", "-lnccl", "-lcudart",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError("riptable_b200: building the NCCL library failed:\n" + out.stdout)
    return NCCL_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_nccl(force="--force" in sys.argv))
