"""Builds libkpms_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

kalman.cu and hmm.cu carry the kernels that are unrolled per (latent_dim, nlags) pair; they are compiled once per
group of pairs (csrc/common.cuh: KPMS_DL_GROUP_g) with -DKPMS_DL_GROUP=g so that the groups build in parallel."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkpms_b200.so")
SOURCES = ["capi.cu", "hmm.cu", "kalman.cu", "elementwise.cu", "stats.cu", "params.cu"]
GROUPED = ("kalman.cu", "hmm.cu")            # compiled once per (latent_dim, nlags) group, longest first
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def dl_groups():
    """Number of (latent_dim, nlags) groups declared in common.cuh."""
    text = open(os.path.join(CSRC, "common.cuh")).read()
    return int(re.search(r"#define\s+KPMS_DL_GROUPS\s+(\d+)", text).group(1))


def build(verbose=False, force=False, jobs=None):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("KPMS_NVCC_EXTRA", "").split()
    jobs = jobs or int(os.environ.get("KPMS_BUILD_JOBS", os.cpu_count() or 4))
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "kpms_b200.h"))
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    units = []                                # (label, source path, object path, extra defines)
    for src in GROUPED:
        for g in range(dl_groups()):
            units.append((f"{src}[{g}]", src, src.replace(".cu", f"_g{g}.o"), [f"-DKPMS_DL_GROUP={g}"]))
    for src in SOURCES:
        if src not in GROUPED:
            units.append((src, src, src.replace(".cu", ".o"), []))
    objs, todo = [], []
    for label, src, obj, defs in units:
        path, obj = os.path.join(CSRC, src), os.path.join(HERE, "build", obj)
        objs.append(obj)
        if force or _stale(obj, [path] + hdrs):
            todo.append((label, [nvcc] + FLAGS + extra + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]))
    running, failed = [], False
    while todo or running:
        while todo and len(running) < jobs:
            label, cmd = todo.pop(0)
            running.append((label, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        label, p = running.pop(0)
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(f"--- {label}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    stale_objs = [o for o in os.listdir(os.path.join(HERE, "build")) if os.path.join(HERE, "build", o) not in objs]
    for o in stale_objs:                      # objects of an older source layout must not be linked
        os.remove(os.path.join(HERE, "build", o))
    if _stale(LIB, objs):
        subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
