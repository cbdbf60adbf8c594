"""In-tree build of the C-ABI library ``csrc/libms_b200.so`` (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU build container;
the resulting .so travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libms_b200.so")
SOURCES = ["ms_api.cu", "ms_stft_fft.cu", "ms_dft_i8.cu", "ms_dft_seg.cu", "ms_detect.cu", "ms_live.cu", "ms_pipeline.cu", "ms_io.cu", "ms_welch_qf.cu", "ms_welch_i8.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build csrc/libms_b200.so")


def dependencies():
    """Every file the library is built from: all sources and headers under csrc/ and include/ (and this recipe)."""
    import glob
    inc = os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include")
    deps = [p for pat in ("*.cu", "*.cuh", "*.h") for p in glob.glob(os.path.join(CSRC, pat))]
    deps += glob.glob(os.path.join(inc, "*.h"))
    deps.append(os.path.abspath(__file__))
    return sorted(deps)


K2_SOURCES = ["ms_dft_i8.cu", "ms_umma.cuh", "ms_async.cuh", "ms_common.cuh"]   # what dft_i8_kernel is compiled from


def source_hash(files=None) -> str:
    """sha256 over the sources the library is built from (stamps the SASS summary), or over ``files`` (names under
    csrc/): ``source_hash(K2_SOURCES)`` stamps profiles/traffic.json, the DRAM traffic of the dominant kernel."""
    import hashlib
    h = hashlib.sha256()
    for p in (dependencies() if files is None else [os.path.join(CSRC, f) for f in files] + [os.path.abspath(__file__)]):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in dependencies())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
