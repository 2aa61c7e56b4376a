"""Builds libnerfpp_b200.so (the C-ABI library of include/nerfpp_b200.h) in-tree with nvcc for sm_100a.

    python -m nerfpp_b200.build            # build if sources are newer than the library
    python -m nerfpp_b200.build --force

nvcc cross-compiles without a GPU.  The library links only the CUDA runtime: no torch, no Triton.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "lib"
LIB = LIB_DIR / "libnerfpp_b200.so"
INCLUDE = ROOT.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = sources() + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("NRF_NVCC_EXTRA", "").split(), "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src.name}\n{out}")
        failed |= p.returncode != 0
    (LIB_DIR / "build.log").write_text("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed, see nerfpp_b200/lib/build.log")
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
    subprocess.run(cmd, check=True)
    return LIB


HOST = ROOT / "host"
HOST_LIB = LIB_DIR / "nerfpp_b200_torch.so"


def _host_stale() -> bool:
    if not HOST_LIB.exists():
        return True
    t = HOST_LIB.stat().st_mtime
    deps = list(HOST.glob("*.cpp")) + list(HOST.glob("*.h")) + list(INCLUDE.glob("*.h"))
    return any(p.stat().st_mtime > t for p in deps)


def build_host(force: bool = False) -> Path:
    """g++ the C++ drop-in layer (nerfpp_b200/host) + its pybind11 front-end against the pip wheel's LibTorch and link it to
    libnerfpp_b200.so (rpath $ORIGIN).  No CUDA code here: every kernel sits behind the C ABI."""
    build()
    if not force and not _host_stale():
        return HOST_LIB
    import sysconfig
    import pybind11
    import torch
    tdir = Path(torch.__file__).resolve().parent
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = ["-I", str(INCLUDE), "-I", str(HOST), "-I", str(tdir / "include"), "-I", str(tdir / "include/torch/csrc/api/include"),
           "-I", f"{cuda}/include", "-I", sysconfig.get_paths()["include"], "-I", pybind11.get_include()]
    flags = ["-std=c++17", "-O2", "-fPIC", "-w", "-D_GLIBCXX_USE_CXX11_ABI=1", "-DTORCH_EXTENSION_NAME=nerfpp_b200_torch"]
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(parents=True, exist_ok=True)
    procs, objs = [], []
    for src in sorted(HOST.glob("*.cpp")):
        obj = obj_dir / ("host_" + src.stem + ".o")
        objs.append(obj)
        if not force and obj.exists() and obj.stat().st_mtime > max(p.stat().st_mtime for p in [src, *HOST.glob("*.h"), *INCLUDE.glob("*.h")]):
            continue
        procs.append((src, subprocess.Popen(["g++", *flags, *inc, "-c", str(src), "-o", str(obj)], stdout=subprocess.PIPE,
                                             stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"== {src.name}\n{out}\n")
    if failed:
        raise RuntimeError("g++ failed on the host layer")
    subprocess.run(["g++", "-shared", "-o", str(HOST_LIB), *map(str, objs), f"-L{LIB_DIR}", "-lnerfpp_b200", f"-L{tdir / 'lib'}",
                    "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", "-ltorch_cuda", "-lc10_cuda", f"-L{cuda}/lib64", "-lcudart",
                    "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tdir / 'lib'}", f"-Wl,-rpath,{cuda}/lib64"], check=True)
    return HOST_LIB


if __name__ == "__main__":
    if "--host" in sys.argv:
        print(build_host(force="--force" in sys.argv))
        sys.exit(0)
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
