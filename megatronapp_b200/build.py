"""In-tree build of the native extensions for sm_100a.

``build_all()`` compiles every ``csrc/*.cu`` with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (cross-compiles without a GPU),
the torch bindings with g++, links ``megatronapp_b200/_C.so`` and builds the pure-C++ dataset
helpers ``megatronapp_b200/data/_helpers.so``.  Objects are cached by content hash (not mtime)
so the snapshot that travels to a GPU box never rebuilds.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC",
              "-diag-suppress", "177", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build the sm_100a extension")
    return cand


def _hash_files(paths, extra="") -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(Path(p).name.encode())
        h.update(Path(p).read_bytes())
    return h.hexdigest()[:20]


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build command failed:\n" + " ".join(map(str, cmd)) + "\n" + r.stdout[-6000:])
    return r.stdout


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce

    inc = []
    for p in ce.include_paths():
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, libdir, abi


def build_all(verbose: bool = False, force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh"))
    cu_sources = sorted(CSRC.glob("*.cu")) + [p for p in sorted(CSRC.glob("*.cpp")) if p.name != "bindings.cpp"]
    stamp_path = BUILD / "stamps.json"
    stamps = json.loads(stamp_path.read_text()) if stamp_path.exists() else {}
    nvcc = _nvcc()
    cuda_home = str(Path(nvcc).resolve().parent.parent)
    inc, libdir, abi = _torch_flags()
    jobs = []
    objs = []
    for src in cu_sources:
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        key = _hash_files([src] + headers, " ".join(NVCC_FLAGS + ARCH_FLAGS))
        if not force and obj.exists() and stamps.get(obj.name) == key:
            continue
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        jobs.append((obj.name, key, cmd))
    bsrc = CSRC / "bindings.cpp"
    bobj = BUILD / "bindings.o"
    objs.append(bobj)
    cxx_flags = ["-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
                 "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
                 "-I", cuda_home + "/include", *inc, "-w"]
    bkey = _hash_files([bsrc] + headers, " ".join(cxx_flags))
    if force or not bobj.exists() or stamps.get(bobj.name) != bkey:
        jobs.append((bobj.name, bkey, ["g++", *cxx_flags, "-c", str(bsrc), "-o", str(bobj)]))

    def _do(job):
        name, key, cmd = job
        if verbose:
            print("[b200 build]", name, flush=True)
        _run(cmd)
        return name, key

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name, key in ex.map(_do, jobs):
                stamps[name] = key
    so = PKG / "_C.so"
    link_key = hashlib.sha256("".join(stamps.get(o.name, "") for o in objs).encode()).hexdigest()[:20]
    if force or jobs or not so.exists() or stamps.get("_C.so") != link_key:
        cmd = ["g++", "-shared", "-o", str(so), *map(str, objs), f"-L{libdir}", f"-L{cuda_home}/lib64",
               "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
               "-lcudart", "-lrt", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{cuda_home}/lib64"]
        _run(cmd)
        stamps["_C.so"] = link_key
    # dataset index helpers (pure C++ / pybind11)
    hsrc = PKG / "data" / "helpers.cpp"
    if hsrc.exists():
        hso = PKG / "data" / "_helpers.so"
        import pybind11

        hflags = ["-O3", "-std=c++17", "-fPIC", "-shared", "-I", pybind11.get_include(), "-I",
                  sysconfig.get_paths()["include"], "-w"]
        hkey = _hash_files([hsrc], " ".join(hflags))
        if force or not hso.exists() or stamps.get("_helpers.so") != hkey:
            _run(["g++", *hflags, str(hsrc), "-o", str(hso)])
            stamps["_helpers.so"] = hkey
    stamp_path.write_text(json.dumps(stamps, indent=1))
    return so


if __name__ == "__main__":
    path = build_all(verbose=True, force="--force" in sys.argv)
    print("built", path)
