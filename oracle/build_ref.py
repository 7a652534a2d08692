"""ORACLE (test infrastructure): build the reference's OWN CUDA extension for sm_100a into
oracle/_ref/ so that GPU tests and tools can call the fork's real kernels on the B200 box.

    python oracle/build_ref.py            (needs /root/reference; ~10-15 min on 8 cores)

Sources are compiled where they lie (/root/reference/submodules/gsplat/gsplat/cuda/csrc/*.cu,
ext.cpp; vendored glm) with nvcc directly — the reference's own build system (setup.py / JIT
`torch.utils.cpp_extension.load`, gsplat/cuda/_backend.py:81-137) is NOT run; flags follow it
(`-O3 --use_fast_math`, _backend.py:93-100).  Nothing is copied into the repository: the only
outputs are object files and `gsplat_ref_csrc.so` under oracle/_ref/ (git-ignored; it travels
to the GPU box with the snapshot).  The module exposes the pybind11 entry points of
CS/ext.cpp:11-56 (`fully_fused_projection_fwd`, `isect_tiles`, `rasterize_to_pixels_fwd`, ...),
which tests/test_gpu_vs_reference.py and tests/reference_cuda_ab.py call with raw tensors —
no reference Python code is needed on the box.
"""
import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CS = Path("/root/reference/submodules/gsplat/gsplat/cuda/csrc")
OUT = ROOT / "oracle" / "_ref"
NAME = "gsplat_ref_csrc"


def main():
    if not CS.exists():
        print("reference sources not present; nothing to build")
        return 0
    import torch
    from torch.utils import cpp_extension as ce

    OUT.mkdir(parents=True, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{CS}",
                                                           f"-I{CS / 'third_party' / 'glm'}"]
    common = ["-O3", "-std=c++17", f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    nvcc_flags = ["--use_fast_math", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                  "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-D__CUDA_NO_HALF_OPERATORS__",
                  "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_BFLOAT16_CONVERSIONS__",
                  "-D__CUDA_NO_HALF2_OPERATORS__", "-w"]
    srcs = sorted(CS.glob("*.cu")) + sorted(CS.glob("*.cpp"))
    so = OUT / f"{NAME}.so"
    if so.exists() and so.stat().st_mtime >= max(p.stat().st_mtime for p in list(CS.glob("*")) if p.is_file()):
        print("up to date:", so)
        return 0

    def compile_one(src: Path):
        obj = OUT / (src.stem + ".o")
        if obj.exists() and obj.stat().st_mtime >= src.stat().st_mtime:
            return obj
        if src.suffix == ".cu":
            cmd = ["nvcc", *common, *nvcc_flags, *inc, "-c", str(src), "-o", str(obj)]
        else:
            cmd = ["g++", *common, "-fPIC", "-w", *inc, "-I/usr/local/cuda/include", "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{src.name}:\n{r.stderr[-4000:]}")
        print("compiled", src.name, flush=True)
        return obj

    with cf.ThreadPoolExecutor(max_workers=int(os.environ.get("MAX_JOBS", "8"))) as ex:
        objs = list(ex.map(compile_one, srcs))
    torch_lib = Path(torch.__file__).parent / "lib"
    cmd = ["g++", "-shared", "-o", str(so), *map(str, objs), f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu",
           "-ltorch_cuda", "-ltorch", "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart",
           f"-Wl,-rpath,{torch_lib}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    for o in objs:
        o.unlink()
    print("built", so, so.stat().st_size // (1 << 20), "MiB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
