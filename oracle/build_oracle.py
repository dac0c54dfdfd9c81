"""Builds the checker artefacts (test infrastructure, never used by the product):

  oracle/_build/libiou3d_oracle.so   gcc, from oracle/iou3d_oracle.c (the plain-C restatement)
  oracle/_ref/iou3d_cpu_ref.so       g++, from the REFERENCE's own pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp where it lies under
                                     /root/reference + oracle/ref_iou3d_binding.cpp - only when /root/reference exists
                                     (this container); the GPU box uses the prebuilt file, which travels with gpurun.

No reference source is copied into the repo; both output directories are git-ignored."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp"
ORACLE_SO = os.path.join(HERE, "_build", "libiou3d_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "iou3d_cpu_ref.so")


def _stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def build_oracle(verbose=False):
    src = os.path.join(HERE, "iou3d_oracle.c")
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    if _stale(ORACLE_SO, [src]):
        # -ffp-contract=off: no fused multiply-adds, the arithmetic is the reference's (x86-64 gcc default for its build)
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", ORACLE_SO, src, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return ORACLE_SO


def build_ref(verbose=False):
    """-> path of the compiled reference CPU IoU, or None when /root/reference is absent and nothing was prebuilt"""
    if not os.path.exists(REF_SRC):
        return REF_SO if os.path.exists(REF_SO) else None
    shim = os.path.join(HERE, "ref_iou3d_binding.cpp")
    os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
    if _stale(REF_SO, [REF_SRC, shim]):
        import torch
        from torch.utils import cpp_extension as ce
        inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", "-I/usr/local/cuda/include",
                                                        f"-I{os.path.dirname(REF_SRC)}"]
        libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
        cmd = (["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-DTORCH_EXTENSION_NAME=iou3d_cpu_ref",
                "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + inc +
               [REF_SRC, shim, "-o", REF_SO, f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", f"-Wl,-rpath,{libdir}"])
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return REF_SO


def load_ref():
    """import the compiled reference module (torch must be imported first)"""
    import importlib.util
    import torch  # noqa: F401
    path = build_ref()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("iou3d_cpu_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_oracle(verbose=True))
    print(build_ref(verbose=True))
