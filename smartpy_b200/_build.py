"""In-tree build of libsmart_b200.so with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SOURCES = [os.path.join(_HERE, "csrc", "smart_kernels.cu"), os.path.join(_HERE, "csrc", "smart_select.cu"),
           os.path.join(_HERE, "csrc", "smart_sample.cu")]
HEADERS = [os.path.join(_HERE, "csrc", "smart_step.cuh"), os.path.join(ROOT, "include", "smart_b200.h")]
LIB_PATH = os.path.join(_HERE, "libsmart_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA kernels + C ABI.  Returns the path of the shared library."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB_PATH] + SOURCES
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    with open(os.path.join(_HERE, "csrc", "ptxas_info.txt"), "w") as f:
        f.write(proc.stdout)
    return LIB_PATH
