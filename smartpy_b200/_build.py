"""In-tree build of libsmart_b200.so with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
_CSRC = os.path.join(_HERE, "csrc")
# (source, extra flags).  smart_kernels.cu holds the binary64 kernels and is compiled without
# implicit FMA contraction so that a member's bits do not depend on the kernel instantiation
# that ran it; smart_kernels_f32.cu re-includes it for the binary32-state kernels with the
# compiler's default contraction (see the note at the top of smart_kernels.cu).
UNITS = [
    ("smart_kernels.cu", ["-fmad=false"]),
    ("smart_kernels_f32.cu", []),
    ("smart_select.cu", []),
    ("smart_sample.cu", []),
    ("smart_io.cpp", []),          # host-side bulk text I/O of the sample database (include/smart_b200_io.h)
]
SOURCES = [os.path.join(_CSRC, name) for name, _ in UNITS]
HEADERS = [os.path.join(_CSRC, "smart_step.cuh"), os.path.join(ROOT, "include", "smart_b200.h"),
           os.path.join(ROOT, "include", "smart_b200_io.h")]
LIB_PATH = os.path.join(_HERE, "libsmart_b200.so")
OBJ_DIR = os.path.join(_CSRC, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def _run(cmd):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    return proc.stdout


def build(force=False, verbose=False):
    """Compile the CUDA kernels + C ABI.  Returns the path of the shared library."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    include = ["-I", os.path.join(ROOT, "include")]
    log, objects = [], []
    procs = []
    for name, extra in UNITS:
        obj = os.path.join(OBJ_DIR, os.path.splitext(name)[0] + ".o")
        objects.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + include + ["-c", os.path.join(_CSRC, name), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, proc in procs:                     # the units compile side by side
        out, _ = proc.communicate()
        log.append(out)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
    log.append(_run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objects + ["-lpthread"]))
    text = "".join(log)
    if verbose:
        print(text)
    with open(os.path.join(_CSRC, "ptxas_info.txt"), "w") as f:
        f.write(text)
    return LIB_PATH
