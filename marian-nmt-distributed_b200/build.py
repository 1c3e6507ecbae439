"""Build recipe of the product library (nvcc, sm_100a) - used by __graft_entry__.build().

Everything is compiled IN-TREE into marian-nmt-distributed_b200/lib/ so the
shared objects travel to the GPU box with the repository snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_SOURCES = [
    "tensors/device_gpu.cu",
    "kernels/tensor_operators.cu",
    "kernels/gemm.cu",
    "kernels/attention.cu",
    "kernels/exchange.cu",
    "kernels/nth_element.cu",
]
# host graph code: instantiates the Element/Add kernel templates, hence nvcc -x cu
ENGINE_SOURCES = [
    "graph/node.cpp",
    "graph/expression_operators.cpp",
    "layers/generic.cpp",
    "rnn/cells.cpp",
    "models/model_factory.cpp",
    "capi/capi.cpp",
]
TEST_SOURCES = [os.path.join(ROOT, "tests", "cpp", "graph_golden.cpp")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fno-gnu-unique,-Wall,-Wno-unused-variable,-Wno-sign-compare,-Wno-unknown-pragmas,-Wno-unused-but-set-variable",
    "--expt-relaxed-constexpr",
    "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[-1])
    return r.stdout + r.stderr


def _deps_stamp():
    h = hashlib.sha1()
    for base, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            if f.endswith((".h", ".cu", ".cpp")):
                with open(os.path.join(base, f), "rb") as fh:
                    h.update(fh.read())
    for f in TEST_SOURCES + [os.path.join(ROOT, "include", "marian_b200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_product(verbose=False, extra_flags=()):
    """Compiles libmarian_b200.so (product) and libmarian_b200_tests.so (golden driver)."""
    os.makedirs(OBJDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "stamp")
    stamp = _deps_stamp() + " ".join(extra_flags)
    lib = os.path.join(LIBDIR, "libmarian_b200.so")
    tests_lib = os.path.join(LIBDIR, "libmarian_b200_tests.so")
    if os.path.exists(lib) and os.path.exists(tests_lib) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return lib

    def compile_one(src):
        path = src if os.path.isabs(src) else os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        cmd = [NVCC] + NVCC_FLAGS + list(extra_flags) + ["-x", "cu", "-c", path, "-o", obj]
        out = _run(cmd)
        if verbose and out.strip():
            print(out)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, CUDA_SOURCES + ENGINE_SOURCES))
        test_objs = list(ex.map(compile_one, TEST_SOURCES))

    # -Bsymbolic: the product and the test oracle define the same C++ symbols; each
    # library must bind to its OWN definitions when both are loaded in one process
    _run([NVCC, "-shared", "-o", lib] + objs + ["-lcudart", "-Xlinker", "-Bsymbolic"])
    _run([NVCC, "-shared", "-o", tests_lib] + test_objs + ["-L" + LIBDIR, "-lmarian_b200", "-Xlinker", "-rpath=$ORIGIN", "-Xlinker", "-Bsymbolic"])
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return lib


def build_oracle():
    """Builds the CPU test oracle (oracle/Makefile) and, where /root/reference exists, oracle/_ref."""
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8"])
    return os.path.join(ROOT, "oracle", "_build", "libmarian_oracle.so")


if __name__ == "__main__":
    print(build_product(verbose="-v" in sys.argv))
    print(build_oracle())
