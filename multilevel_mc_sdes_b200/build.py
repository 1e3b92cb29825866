"""Build libmlmcb.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmlmcb.so")
SOURCES = ["mlmcb_api.cu"]
HEADERS = ["mlmcb_kernels.cuh", os.path.join("..", "..", "include", "mlmcb.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libmlmcb.so")
    return p


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> csrc/libmlmcb.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libmlmcb.so")
    if verbose:
        sys.stderr.write(r.stderr)
    write_sass_mix()
    return LIB


def write_sass_mix():
    """Static instruction mix of the benchmarked kernel's inner loop -> csrc/sass_mix.json
    (bench.py reports it next to the measured cycles per step)."""
    import json
    tools = os.path.join(os.path.dirname(HERE), "tools")
    sys.path.insert(0, tools)
    try:
        import re
        import sass_mix
        hdr = open(os.path.join(CSRC, "mlmcb_kernels.cuh")).read()
        unroll = int(re.search(r"#define MLMCB_UNROLL (\d+)", hdr).group(1))
        out = {k: sass_mix.hot_loop_mix(LIB, pat, 4 * (unroll if k == "f32bm" else 1)) for k, pat in
               (("f32bm", "level_kernel<double, 0, 0, 4, 0, 0>"), ("f64bm", "level_kernel<double, 0, 0, 4, 1, 0>"))}
        json.dump(out, open(os.path.join(CSRC, "sass_mix.json"), "w"), indent=1)
    except Exception as e:  # cuobjdump missing: not fatal
        sys.stderr.write(f"sass_mix skipped: {e}\n")
    finally:
        sys.path.remove(tools)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
