"""Build libmlmcb.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmlmcb.so")
# one translation unit per (kernel family, sampler), compiled side by side: (source, object name, extra -D flags)
UNITS = [("mlmcb_api.cu", "api.o", [])] + [
    ("mlmcb_family.cu", f"family{fam}_s{samp}.o", [f"-DMLMCB_TU_FAMILY={fam}", f"-DMLMCB_TU_SAMP={samp}"])
    for fam in (0, 1, 2, 3, 4, 5) for samp in (0, 1)]
SOURCES = ["mlmcb_api.cu", "mlmcb_family.cu"]
HEADERS = ["mlmcb_kernels.cuh", "mlmcb_pick.cuh", "mlmcb_f64tab.inc", os.path.join("..", "..", "include", "mlmcb.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
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


def compile_lib(out, extra=(), verbose=False, only=None):
    """nvcc -c every unit (in parallel), then link them into the shared library `out`.
    `only`: object names (e.g. ["family0_s1.o"]) to compile with `extra`; the other objects are taken from the
    main build's object directory (csrc/libmlmcb.so.obj) - how a kernel experiment rebuilds one family in seconds."""
    objdir = out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    main_objdir = LIB + ".obj"
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    procs, objs = [], []
    for src, objname, defs in UNITS:
        obj = os.path.join(objdir, objname)
        if only is not None and objname not in only:
            objs.append(os.path.join(main_objdir, objname))
            continue
        objs.append(obj)
        procs.append((obj, subprocess.Popen([nvcc_path()] + flags + list(extra) + defs + ["-c", "-o", obj, src], cwd=CSRC,
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, ok = "", True
    for obj, p in procs:
        log += p.communicate()[0]
        ok = ok and p.returncode == 0
    if ok:
        r = subprocess.run([nvcc_path(), "-shared", "-o", out] + objs, cwd=CSRC, capture_output=True, text=True)
        log += r.stdout + r.stderr
        ok = r.returncode == 0
    if not ok:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building " + os.path.basename(out))
    if verbose:
        sys.stderr.write(log)
    return out


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> csrc/libmlmcb.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    compile_lib(LIB, verbose=verbose)
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
        unroll64 = int(re.search(r"#define MLMCB_UNROLL_F64 (\d+)", hdr).group(1))
        # fp32 sampler: four steps per draw block; fp64 sampler: eight (an octet of three Philox blocks)
        out = {k: sass_mix.hot_loop_mix(LIB, pat, 4 * unroll if k == "f32bm" else 8 * unroll64) for k, pat in
               (("f32bm", "level_kernel<double, 0, 0, 4, 0, 0, 0>"), ("f64bm", "level_kernel<double, 0, 0, 4, 1, 0, 2>"))}
        json.dump(out, open(os.path.join(CSRC, "sass_mix.json"), "w"), indent=1)
    except Exception as e:  # cuobjdump missing: not fatal
        sys.stderr.write(f"sass_mix skipped: {e}\n")
    finally:
        sys.path.remove(tools)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
