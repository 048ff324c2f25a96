"""In-tree build of the native core (C++17 host code + sm_100a CUDA).

`python -m accl_b200.utils.build` compiles every source under csrc/ and links
`accl_b200/_C*.so` (pybind11 module, static cudart, no libcuda link: the driver
API is resolved lazily so the module imports on GPU-less machines) plus the
standalone tools under build/bin.  nvcc cross-compiles for sm_100a without a
GPU; the resulting .so travels to the GPU box as-is.

Replaces the reference's CMake/Make/Vitis flow (driver/xrt/CMakeLists.txt,
kernels/cclo/Makefile, test/refdesigns/Makefile).
"""
import concurrent.futures as cf
import os
import shlex
import subprocess
import sys
import sysconfig
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
CSRC = ROOT / "csrc"
OBJ = ROOT / "build" / "obj"
BIN = ROOT / "build" / "bin"

NVCC = os.environ.get("ACCL_NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("ACCL_CXX", "/usr/bin/g++")  # NOT $CXX: the image exports a toolchain whose libstdc++ is static
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

HOST_SOURCES = [
    "src/host/accl.cpp", "src/host/arithconfig.cpp", "src/host/bootstrap.cpp", "src/host/common.cpp",
    "src/host/communicator.cpp", "src/host/constants.cpp",
    "src/emu/engine.cpp", "src/emu/firmware.cpp", "src/emu/fabric.cpp", "src/emu/emudevice.cpp", "src/emu/remote.cpp",
]
CUDA_HOST_SOURCES = ["src/cuda/driver_api.cpp", "src/cuda/symheap.cpp"]
CUDA_SOURCES = []  # filled below from csrc/src/cuda/*.cu
BINDING = "bindings/pyaccl.cpp"


def _includes():
    import pybind11
    return ["-I" + str(CSRC / "include"), "-I" + pybind11.get_include(),
            "-I" + sysconfig.get_paths()["include"], "-I/usr/local/cuda/include"]


def _newest_header_mtime():
    m = 0.0
    for p in (CSRC / "include").rglob("*"):
        if p.is_file():
            m = max(m, p.stat().st_mtime)
    for p in (CSRC / "src").rglob("*.hpp"):
        m = max(m, p.stat().st_mtime)
    for p in (CSRC / "src").rglob("*.cuh"):
        m = max(m, p.stat().st_mtime)
    return m


def _run(cmd, verbose):
    if verbose:
        print(" ".join(shlex.quote(c) for c in cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd[:6]) + " ...")
    if verbose and (r.stdout or r.stderr):
        sys.stderr.write(r.stdout + r.stderr)


def _compile(src, with_cuda, hdr_mtime, verbose, force):
    srcp = CSRC / src
    obj = OBJ / (src.replace("/", "_") + ".o")
    if not force and obj.exists() and obj.stat().st_mtime > max(srcp.stat().st_mtime, hdr_mtime):
        return obj
    defs = ["-DACCL_WITH_CUDA"] if with_cuda else []
    # extra -D switches for experiments, e.g.
    # ACCL_EXTRA_DEFINES=MY_SWITCH python -m accl_b200.utils.build -f
    defs += ["-D" + d for d in os.environ.get("ACCL_EXTRA_DEFINES", "").split(",") if d]
    if src.endswith(".cu"):
        cmd = [NVCC, "-ccbin", CXX, "-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-fvisibility=hidden",
               "--expt-relaxed-constexpr", "-Xptxas", "-v" if verbose else "-O3",
               *defs, *_includes(), "-c", str(srcp), "-o", str(obj)]
    else:
        cmd = [CXX, "-std=c++17", "-O2", "-g1", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unused-function",
               *defs, *_includes(), "-c", str(srcp), "-o", str(obj)]
    _run(cmd, verbose)
    return obj


def ext_path():
    return ROOT / "accl_b200" / ("_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(with_cuda=True, verbose=False, force=False, tools=True, out_path=None):
    OBJ.mkdir(parents=True, exist_ok=True)
    BIN.mkdir(parents=True, exist_ok=True)
    cu = sorted(str(p.relative_to(CSRC)) for p in (CSRC / "src" / "cuda").glob("*.cu"))
    cuda_host = sorted(str(p.relative_to(CSRC)) for p in (CSRC / "src" / "cuda").glob("*.cpp"))
    srcs = list(HOST_SOURCES) + [BINDING]
    if with_cuda:
        srcs += cuda_host + cu
    hdr = _newest_header_mtime()
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, with_cuda, hdr, verbose, force), srcs))
    out = Path(out_path) if out_path else ext_path()
    if force or not out.exists() or any(o.stat().st_mtime > out.stat().st_mtime for o in objs):
        if with_cuda:
            link = [NVCC, "-ccbin", CXX, "-shared", *ARCH, "-cudart", "static", "-Xcompiler", "-fPIC",
                    *[str(o) for o in objs], "-o", str(out), "-lpthread", "-ldl", "-lrt"]
        else:
            link = [CXX, "-shared", "-fPIC", *[str(o) for o in objs], "-o", str(out), "-lpthread"]
        _run(link, verbose)
    if tools:
        # the C++ library on its own (no Python): build/lib/libaccl.a + headers under csrc/include.  Static,
        # because the objects are compiled with hidden visibility for the Python extension; link with
        #   g++ app.cpp -Icsrc/include build/lib/libaccl.a [libcudart_static.a] -lpthread -ldl -lrt
        lib = ROOT / "build" / "lib" / "libaccl.a"
        lib.parent.mkdir(parents=True, exist_ok=True)
        core_objs = [o for o, s_ in zip(objs, srcs) if s_ != BINDING and "bind_" not in s_]
        if force or not lib.exists() or any(o.stat().st_mtime > lib.stat().st_mtime for o in core_objs):
            if lib.exists():
                lib.unlink()
            _run(["ar", "rcs", str(lib), *[str(o) for o in core_objs]], verbose)
    if tools and not with_cuda:
        for name in ("cclo_emu", "emu_selftest", "emu_suite", "emu_bench", "emu_fuzz"):
            build_tool(name, verbose)
    if tools and with_cuda:
        lib_objs = [o for o, s in zip(objs, srcs) if s != BINDING and "bind_" not in s]
        for tool in sorted((CSRC / "tools").glob("*.cu")) + sorted((CSRC / "tools").glob("*.cpp")):
            exe = BIN / tool.stem
            if not force and exe.exists() and exe.stat().st_mtime > max(tool.stat().st_mtime, hdr,
                                                                        max(o.stat().st_mtime for o in lib_objs)):
                continue
            _run([NVCC, "-ccbin", CXX, "-std=c++17", "-O3", "-lineinfo", *ARCH, "--expt-relaxed-constexpr", "-DACCL_WITH_CUDA",
                  "-I" + str(CSRC / "include"), str(tool), *[str(o) for o in lib_objs], "-o", str(exe),
                  "-lpthread", "-ldl", "-lrt"], verbose)
    return out


def check_experimental(defines, verbose=True):
    """Compile (do not link) every CUDA-backend source with extra -D switches into build/obj_exp: the way to
    syntax- and resource-check code that is kept out of the default build (any -D switch)."""
    out = ROOT / "build" / "obj_exp"
    out.mkdir(parents=True, exist_ok=True)
    srcs = sorted(str(p.relative_to(CSRC)) for p in (CSRC / "src" / "cuda").glob("*.cu")) + \
        sorted(str(p.relative_to(CSRC)) for p in (CSRC / "src" / "cuda").glob("*.cpp"))
    dflags = ["-D" + d for d in defines] + ["-DACCL_WITH_CUDA"]
    for src in srcs:
        obj = out / (src.replace("/", "_") + ".o")
        if src.endswith(".cu"):
            cmd = [NVCC, "-ccbin", CXX, "-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                   "-Xptxas", "-v", *dflags, *_includes(), "-c", str(CSRC / src), "-o", str(obj)]
        else:
            cmd = [CXX, "-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-unused-function", *dflags, *_includes(), "-c",
                   str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("experimental compile failed: " + src)
        if verbose:
            for line in (r.stdout + r.stderr).splitlines():
                if "Used" in line or "spill" in line or "Compiling entry" in line or "warning" in line:
                    print(line)
    return out


def build_variant(name, defines, verbose=False):
    """A second, self-contained copy of the package built with extra -D switches:
    build/variants/<name>/accl_b200 (python sources copied, own _C extension, own object directory).
    `PYTHONPATH=build/variants/<name> python ...` then runs that build next to the default one — one GPU
    session can compare both (any set of -D switches)."""
    import shutil
    global OBJ
    pkg = ROOT / "build" / "variants" / name / "accl_b200"
    if pkg.exists():
        shutil.rmtree(pkg)
    shutil.copytree(ROOT / "accl_b200", pkg, ignore=shutil.ignore_patterns("__pycache__", "*.so"))
    saved_obj, saved_env = OBJ, os.environ.get("ACCL_EXTRA_DEFINES")
    OBJ = ROOT / "build" / ("obj_" + name)
    os.environ["ACCL_EXTRA_DEFINES"] = ",".join(defines)
    try:
        out = build(with_cuda=True, verbose=verbose, force=False, tools=False, out_path=pkg / ext_path().name)
    finally:
        OBJ = saved_obj
        if saved_env is None:
            os.environ.pop("ACCL_EXTRA_DEFINES", None)
        else:
            os.environ["ACCL_EXTRA_DEFINES"] = saved_env
    return out


def build_tool(name, verbose=False):
    """Host-only tools (cclo_emu, emu_selftest) need no CUDA toolchain: g++ build into build/bin."""
    BIN.mkdir(parents=True, exist_ok=True)
    exe = BIN / name
    srcs = [str(CSRC / s) for s in HOST_SOURCES] + [str(CSRC / "tools" / (name + ".cpp"))]
    _run([CXX, "-std=c++17", "-O2", "-g1", "-I" + str(CSRC / "include"), *srcs, "-o", str(exe), "-lpthread"], verbose)
    return exe


def build_sanitized(kind="address", verbose=False, tool="emu_selftest"):
    """CPU-only build of the emulator self-test under a sanitizer (`address` = ASan+UBSan,
    `thread` = TSan): build/bin/emu_selftest_<kind>.  The reference configures none
    (SURVEY 5.2); the emulator is the concurrent part of the host code (control thread, data
    mover, fabric threads per rank), so this is where races and lifetime bugs would live."""
    BIN.mkdir(parents=True, exist_ok=True)
    flags = {"address": ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"],
             "thread": ["-fsanitize=thread"]}[kind]
    exe = BIN / f"{tool}_{kind}"
    srcs = [str(CSRC / s) for s in HOST_SOURCES] + [str(CSRC / "tools" / (tool + ".cpp"))]
    _run([CXX, "-std=c++17", "-O1", "-g", *flags, "-I" + str(CSRC / "include"), *srcs, "-o", str(exe), "-lpthread"],
         verbose)
    return exe


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("-f", "--force", action="store_true")
    ap.add_argument("--no-tools", action="store_true")
    ap.add_argument("--sanitize", choices=["address", "thread"], help="build build/bin/<tool>_<kind> only")
    ap.add_argument("--suite", action="store_true", help="with --sanitize: build the full emu_suite instead of emu_selftest")
    ap.add_argument("--tool", default=None, help="with --sanitize: which csrc/tools/<tool>.cpp to build (emu_selftest, emu_suite, emu_fuzz)")
    ap.add_argument("--variant", metavar="NAME", help="with --define: build build/variants/NAME/accl_b200 with the extra macros")
    ap.add_argument("--define", action="append", default=[], metavar="MACRO")
    ap.add_argument("--check-define", action="append", default=[], metavar="MACRO",
                    help="compile the CUDA backend with -DMACRO into build/obj_exp (no link): syntax / resource check")
    a = ap.parse_args()
    if a.variant:
        print(build_variant(a.variant, a.define, a.verbose))
        sys.exit(0)
    if a.check_define:
        print(check_experimental(a.check_define))
        sys.exit(0)
    if a.sanitize:
        print(build_sanitized(a.sanitize, a.verbose, a.tool or ("emu_suite" if a.suite else "emu_selftest")))
        sys.exit(0)
    print(build(with_cuda=not a.cpu_only, verbose=a.verbose, force=a.force, tools=not a.no_tools))
