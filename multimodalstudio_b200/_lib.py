"""ctypes binding of libmms_b200.so (the C ABI declared in include/mms_b200.h).

The prototypes are parsed from the header itself so the Python side cannot drift from the ABI.
There is no fallback: if the shared library is missing (or a CUDA tensor is not given) the ops
raise — nothing in the product path computes on the CPU.
"""
import ctypes
import os
import re
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_ROOT = os.path.dirname(_PKG_DIR)
HEADER_PATH = os.path.join(_REPO_ROOT, "include", "mms_b200.h")
LIB_PATH = os.environ.get("MMSB_LIB") or os.path.join(_PKG_DIR, "libmms_b200.so")   # MMSB_LIB: dev builds only
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

MMSB_MAX_LEVELS = 32


class MmsbHashGridDesc(ctypes.Structure):
    _fields_ = [
        ("num_levels", ctypes.c_int32),
        ("features_per_level", ctypes.c_int32),
        ("log2_hashmap_size", ctypes.c_int32),
        ("interpolation", ctypes.c_int32),
        ("radius", ctypes.c_float),
        ("resolution", ctypes.c_float * MMSB_MAX_LEVELS),
    ]


_SCALARS = {
    "int": ctypes.c_int,
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
    "mmsb_stream_t": ctypes.c_void_p,
}


def _ctype_of(decl: str):
    decl = decl.strip()
    if "*" in decl:
        if "char" in decl:
            return ctypes.c_char_p
        if "MmsbHashGridDesc" in decl:
            return ctypes.POINTER(MmsbHashGridDesc)
        return ctypes.c_void_p
    typ = decl.replace("const", "").split()[0]
    return _SCALARS[typ]


def parse_header(path: str = HEADER_PATH):
    """Returns {name: (restype, [argtypes], [argnames])} for every `mmsb_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int64_t|int)\s+(mmsb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = " ".join(args.split())
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                nm = re.search(r"(\w+)$", a).group(1)
                argtypes.append(_ctype_of(a[: -len(nm)]))
                argnames.append(nm)
        restype = {"const char*": ctypes.c_char_p, "int64_t": ctypes.c_int64, "int": ctypes.c_int}[ret]
        protos[name] = (restype, argtypes, argnames)
    return protos


def build_library(verbose: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a into libmms_b200.so (in-tree). nvcc cross-compiles without a GPU."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libmms_b200.so failed")
    return LIB_PATH


class MmsbError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load_library():
    """Loads (once) the in-tree shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MmsbError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes, _) in parse_header().items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    return load_library().mmsb_last_error().decode()


def check(code: int, what: str = ""):
    """Maps the C error convention onto the reference's Python one."""
    if code == 0:
        return
    msg = last_error()
    if code in (-1, -2):
        raise ValueError(f"{what}: {msg}")
    raise MmsbError(f"{what}: {msg} (code {code})")


def launch_count() -> int:
    return int(load_library().mmsb_launch_count())


def ptr(t):
    """Device pointer of a CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("mms_b200 ops need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_timed = None      # set of entry points to bracket with CUDA events (bench.py roofline pass), else None
_records = []


class _Everything:
    def __contains__(self, name):
        return True


def start_kernel_timing(names=None):
    """names: entry points to bracket with CUDA events; None = every entry point."""
    global _timed, _records
    _timed, _records = (set(names) if names is not None else _Everything()), []


def stop_kernel_timing():
    """-> [(name, milliseconds, args)] for every instrumented call since start_kernel_timing (caller synchronised)."""
    global _timed
    _timed = None
    return [(n, s.elapsed_time(e), a) for n, s, e, a in _records]


def call(name: str, *args):
    lib = load_library()
    if _timed is not None and name in _timed:
        import torch

        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = getattr(lib, name)(*args)
        e.record()
        _records.append((name, s, e, args))
        check(rc, name)
        return
    check(getattr(lib, name)(*args), name)
