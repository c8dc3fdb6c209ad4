"""ctypes binding of csrc/libcmx_b200.so.  The prototypes are parsed from include/cmx_b200.h so the
header stays the single source of truth for the C ABI.  The library is loaded lazily, per process
(the reference's evaluator pickles the model into spawned children: engine/evaluator.py:131-137)."""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "cmx_b200.h")
LIB_PATH = os.path.join(_HERE, "csrc", "libcmx_b200.so")

_lock = threading.Lock()
_lib = None


class CmxGemm(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("C", ctypes.c_void_p),
        ("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("row_scale", ctypes.c_void_p),
        ("M", ctypes.c_int64), ("N", ctypes.c_int64), ("K", ctypes.c_int64),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldc", ctypes.c_int64), ("ldr", ctypes.c_int64),
        ("trans_a", ctypes.c_int32), ("trans_b", ctypes.c_int32),
        ("batch1", ctypes.c_int32), ("batch2", ctypes.c_int32),
        ("sA1", ctypes.c_int64), ("sA2", ctypes.c_int64), ("sB1", ctypes.c_int64), ("sB2", ctypes.c_int64),
        ("sC1", ctypes.c_int64), ("sC2", ctypes.c_int64),
        ("c_dtype", ctypes.c_int32), ("r_dtype", ctypes.c_int32),
        ("act", ctypes.c_int32), ("alpha", ctypes.c_float),
        ("accumulate", ctypes.c_int32), ("split_k", ctypes.c_int32),
        ("rows_per_sample", ctypes.c_int32), ("impl", ctypes.c_int32),
        ("sBias1", ctypes.c_int64), ("sR1", ctypes.c_int64), ("sS1", ctypes.c_int64),
    ]


_CTYPE = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
    "double": ctypes.c_double, "long long": ctypes.c_longlong,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every `int|long long|const char* cmx_*(...)` prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int64_t|int)\s+(cmx_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        args = " ".join(args.split())
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const ", "").rsplit(" ", 1)[0].strip()
                    argtypes.append(_CTYPE[ty])
        restype = {"const char*": ctypes.c_char_p, "long long": ctypes.c_longlong, "int64_t": ctypes.c_int64, "int": ctypes.c_int}[ret]
        protos[name] = (restype, argtypes)
    return protos


def lib_available():
    return os.path.exists(LIB_PATH)


def load():
    """Load (once per process) and return the ctypes library; raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "cmx_b200: CUDA library %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error():
    return load().cmx_last_error().decode(errors="replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("cmx_b200.%s failed (rc=%d): %s" % (what, rc, last_error()))
