"""ctypes binding of librefil_b200.so.

The signatures are parsed from include/refil_b200.h (the C ABI is the single source of truth).  There is NO
fallback: if the library is missing or a symbol fails to resolve, importing / calling raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "librefil_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "refil_b200.h")

_SCALARS = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong, "cudaStream_t": ctypes.c_void_p, "size_t": ctypes.c_size_t,
}


class RefilError(RuntimeError):
    pass


def parse_header(path=HEADER):
    """-> {name: (restype_str, [(ctype, c_type_str, arg_name), ...])} for every `refil_*` prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(refil_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    parsed.append((ctypes.c_void_p, a[:a.rindex("*") + 1].strip(), a[a.rindex("*") + 1:].strip()))
                else:
                    ty, nm = a.rsplit(" ", 1)
                    ty = ty.replace("const ", "").strip()
                    parsed.append((_SCALARS[ty], ty, nm))
        protos[name] = (ret, parsed)
    return protos


PROTOS = parse_header()
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RefilError("librefil_b200.so is not built (%s missing): run `python -m refil_b200.build` "
                         "or __graft_entry__.build(); there is no CPU fallback" % SO_PATH)
    import torch  # noqa: F401  (loads libcudart.so.12 first so the runtime instance is shared with PyTorch)
    lib = ctypes.CDLL(SO_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (ret, args) in PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.argtypes = [a[0] for a in args]
        fn.restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
    _lib = lib
    return lib


def last_error():
    return load().refil_last_error().decode(errors="replace")


def call(name, *args):
    """Call `refil_<name>` and raise RefilError on a non-zero return code."""
    fn = getattr(load(), "refil_" + name)
    rc = fn(*args)
    if rc != 0:
        raise RefilError("refil_%s failed (%d): %s" % (name, rc, last_error()))
