"""ctypes binding of libdcnet_sm100.so.  Signatures are parsed from include/dcnet_b200.h, so the header is the single
source of truth for the C ABI.  There is NO fallback: if the library is missing or a call fails, an exception is raised."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "dcnet_b200.h")
LIB_PATH = os.path.join(HERE, "libdcnet_sm100.so")
ABI_VERSION = 4

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t, "unsigned int": ctypes.c_uint,
    "void": None,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [(argtype, argname), ...])} for every DCNET_API prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    protos = {}
    for m in re.finditer(r"DCNET_API\s+(.*?)\s*\b(dcnet_\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                alist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, alist)
    return protos


def _ctype(t):
    t = t.replace("const ", "").strip()
    if t.endswith("*"):
        return ctypes.c_char_p if t == "char*" else ctypes.c_void_p
    return _CTYPES[t]


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "dcnet_b200: %s is missing -- build it with `python -m dcnet_b200.build` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (ret, args) in _protos.items():
            fn = getattr(L, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = _ctype(ret)
            fn.argtypes = [_ctype(t) for t, _ in args]
        if L.dcnet_abi_version() != ABI_VERSION:
            raise RuntimeError("dcnet_b200: ABI version mismatch")
        _lib = L
    return _lib


def last_error():
    return lib().dcnet_last_error().decode()


def launch_count():
    return int(lib().dcnet_launch_count())


def call(name, *args):
    """Calls an int-returning entry point; raises RuntimeError(dcnet_last_error()) on failure."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError("dcnet_b200.%s failed (%d): %s" % (name, rc, last_error()))
