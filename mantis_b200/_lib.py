"""ctypes loader for libmantis_b200.so (the C-ABI CUDA library).

There is NO CPU / PyTorch fallback: if the shared library is missing the import of any op fails loudly.
Prototypes are parsed from include/mantis_b200.h so the header is the single source of truth.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmantis_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "mantis_b200.h")
if not os.path.isfile(HEADER_PATH):                      # installed layout: header copied next to the package
    HEADER_PATH = os.path.join(_HERE, "include", "mantis_b200.h")

_CTYPES = {
    "int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "int64_t": ctypes.c_int64,
    "void": None, "const char*": ctypes.c_char_p,
}


def _ctype_of(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_char_p if decl.replace(" ", "") == "constchar*" else ctypes.c_void_p
    decl = decl.replace("const ", "").strip()
    return _CTYPES[decl]


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes], [argnames])} for every `mb200_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#") and "extern" not in l)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(mb200_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_ctype_of(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype_of(ret), argtypes, argnames)
    return protos


class MantisB200Error(RuntimeError):
    pass


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MantisB200Error(
                f"{LIB_PATH} not found. mantis_b200 has no CPU/PyTorch fallback: build the CUDA library first "
                f"(python -c 'import __graft_entry__ as g; g.build()' or make -C mantis_b200/csrc).")
        _lib = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (res, argtypes, _) in _protos.items():
            fn = getattr(_lib, name)          # AttributeError here == header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = argtypes
    return _lib


def protos():
    lib()
    return _protos


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().mb200_last_error()
        msg = msg.decode() if msg else ""
        if rc == -22:
            raise ValueError(f"{what}: invalid argument (-EINVAL) {msg}")
        raise MantisB200Error(f"{what} failed with {rc} {msg}")
