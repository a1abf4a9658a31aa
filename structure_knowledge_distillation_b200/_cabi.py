"""ctypes binding of libskd_b200.so (the C ABI declared in include/skd.h).

The signatures are parsed from the header itself, so the Python side can never drift from the declared ABI.
Error convention follows the reference (libs/functions.py:13-16): a 0 return raises
RuntimeError("CUDA Error encountered in <fn>").  There is NO fallback: if the library is missing the import of any
op raises -- the product path never silently routes through torch/CPU code.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libskd_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "skd.h")

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong, "void": None,
    "cudaStream_t": ctypes.c_void_p,
}

_DECL = re.compile(r"^\s*(const\s+char\s*\*|int|long long|void)\s+(skd_\w+)\s*\(([^;{]*)\)\s*;", re.M | re.S)


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every function declared in include/skd.h."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for ret, name, args in _DECL.findall(text):
        ret = ret.strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret]
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                elif a.startswith("cudaStream_t"):
                    argtypes.append(ctypes.c_void_p)
                elif a.startswith("long long"):
                    argtypes.append(ctypes.c_longlong)
                elif a.startswith("float"):
                    argtypes.append(ctypes.c_float)
                elif a.startswith("int"):
                    argtypes.append(ctypes.c_int)
                else:
                    raise ValueError("unparsed argument %r in %s" % (a, name))
        out[name] = (restype, argtypes)
    return out


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libskd_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C structure_knowledge_distillation_b200/csrc`. There is no CPU/torch fallback." % LIB_PATH)
        self._dll = ctypes.CDLL(LIB_PATH)
        self.signatures = parse_header()
        self._checked = set()
        for name, (restype, argtypes) in self.signatures.items():
            fn = getattr(self._dll, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
            if restype is ctypes.c_int and argtypes and argtypes[-1] is ctypes.c_void_p and name != "skd_pool_out_size_ceil":
                self._checked.add(name)

    def last_error(self):
        return self._dll.skd_last_error().decode()

    def __getattr__(self, name):
        fn = getattr(self._dll, name)
        if name not in self._checked:
            return fn

        def checked(*args):
            if not fn(*args):
                raise RuntimeError("CUDA Error encountered in {} ({})".format(name, self.last_error()))
            return 1
        checked.__name__ = name
        setattr(self, name, checked)
        return checked


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
