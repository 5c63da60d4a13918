"""CPU suite: include/mst_b200.h is a plain-C header (compiles with gcc -std=c99 -pedantic) and every argument struct has the
same size and field offsets in the ctypes binding (mastermetastyletransfer_b200/_lib.py) as in C -- a mismatch would make the
kernels read the wrong pointers without any error."""
import ctypes
import os
import shutil
import subprocess

import pytest

from mastermetastyletransfer_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = ["MstGemm", "MstWgrad", "MstWindowAttnBwd", "MstWindowAttn", "MstAttnBlock", "MstMlp", "MstTensorTable", "MstLossTap", "MstLossTaps"]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_ctypes_structs_match_the_c_header(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mst_b200.h"', "int main(void) {"]
    for name in STRUCTS:
        st = getattr(_lib, name)
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out.splitlines()}
    for name in STRUCTS:
        st = getattr(_lib, name)
        assert got[(name, "size")] == ctypes.sizeof(st), name
        for field, _ in st._fields_:
            assert got[(name, field)] == getattr(st, field).offset, (name, field)


def test_header_declares_exactly_the_bound_structs():
    import re
    text = open(os.path.join(REPO, "include", "mst_b200.h")).read()
    declared = set(re.findall(r"typedef struct (\w+)", text))
    assert declared == set(STRUCTS), declared ^ set(STRUCTS)


def test_bound_prototypes_match_the_header():
    """Argument count and class (pointer / int / float / size_t) of every SYMBOLS entry against the header's prototypes."""
    import re
    text = re.sub(r"/\*.*?\*/", " ", open(os.path.join(REPO, "include", "mst_b200.h")).read(), flags=re.S)
    protos = re.findall(r"\b(int|size_t|const char\s*\*)\s+(mst_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert {p[1] for p in protos} == set(_lib.SYMBOLS)

    def klass_c(arg: str) -> str:
        arg = " ".join(arg.split())
        if "*" in arg:
            return "ptr"
        for t, k in (("size_t", "size"), ("float", "float"), ("double", "double"), ("int", "int")):
            if re.search(rf"\b{t}\b", arg):
                return k
        raise AssertionError(arg)

    def klass_py(t) -> str:
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) is not None and issubclass(t, ctypes._Pointer):
            return "ptr"
        return {ctypes.c_int: "int", ctypes.c_float: "float", ctypes.c_double: "double", ctypes.c_size_t: "size"}[t]

    for ret, name, args in protos:
        restype, argtypes = _lib.SYMBOLS[name]
        args = [a for a in args.split(",") if a.strip() and a.strip() != "void"]
        assert len(args) == len(argtypes), (name, args, argtypes)
        for a, t in zip(args, argtypes):
            assert klass_c(a) == klass_py(t), (name, a, t)
        want = {"int": ctypes.c_int, "size_t": ctypes.c_size_t}.get(ret, ctypes.c_char_p)
        assert restype is want, (name, ret, restype)
