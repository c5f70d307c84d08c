"""struct mfb_config is bound by hand in two places -- the package's ctypes Structure and the stub INTEGRATION.md gives
a maintainer of the reference.  Both must list the header's fields, in the header's order, with the header's types."""
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mean-field-multi-agent-reinforcement-learning_b200", "python"))


def header_fields():
    text = open(os.path.join(ROOT, "include", "mfmarl_batched.h")).read()
    body = text[text.index("typedef struct mfb_config {"):text.index("} mfb_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("{", 1)[1]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        for name in names.split(","):
            m = re.match(r"\s*(\w+)(\[(\d+)\])?\s*$", name)
            fields.append((m.group(1), ctype, int(m.group(3)) if m.group(3) else 1))
    return fields


def test_package_binding_matches_the_header():
    import ctypes
    from mfmarl_b200.lib import MfbConfig
    kinds = {"int": ctypes.c_int, "unsigned": ctypes.c_uint, "float": ctypes.c_float}
    want = [(n, kinds[t] * k if k > 1 else kinds[t]) for n, t, k in header_fields()]
    got = list(MfbConfig._fields_)
    assert [n for n, _ in got] == [n for n, _ in want]
    for (n, a), (_, b) in zip(got, want):
        assert ctypes.sizeof(a) == ctypes.sizeof(b) and a._type_ == b._type_, n
    assert len(want) >= 28


def test_integration_stub_matches_the_header():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = text[text.index("class _Cfg(ctypes.Structure)"):text.index("def make(")]
    names = re.findall(r'"(\w+)"', stub)
    assert names == [n for n, _, _ in header_fields()]


def header_prototypes():
    """{name: [parameter type strings]} of every `int mfb_*(...)` prototype in include/mfmarl_batched.h"""
    text = open(os.path.join(ROOT, "include", "mfmarl_batched.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(mf[bi]_\w+)\s*\(([^)]*)\)\s*;", text):
        params = [p.strip() for p in m.group(2).split(",")] if m.group(2).strip() not in ("", "void") else []
        protos[m.group(1)] = params
    return protos


def test_package_argtypes_match_the_header_prototypes():
    """every function the package binds: same number of arguments as the header, pointers where the header has
    pointers, and the scalar kinds of the header"""
    import ctypes
    from mfmarl_b200.lib import load_library
    lib = load_library()
    protos = header_prototypes()
    assert len(protos) >= 20 and "mfb_step" in protos and "mfi_run" in protos
    scalar = {"int": ctypes.c_int, "unsigned": ctypes.c_uint, "unsigned long": ctypes.c_ulong, "double": ctypes.c_double}
    checked = 0
    for name, params in protos.items():
        fn = getattr(lib, name)                       # (also: every prototype is exported)
        if fn.argtypes is None:
            continue                                  # bound elsewhere (mfi_* in ising.py, on a GPU box)
        assert len(fn.argtypes) == len(params), (name, len(fn.argtypes), params)
        for at, p in zip(fn.argtypes, params):
            if "*" in p:
                assert at in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(at, "contents") or issubclass(at, ctypes._Pointer), (name, p)
            else:
                kind = " ".join(p.split()[:-1])
                assert at is scalar[kind], (name, p, at)
        checked += 1
    assert checked >= 18
