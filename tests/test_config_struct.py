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
