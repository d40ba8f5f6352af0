"""The C-ABI library loads and exports every symbol include/deer_engine.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "deer_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(deer_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__ as g
    g.build()
    from deer_vla_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 24
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/deer_engine.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes binding"
    assert lib.deer_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.deer_build_info()


def test_struct_layouts_match_header():
    from deer_vla_b200 import _lib
    src = open(os.path.join(ROOT, "include", "deer_engine.h")).read()
    body = re.search(r"typedef struct deer_config \{(.*?)\} deer_config;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl.startswith("int32_t"):
            fields += [f.strip() for f in decl[len("int32_t"):].split(",")]
    assert fields == [f for f, _ in _lib.DeerCConfig._fields_]
    assert ctypes.sizeof(_lib.DeerCConfig) == 4 * len(fields)


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "deer_vla_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn
