"""The C-ABI boundary (include/hordeb200.h <-> libhordeb200.so <-> ffi.py),
checked without a GPU: the library loads, exports every declared symbol, the
ctypes table binds every one of them, enums agree with the header, and the
product path fails LOUDLY (no CPU fallback) when there is no device."""
import ctypes as C
import re

import pytest

from horde_ad_b200 import ffi


def _header():
    return open(ffi.HEADER_PATH).read()


def test_library_loads_and_exports_every_declared_symbol():
    lib = ffi.lib()
    declared = ffi.header_symbols()
    assert len(declared) >= 80
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in hordeb200.h but not exported"


def test_ctypes_table_matches_header():
    declared = set(ffi.header_symbols())
    bound = set(ffi.PROTOTYPES)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    # argument counts agree with the C declarations
    text = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    for name, argtypes in ffi.PROTOTYPES.items():
        m = re.search(r"\bint\s+" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("void", "") else args.count(",") + 1
        assert n == len(argtypes), (name, n, len(argtypes))


def _enum(name):
    text = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    body = re.search(r"typedef enum " + name + r"\s*\{(.*?)\}", text, flags=re.S).group(1)
    return [x.split("=")[0].strip() for x in body.split(",") if x.strip()]


def test_enums_agree_with_header():
    un = [e for e in _enum("hb_unop") if e != "HB_UNOP_COUNT"]
    assert [e[3:].lower() for e in un] == ffi.UNOPS
    bi = [e for e in _enum("hb_binop") if e != "HB_BINOP_COUNT"]
    assert [e[3:].lower() for e in bi] == ffi.BINOPS
    vm = [e for e in _enum("hb_vmop") if e != "HB_VM_OP_COUNT"]
    assert [e[6:] for e in vm] == ffi.VMOPS
    dts = _enum("hb_dtype")
    assert [d[3:].lower() for d in dts] == ffi.DTYPE_NAMES
    assert C.sizeof(ffi.Insn) == 16


def test_no_cpu_fallback_without_a_device():
    if ffi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ffi.HordeError, match="no CPU fallback"):
        ffi.call("hb_init", 0)
    # every compute entry point refuses to run before hb_init
    out = C.c_void_p()
    with pytest.raises(ffi.HordeError, match="hb_init"):
        ffi.call("hb_iota", ffi.I64, 4, C.byref(out))


def test_product_package_never_imports_the_oracle():
    import os
    root = os.path.dirname(ffi.__file__)
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_generated_haskell_ffi_module_covers_the_header():
    """bindings/haskell/HordeAd/Core/DeviceFFI.hs (tools/gen_haskell_ffi.py) declares one `foreign import ccall`
    per symbol of include/hordeb200.h with the arity of the ctypes prototype (no GHC here: text-level check)."""
    import os
    import re
    from horde_ad_b200 import ffi
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bindings", "haskell", "HordeAd",
                        "Core", "DeviceFFI.hs")
    hs = open(path).read()
    sigs = dict(re.findall(r'"hordeb200.h (hb_[a-z0-9_]+)"\n  c_\w+ :: (.*)', hs))
    assert sorted(sigs) == ffi.header_symbols()
    for name, sig in sigs.items():
        depth, arrows = 0, 0
        for tok in re.split(r"(\(|\)| -> )", sig):
            depth += (tok == "(") - (tok == ")")
            arrows += tok == " -> " and depth == 0
        assert arrows == len(ffi.PROTOTYPES[name]), (name, sig)

