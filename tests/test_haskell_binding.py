"""The typed Haskell layer (bindings/haskell: CarriersDevice.hs, OpsDevice.hs, ADEngineDevice.hs) cannot be compiled
in this image (no GHC).  What CAN be checked on the CPU:
  * every method of the reference's tensor classes (LetTensor, ShareTensor, BaseTensor: Ops.hs:157-1255;
    ConvertTensor: ConvertTensor.hs:25-228 -- names in tests/golden/class_methods.json, parsed from the reference by
    tests/golden/make_class_methods.py) is DEFINED in the corresponding `instance ... Device` of OpsDevice.hs;
  * every raw FFI name the typed layer uses exists in the generated DeviceFFI.hs (which test_abi.py ties to the header);
  * the bytecode op codes and element-type codes it hard-wires equal the enums of include/hordeb200.h."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS = os.path.join(ROOT, "bindings", "haskell", "HordeAd")


def read(*p):
    return open(os.path.join(*p)).read()


def instance_body(src, cls):
    m = re.search(r"^instance %s Device where\n" % cls, src, re.M)
    assert m, f"no `instance {cls} Device`"
    rest = src[m.end():]
    end = re.search(r"^(?=\S)", rest, re.M)
    return rest[:end.start()] if end else rest


def test_every_class_method_is_defined_in_the_device_instances():
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "class_methods.json")))
    src = read(HS, "Core", "OpsDevice.hs")
    assert sum(len(v) for v in want.values()) >= 210
    for cls, names in want.items():
        body = instance_body(src, cls)
        missing = [n for n in names
                   if not re.search(r"(^|[;\s])%s(\s+[^=\n;]*)?\s=(?!=)" % re.escape(n), body, re.M)]
        assert not missing, f"instance {cls} Device lacks {missing}"


def test_typed_layer_only_uses_ffi_names_that_exist():
    ffi = set(re.findall(r"^\s+(c_hb_\w+|p_hb_\w+) ::", read(HS, "Core", "DeviceFFI.hs"), re.M))
    assert len(ffi) >= 100
    for rel in (("Core", "CarriersDevice.hs"), ("Core", "OpsDevice.hs"), ("ADEngineDevice.hs",)):
        used = set(re.findall(r"\b(c_hb_\w+|p_hb_\w+)\b", read(HS, *rel)))
        assert used, rel
        assert used <= ffi, (rel, sorted(used - ffi))


def test_hardwired_codes_match_the_header():
    hdr = read(ROOT, "include", "hordeb200.h")

    def enum(name):
        body = re.search(r"typedef enum %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out, nxt = {}, 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = [x.strip() for x in item.split("=")]
                nxt = int(v)
            else:
                k = item
            out[k] = nxt
            nxt += 1
        return out
    vm = enum("hb_vmop")
    ops = read(HS, "Core", "OpsDevice.hs")
    for name, val in re.findall(r"\bvm([A-Z0-9]+) = (\d+)", ops):
        assert vm["HB_VM_" + name] == int(val), name
    un, bi, dt = enum("hb_unop"), enum("hb_binop"), enum("hb_dtype")
    # the op codes used by the Num / Floating instances
    assert (un["HB_NEGATE"], un["HB_ABS"], un["HB_SIGNUM"], un["HB_RECIP"], un["HB_EXP"], un["HB_ATANH"]) == (0, 1, 2, 3, 4, 18)
    assert (bi["HB_ADD"], bi["HB_SUB"], bi["HB_MUL"], bi["HB_DIVIDE"], bi["HB_POWER"], bi["HB_LOGBASE"], bi["HB_ATAN2"],
            bi["HB_QUOT"], bi["HB_REM"]) == (0, 1, 2, 3, 4, 5, 6, 7, 8)
    car = read(HS, "Core", "CarriersDevice.hs")
    for ty, code in (("Double", "HB_F64"), ("Float", "HB_F32"), ("Int64", "HB_I64"), ("Int32", "HB_I32"),
                     ("Int16", "HB_I16"), ("Int8", "HB_I8"), ("Bool", "HB_BOOL")):
        m = re.search(r"instance HbElt %s\s+where \{ hbDtype = (\d+)" % ty, car)
        assert m and int(m.group(1)) == dt[code], ty
    # struct hb_insn is 16 bytes with the immediate at offset 8 (pokeInsn)
    assert re.search(r"uint8_t op;.*uint8_t dst;.*uint8_t a, b, c;.*uint8_t n;.*uint8_t pad\[2\];", hdr, re.S)
