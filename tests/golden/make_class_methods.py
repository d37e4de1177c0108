"""Writes tests/golden/class_methods.json: the method names of the reference's tensor classes -- LetTensor,
ShareTensor, BaseTensor (src/HordeAd/Core/Ops.hs:157-1255) and ConvertTensor (src/HordeAd/Core/ConvertTensor.hs:25-228)
-- parsed out of the reference source.  Run in the build container (needs /root/reference); the JSON is committed."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def methods(path, start_pat):
    lines = open(path).read().split("\n")
    i = [k for k, l in enumerate(lines) if re.match(start_pat, l)][0]
    j = i + 1
    while "where" not in lines[j - 1]:
        j += 1
    names = []
    k = j
    while k < len(lines):
        l = lines[k]
        if l and not l.startswith(" ") and not l.startswith("--") and not l.startswith("{-") and not l.startswith("#"):
            break
        m = re.match(r"^  ([a-z][A-Za-z0-9_']*)\s*(::.*)?$", l)
        if m and m.group(1) not in ("default", "where", "type", "let", "in"):
            if m.group(2):
                names.append(m.group(1))
            else:
                kk = k + 1
                while kk < len(lines) and lines[kk].startswith("    ") and "::" not in lines[kk] and "=" not in lines[kk]:
                    kk += 1
                if kk < len(lines) and lines[kk].lstrip().startswith("::"):
                    names.append(m.group(1))
        # `tsargMin, tsargMax  -- partial` style: several names before one signature
        m2 = re.match(r"^  ([a-z][A-Za-z0-9_']*(?:, [a-z][A-Za-z0-9_']*)+)\s*(--.*)?$", l)
        if m2:
            names.extend(n.strip() for n in m2.group(1).split(","))
        k += 1
    seen = []
    for n in names:
        if n not in seen:
            seen.append(n)
    return seen


ops = os.path.join(REF, "src/HordeAd/Core/Ops.hs")
out = {
    "LetTensor": methods(ops, r"^class LetTensor"),
    "ShareTensor": methods(ops, r"^class ShareTensor"),
    "BaseTensor": methods(ops, r"^class \( Num \(IntOf target\)"),
    "ConvertTensor": methods(os.path.join(REF, "src/HordeAd/Core/ConvertTensor.hs"), r"^class ConvertTensor"),
}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "class_methods.json")
json.dump(out, open(dst, "w"), indent=1)
print({k: len(v) for k, v in out.items()})
