"""Writes tests/golden/conv_goldens.json: the literal expected gradients of horde-ad's conv2dPreserving tests
over the fixtures t16b / t128b / t128c (test/simplified/TestConvSimplified.hs:196-470), transcribed by parsing the
reference's test source in the build container (the reference is Haskell and cannot be executed here; these
literals -- not outputs of a run -- pin the oracle).  The tests read only the JSON.  Run from the repository root:

    python tests/golden/make_conv_goldens.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
path = os.path.join(ref, "test/simplified/TestConvSimplified.hs")
lines = open(path).read().split("\n")
src = "\n".join(lines)

# which tests are enabled in the reference's test tree (commented-out entries are skipped)
enabled = set(re.findall(r'^\s*[\[,]\s*testCase\s+"[^"]+"\s+(\w+)', src, re.M))

FUNCS = {   # test function -> (fixture that is held constant, which operand the test differentiates)
    "conv2dB": ("t16b", "A"), "conv2dC": ("t16b", "K"),
    "conv2dB128b": ("t128b", "A"), "conv2dC128b": ("t128b", "K"),
    "conv2dB128c": ("t128c", "A"), "conv2dC128c": ("t128c", "K"),
}
pat = re.compile(
    r"^(test\w+) =\n\s+assertEqualUpToEpsilon'? (\S+)\n"
    r"\s+\((?:ringestData|rconcrete \$ Nested\.rfromListPrimLinear) \[([\d, ]+)\] \[([^\]]*)\]\)\n"
    r"\s+\((?:rev' @Double @4|grad \(rsum0 @4 @Double \.) (\w+)\)?\s*\n?\s*\((.*?)\)\)\s*$", re.M | re.S)
out = []
for m in pat.finditer(src):
    name, eps, shape, vals, func, arg = m.groups()
    if name not in enabled or func not in FUNCS:
        continue
    arg = " ".join(arg.split())
    a = re.match(r"(rreplicate0N|rrepl) \[([\d, ]+)\] (\S+)$", arg)
    r = re.match(r"ringestData \[([\d, ]+)\] \[37, 36 \.\. -10\]$", arg)
    if a:
        point = {"kind": "constant", "shape": [int(x) for x in a.group(2).split(",")], "value": float(a.group(3))}
    elif r:
        point = {"kind": "countdown_37_to_-10", "shape": [int(x) for x in r.group(1).split(",")]}
    else:
        continue
    line = src[:m.start()].count("\n") + 1
    fixture, wrt = FUNCS[func]
    out.append({"name": name, "cite": f"test/simplified/TestConvSimplified.hs:{line}", "function": func,
                "fixture": fixture, "wrt": wrt, "point": point, "eps": float(eps),
                "expected_shape": [int(x) for x in shape.split(",")],
                "expected": [float(x) for x in vals.split(",")]})
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conv_goldens.json")
json.dump({"source": "horde-ad v0.4.0.0, test/simplified/TestConvSimplified.hs (literals)", "tests": out}, open(dst, "w"), indent=1)
print(f"wrote {dst}: {len(out)} tests:", ", ".join(t["name"] for t in out))
