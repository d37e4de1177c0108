"""Extracts the printed gradient ARTIFACTS of the reference's CNN pretty-printing test (testCNNOPP2S,
test/simplified/TestMnistPP.hs:700-735: MnistCnnShaped2.convMnistTwoS on a constant 14x23 glyph batch of 7, kernels
3x4, 4 channels, 5 hidden units) into tests/golden/cnn_artifacts.json.  These strings are GOLDEN PROGRAMS: the literal
op sequences horde-ad's interpreter sends to a backend.  Run in the build container (needs /root/reference)."""
import json
import os
import re

SRC = "/root/reference/test/simplified/TestMnistPP.hs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cnn_artifacts.json")

text = open(SRC).read()
start = text.index("testCNNOPP2S = do")
body = text[start:]
names = ["primal_simplified", "primal", "rev", "rev_simplified"]
found = re.findall(r'@\?= "((?:[^"\\]|\\.)*)"', body)[:4]
assert len(found) == 4
arts = {n: s.replace("\\\\", "\\") for n, s in zip(names, found)}
json.dump({"source": "test/simplified/TestMnistPP.hs:728-735 (testCNNOPP2S)", "shapes": {
    "glyph": [7, 1, 14, 23], "ker1": [4, 1, 3, 4], "bias1": [4], "ker2": [4, 4, 3, 4], "bias2": [4],
    "weightsDense": [5, 60], "biasesDense": [5], "weightsReadout": [10, 5], "biasesReadout": [10]},
    "artifacts": arts}, open(OUT, "w"), indent=1)
print("wrote", OUT, {k: len(v) for k, v in arts.items()})
