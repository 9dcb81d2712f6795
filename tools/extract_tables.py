#!/usr/bin/env python3
"""Extract the NUMERIC TABLES the hot path needs from the reference tree into
compact binary blobs under pbrt-r3_b200/data/ (data only, no code).

  sobol32.bin   : SOBOL_MATRICES_32  u32[1024*52]   (core/lowdiscrepancy/sobol/sobolmatrices.rs:4)
  vdc.bin       : VDC_SOBOL_MATRICES u64[25*52] ++ VDC_SOBOL_MATRICES_INV u64[26*52] (:6664, :6823)
  cie.bin       : CIE_X, CIE_Y, CIE_Z, CIE_LAMBDA f32[471] each (build/spectrum/cie_data.rs)
  rgb2spect.bin : RGB2SPECT_LAMBDA + 14 curves f32[32] each (build/spectrum/rgb_data.rs)

Run in the build container (needs /root/reference); the blobs are committed.
"""
import re, struct, sys, os
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pbrt-r3_b200", "data")

def const_block(text, name):
    m = re.search(r"pub const %s\s*:[^=]*=\s*\[" % re.escape(name), text)
    assert m, name
    i = m.end()
    depth = 1
    j = i
    while depth:
        c = text[j]
        if c == "[": depth += 1
        elif c == "]": depth -= 1
        j += 1
    body = text[i:j - 1]
    body = re.sub(r"//[^\n]*", "", body)
    return body

def ints(body):
    return [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]

def floats(body):
    return [float(t) for t in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", body)]

sob = open(os.path.join(REF, "src/core/lowdiscrepancy/sobol/sobolmatrices.rs")).read()
m32 = ints(const_block(sob, "SOBOL_MATRICES_32"))
assert len(m32) == 1024 * 52, len(m32)
np.array(m32, dtype=np.uint32).tofile(os.path.join(OUT, "sobol32.bin"))

def rows(body, nrows):
    # nested arrays; rows may be shorter than 52 (implicitly? no: rust needs full) -> parse per row
    out = []
    depth = 0; cur = ""
    for c in body:
        if c == "[":
            depth += 1; cur = ""
        elif c == "]":
            depth -= 1
            r = ints(cur)
            assert len(r) == 52, len(r)
            out.append(r)
        else:
            cur += c
    assert len(out) == nrows, len(out)
    return out

vdc = rows(const_block(sob, "VDC_SOBOL_MATRICES"), 25)
vdci = rows(const_block(sob, "VDC_SOBOL_MATRICES_INV"), 26)
np.array(vdc + vdci, dtype=np.uint64).tofile(os.path.join(OUT, "vdc.bin"))

cie = open(os.path.join(REF, "build/spectrum/cie_data.rs")).read()
arrs = []
for n in ["CIE_X", "CIE_Y", "CIE_Z", "CIE_LAMBDA"]:
    v = floats(const_block(cie, n))
    assert len(v) == 471, (n, len(v))
    arrs.append(v)
np.array(arrs, dtype=np.float32).tofile(os.path.join(OUT, "cie.bin"))

rgb = open(os.path.join(REF, "build/spectrum/rgb_data.rs")).read()
names = ["RGB2SPECT_LAMBDA"] + ["RGBREFL2SPECT_" + c for c in
         ["WHITE", "CYAN", "MAGENTA", "YELLOW", "RED", "GREEN", "BLUE"]] + \
        ["RGBILLUM2SPECT_" + c for c in ["WHITE", "CYAN", "MAGENTA", "YELLOW", "RED", "GREEN", "BLUE"]]
arrs = []
for n in names:
    v = floats(const_block(rgb, n))
    assert len(v) == 32, (n, len(v))
    arrs.append(v)
np.array(arrs, dtype=np.float32).tofile(os.path.join(OUT, "rgb2spect.bin"))
print("ok")
