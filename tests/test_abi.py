"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/obmd_b200.h declares, and fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "obmd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(obmd_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import obmd_b200 as ob
    L = C.CDLL(ob.lib_path())
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    from importlib import import_module
    binding = import_module("obmd-lammps-extension_b200.binding")
    assert set(names) == set(binding.SIGNATURES), set(names) ^ set(binding.SIGNATURES)
    assert L.obmd_abi_version() == ob.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    import obmd_b200 as ob
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ob.ObmdError) as ei:
        ob.ObmdB200(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "obmd-lammps-extension_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dp, f)).read()
                assert "oracle" not in text.lower().replace("no oracle", ""), os.path.join(dp, f)


def test_parse_fix_obmd_matches_reference_syntax():
    import obmd_b200 as ob
    box_lo, box_hi = (0.0, 0.0, 0.0), (33.59462486002239, 11.198208286674133, 11.198208286674133)
    regs = dict(zip(["leftB", "rightB", "leftshear", "rightshear", "leftBin", "rightBin"],
                    ob.default_regions(box_lo, box_hi, 5.0391)))
    line = ("1 1 6111 188.0 0.0 0.0 0.0 0.0 0.7 0.005 1327 region1 leftB region2 rightB region3 leftshear "
            "region4 rightshear region5 leftBin region6 rightBin buffersize 5.0391 gfac 0.25 "
            "stepparallel 0 stepperp 1 maxattempt 1 usher 1 31.03 1.0 0.02 10000 1.5 1.0 40 charged 0")
    a = ob.parse_fix_obmd(line.split(), regs, box_lo, box_hi)
    assert (a.ntype, a.nfreq, a.seed) == (1, 1, 6111)
    assert (a.pxx, a.alpha, a.tau, a.nbuf) == (188.0, 0.7, 0.005, 1327.0)
    assert (a.etarget, a.ds0, a.uovlp, a.dsovlp, a.eps, a.nattempt) == (31.03, 1.0, 10000.0, 1.5, 1.0, 40)
    assert a.buffer_size == 5.0391 and a.region[1].lo[0] == box_hi[0] - 5.0391
    # error behaviour follows the reference's messages
    for bad, msg in ((line.replace("stepparallel 0", "stepparallel 1"), "smooth transition"),
                     (line.replace("stepperp 1", "stepperp 0"), "heaviside"),
                     (line.replace(" usher 1 31.03 1.0 0.02 10000 1.5 1.0 40", ""), "either usher or near"),
                     (line.replace("6111", "0", 1), "Illegal fix obmd"),
                     (line.replace("region1 leftB", "region1 nope"), "does not exist"),
                     ("1 1 6111 188.0", "check # of arguments")):
        with pytest.raises(ob.ObmdError) as ei:
            ob.parse_fix_obmd(bad.split(), regs, box_lo, box_hi)
        assert msg in str(ei.value)
    # defaults (fix_obmd_merged.cpp:1909-1921)
    short = "1 1 6111 188.0 0 0 0 0 0.7 0.005 1327 region1 leftB region2 rightB region3 leftshear " \
            "region4 rightshear region5 leftBin region6 rightBin usher 1 31.03 1.0 0.02 10000 1.5 1.0 40"
    d = ob.parse_fix_obmd(short.split(), regs, box_lo, box_hi)
    assert d.g_fac == 0.25 and d.buffer_size == 0.30 * box_hi[0] and d.maxattempt == 1
