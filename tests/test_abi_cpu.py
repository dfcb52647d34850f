"""CPU-side checks of the drop-in boundary: the shared library builds/loads, exports every symbol
the public header declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import simspread_b200 as ss
from simspread_b200 import _lib


@pytest.fixture(scope="module")
def built():
    path = ss.build()
    assert os.path.exists(path)
    return path


def test_header_and_binding_agree(built):
    hdr = ss.header_symbols()
    assert len(hdr) == len(set(hdr)) >= 40
    assert set(hdr) == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in ss.header_symbols() if s not in exported]
    assert not missing, missing
    # nothing but the ABI leaks out of the library
    assert all(s.startswith("ss_") for s in exported), sorted(exported)[:10]
    h = ss.lib()
    for s in ss.header_symbols():
        assert getattr(h, s) is not None
    assert h.ss_version() == 100


def test_sass_is_blackwell_native(built):
    """The chain-product kernel must carry the FP64 tensor pipe (DMMA) and TMA (UTMALDG) in SASS."""
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert sass.count("DMMA.8x8x4") >= 256
    assert "UTMALDG.2D" in sass
    assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass  # mbarrier try_wait


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_cpu_fallback(built):
    n = C.c_int32(-1)
    assert ss.lib().ss_device_count(C.byref(n)) == 0 and n.value == 0
    with pytest.raises(ss.SimSpreadError) as e:
        ss.Context()
    assert e.value.status == _lib.SS_ERR_NO_DEVICE
    with pytest.raises(ss.SimSpreadError):
        ss.cutoff(np.zeros((2, 2)), 0.5)


def test_host_bookkeeping_without_gpu():
    """Name logic of construct / split / NamedArray is host code and must match the reference."""
    X = ss.NamedArray(np.array([[1., 0, 1], [1, 1, 0], [0, 1, 1]]), (["s1", "s2", "s3"], ["s1", "s2", "s3"]))
    y = ss.NamedArray(np.array([[0., 1], [1, 1], [1, 0]]), (["s1", "s2", "s3"], ["t1", "t2"]))
    with pytest.raises(AssertionError, match="Source and Features nodes have the same names!"):
        ss.construct(y, X, ["s1"])
    X2 = ss.NamedArray(np.ones((2, 3)), (["s1", "s2"], ["fs1", "fs2", "fs3"]))
    with pytest.raises(AssertionError, match="Labels and features have different number of source nodes"):
        ss.construct(y, X2, ["s1"])
    folds = ss.split(y, 2, seed=3)
    assert sorted(sum(folds, [])) == ["s1", "s2", "s3"]
    assert [len(f) for f in folds] == [1, 2]  # i = 1,2,3 -> folds mod(i,2)+1 = 2,1,2
    sub = X[["s3", "s1"], ["s2"]]
    assert sub.names(1) == ["s3", "s1"] and sub.array.tolist() == [[1.0], [0.0]]
    assert ss.f1score(3, 2, 2, 3) == pytest.approx(0.6)
    assert ss.mcc(3, 2, 2, 3) == pytest.approx(0.2)
    assert ss.accuracy(3, 2, 2, 3) == pytest.approx(0.6)
    assert ss.balancedaccuracy(3, 2, 2, 3) == pytest.approx(0.6)
    assert ss.recall(3, 2, 2, 3) == pytest.approx(0.6)
    assert ss.precision(3, 2, 2, 3) == pytest.approx(0.6)
    assert ss.mcc(0, 5, 0, 5) - ss.mcc(5, 5) < 1e-5


def test_save_matches_reference_fixtures(tmp_path, kats):
    """`save` is host-only I/O: byte-exact against test/data/save1..4 (test/runtests.jl:185-203)."""
    sv = kats["save"]
    y = ss.NamedArray(np.array(sv["y"]), (sv["rows"], sv["cols"]))
    yhat = y.copy()
    for name, args, kw in (("save1", (y, yhat), {}), ("save2", (y, yhat), {"delimiter": " "}),
                           ("save3", (1, y, yhat), {}), ("save4", (1, y, yhat), {"delimiter": " "})):
        p = tmp_path / name
        ss.save(str(p), *args, **kw)
        assert p.read_text() == sv["files"][name]
    ss.save(str(tmp_path / "save1"), y, yhat)  # "a+": appends
    assert (tmp_path / "save1").read_text() == sv["files"]["save1"] * 2
