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


def test_native_save_number_format_matches_julia_rules(tmp_path):
    """ss_save_rows formats Float64 as Julia's string(x) (Base.Ryu.writeshortest: shortest round-trip digits, fixed
    notation for decimal exponents -4..5, d.ddde-7 otherwise): compared with the host mirror's per-cell formatter on
    awkward values, on random doubles of every magnitude, and on a block large enough to go through the threaded path."""
    from simspread_b200.host import _jl_string
    rng = np.random.default_rng(0)
    special = [0.0, -0.0, 1.0, -1.0, 0.5, 0.1, 1 / 3, 2 / 3, 1e-5, 9.999e-5, 1e-4, 123456.0, 999999.9, 1e6, 1.5e6, 1e21, 1e22, 1e23,
               5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, float("nan"), float("inf"), float("-inf"), -99.0,
               0.30000000000000004, 100.0, 1e5, 12345.678, 7e-5, 6.02214076e23, 4.35e-6]
    vals = np.array(special + list(rng.random(200)) + list(10.0 ** rng.uniform(-300, 300, 300) * rng.choice([-1, 1], 300)))
    nq, nt = 4, len(vals) // 4
    vals = vals[:nq * nt]
    yhat = ss.NamedArray(np.asfortranarray(vals.reshape(nq, nt)), ([f"q{i}" for i in range(nq)], [f"t{j}" for j in range(nt)]))
    y = ss.NamedArray((rng.random((nq, nt)) < 0.3).astype(np.int64), yhat.names())
    p = tmp_path / "native.tsv"
    ss.save(str(p), 7, yhat, y)
    lines = p.read_text().splitlines()
    assert len(lines) == nq * nt
    k = 0
    for qi in range(nq):
        for ti in range(nt):
            want = "\t".join(["7", f'"q{qi}"', f'"t{ti}"', _jl_string(float(yhat.array[qi, ti])), str(int(y.array[qi, ti]))])
            assert lines[k] == want, (lines[k], want)
            k += 1
    for line, x in zip(lines, yhat.array.ravel(order="C")):
        tok = line.split("\t")[3]
        back = float(tok.replace("Inf", "inf").replace("NaN", "nan"))
        assert back == x or (back != back and x != x)   # shortest digits round-trip
    # a larger block: several threads, several rounds; fold column = 1-based query index
    nq, nt = 700, 300
    big = ss.NamedArray(np.round(rng.random((nq, nt)), 5), ([f"D{i}" for i in range(nq)], [f"T{j}" for j in range(nt)]))
    lab = ss.NamedArray((rng.random((nq, nt)) < 0.1).astype(float), big.names())
    p2 = tmp_path / "big.tsv"
    ss.save(str(p2), big, lab, delimiter=" ")
    rows = p2.read_text().splitlines()
    assert len(rows) == nq * nt
    for r in (0, 1, nt, 12345, nq * nt - 1):
        qi, ti = divmod(r, nt)
        assert rows[r] == " ".join([str(qi + 1), f'"D{qi}"', f'"T{ti}"', _jl_string(float(big.array[qi, ti])), _jl_string(float(lab.array[qi, ti]))])


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


# ---------------------------------------------------------------------------------------------
# host I/O (no GPU involved): read_namedmatrix / writedlm, reference src/utils.jl
# ---------------------------------------------------------------------------------------------


def test_read_namedmatrix_reference_fixtures(built, kats, tmp_path):
    """test/runtests.jl:8-18 on test/data/mat1..4, through the native reader, against the oracle restatement."""
    from oracle import simspread_oracle as o
    g = kats["read_namedmatrix"]
    for name, exp in g["expect"].items():
        p = tmp_path / name
        p.write_text(g["files"][name])
        got = ss.read_namedmatrix(str(p), rows=exp["rows"], cols=exp["cols"])
        assert np.array_equal(got.array, np.array(g["values"]))
        assert got.names(1) == exp["rownames"] and got.names(2) == exp["colnames"]
        v, r, c = o.read_namedmatrix(g["files"][name], " ", exp["rows"], exp["cols"])
        assert np.array_equal(v, got.array) and r == got.names(1) and c == got.names(2)
    # the tutorial's label file (docs/src/tutorial/fishers-flowers.jl:12)
    p = tmp_path / "iris.classes"
    p.write_text(g["files"]["iris.classes"])
    got = ss.read_namedmatrix(str(p))
    v, r, c = o.read_namedmatrix(g["files"]["iris.classes"])
    assert got.array.shape == (150, 3) and np.array_equal(v, got.array) and r == got.names(1) and c == got.names(2)
    assert got.array.sum() == 150.0


def test_read_namedmatrix_large_random_matches_oracle_bit_for_bit(built, tmp_path):
    from oracle import simspread_oracle as o
    rng = np.random.default_rng(5)
    n, m = 257, 131
    vals = rng.random((n, m)) * 10.0 ** rng.integers(-12, 12, size=(n, m))
    vals[0, 0], vals[1, 1], vals[2, 2], vals[3, 3] = np.inf, -np.inf, 0.0, -0.0
    rn = [f"s{rng.integers(0, 10**6):06d}_{i}" for i in range(n)]  # unsorted names: the reader sorts them
    cn = [f"t{j}" for j in range(m)]                                # "t10" < "t2": string order
    text = "\t".join([""] + cn) + "\n" + "".join("\t".join([rn[i]] + [repr(float(x)) for x in vals[i]]) + "\n" for i in range(n))
    text = text.replace("inf", "Inf")
    p = tmp_path / "big.tsv"
    p.write_text(text)
    got = ss.read_namedmatrix(str(p), "\t")
    v, r, c = o.read_namedmatrix(text, "\t")
    assert r == got.names(1) == sorted(rn) and c == got.names(2) == sorted(cn)
    assert np.array_equal(v.view(np.uint64), got.array.view(np.uint64))  # including the sign of -0.0
    # CRLF line ends, no trailing newline, '+' signs
    p2 = tmp_path / "crlf.txt"
    p2.write_bytes(b" a b\r\nx +1.5 2\r\ny 3e-2 -4")
    got2 = ss.read_namedmatrix(str(p2))
    assert got2.array.tolist() == [[1.5, 2.0], [0.03, -4.0]] and got2.names(1) == ["x", "y"] and got2.names(2) == ["a", "b"]
    # ragged line -> error, not garbage
    p3 = tmp_path / "ragged.txt"
    p3.write_text(" a b\nx 1 2\ny 3\n")
    with pytest.raises(ss.SimSpreadError, match="line 3"):
        ss.read_namedmatrix(str(p3))


def test_writedlm_layout_and_julia_number_format(built, tmp_path):
    """src/utils.jl:6-11: `["" names(M, 2)...; names(M, 1) M]`, cells printed like Julia prints Float64."""
    from oracle import simspread_oracle as o
    from simspread_b200.host import _jl_string
    X = ss.NamedArray(np.array([[0.0, 1.0, 0.5], [1e-5, 1234567.8, -99.0]]), (["s1", "s2"], ["t1", "t2", "t3"]))
    p = tmp_path / "out.tsv"
    ss.writedlm(str(p), X)
    assert p.read_text() == "\tt1\tt2\tt3\ns1\t0.0\t1.0\t0.5\ns2\t1.0e-5\t1.2345678e6\t-99.0\n"
    want = o.namedmatrix2matrix(X.array, X.names(1), X.names(2))
    assert [ln.split("\t") for ln in p.read_text().splitlines()] == [[c if isinstance(c, str) else _jl_string(c) for c in row] for row in want]
    back = ss.read_namedmatrix(str(p), "\t")
    assert np.array_equal(back.array, X.array) and back.names(1) == X.names(1) and back.names(2) == X.names(2)
    for v, s_ in [(100000.0, "100000.0"), (1e6, "1.0e6"), (0.0001, "0.0001"), (0.1 + 0.2, "0.30000000000000004"), (5e-324, "5.0e-324")]:
        assert _jl_string(v) == s_


# ---------------------------------------------------------------------------------------------
# the Julia host layer cannot be executed here (no Julia in the image): static checks of its ccalls
# ---------------------------------------------------------------------------------------------


def _split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_ccalls_match_the_header():
    """Every `ccall((:ss_x, libss), Cint, (argtypes...), args...)` of SimSpreadB200.jl must name a function the header
    declares, with as many argument types (and arguments) as the C prototype has parameters, and with Julia types whose
    width matches the C parameter (pointers <-> Ptr/Cstring/Ref, int32 <-> Cint/Cuint, int64 <-> Int64, double <-> Float64)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    jl = open(os.path.join(root, "simspread.jl_b200", "julia", "SimSpreadB200.jl")).read()
    hdr = open(os.path.join(root, "include", "simspread_b200.h")).read()
    protos = {}
    for m in re.finditer(r"SS_API\s+[\w\s\*]+?\b(ss_[a-z0-9_]+)\(([^;]*?)\);", hdr, flags=re.S):
        params = [p.strip() for p in _split_top_level(m.group(2).replace("\n", " "))]
        protos[m.group(1)] = [] if params == ["void"] else params

    def c_class(p):
        if "*" in p:
            return "ptr"
        if re.search(r"\bdouble\b", p):
            return "f64"
        if re.search(r"\b(int64_t|uint64_t)\b", p):
            return "i64"
        if re.search(r"\b(int32_t|uint32_t|int)\b", p):
            return "i32"
        raise AssertionError(p)

    def jl_class(t):
        t = t.strip()
        if t.startswith("Ptr{") or t.startswith("Ref{") or t in ("Cstring",):
            return "ptr"
        return {"Float64": "f64", "Int64": "i64", "Cint": "i32", "Cuint": "i32", "Int32": "i32", "UInt32": "i32"}[t]

    calls = 0
    for m in re.finditer(r"ccall\(\(:(ss_[a-z0-9_]+),\s*libss\)", jl):
        name = m.group(1)
        assert name in protos, f"{name} is not declared in the header"
        # the full ccall expression: balance parentheses from the opening one
        i = m.start() + len("ccall")
        depth, j = 0, i
        while True:
            depth += jl[j] == "("
            depth -= jl[j] == ")"
            j += 1
            if depth == 0:
                break
        parts = _split_top_level(jl[i + 1:j - 1])
        # parts: (:name, libss) | return type | (argtypes) | args...
        assert parts[1] in ("Cint", "Cstring"), (name, parts[1])
        argtypes = _split_top_level(parts[2].strip()[1:-1]) if parts[2].strip() not in ("()",) else []
        args = parts[3:]
        want = protos[name]
        assert len(argtypes) == len(want), f"{name}: {len(argtypes)} Julia argument types, {len(want)} C parameters"
        assert len(args) == len(want), f"{name}: {len(args)} arguments passed, {len(want)} C parameters"
        for t, p in zip(argtypes, want):
            assert jl_class(t) == c_class(p), f"{name}: Julia type {t} against C parameter `{p}`"
        calls += 1
    assert calls >= 30


# exported generics of the reference on the hot path (src/SimSpread.jl:21-56) with the signatures of src/core.jl,
# src/performance.jl, src/graphs.jl, src/utils.jl: (name, positional parameters, of which optional, keyword names)
_REFERENCE_API = [
    ("split", 2, 0, {"seed"}),                                              # Base.split, src/core.jl:11
    ("cutoff", 3, 1, set()), ("cutoff!", 3, 1, set()),                      # :37, :55, :72, :87
    ("featurize", 3, 1, set()), ("featurize!", 3, 1, set()),                # :106, :129
    ("construct", 3, 0, set()), ("construct", 2, 0, set()), ("construct", 4, 0, set()),   # :148, :217 / :308, :294
    ("spread", 1, 0, set()),                                                # :365-380
    ("predict", 2, 0, {"GPU"}), ("predict", 3, 0, {"GPU"}),                 # :402, :446, forwarder :424-425
    ("clean!", 3, 0, set()),                                                # :478
    ("save", 3, 0, {"delimiter"}), ("save", 4, 0, {"delimiter"}),           # :503, :542
    ("BEDROC", 2, 0, {"rev", "α"}), ("AuROC", 2, 0, set()), ("AuPRC", 2, 0, set()),       # performance.jl:22, :49, :74
    ("f1score", 4, 0, set()), ("mcc", 3, 1, set()), ("mcc", 4, 0, set()), ("accuracy", 4, 0, set()),
    ("balancedaccuracy", 4, 0, set()), ("recall", 4, 0, set()), ("precision", 4, 0, set()),  # :102-296
    ("recallatL", 3, 1, set()), ("recallatL", 4, 1, set()), ("precisionatL", 3, 1, set()), ("precisionatL", 4, 1, set()),
    ("maxperformance", 3, 0, set()), ("meanperformance", 3, 0, set()), ("meanstdperformance", 3, 0, set()),
    ("validity_ratio", 1, 0, set()),                                        # :558
    ("k", 2, 0, set()), ("k", 1, 0, set()),                                 # graphs.jl:9-11
    ("read_namedmatrix", 3, 2, {"rows", "cols"}),                           # utils.jl:50
    ("writedlm", 2, 0, set()), ("writedlm", 3, 0, set()),                   # utils.jl:8-11
]


def _julia_definitions(src):
    """(name, positional, optional, keyword names) of every `function name(...)` / `name(...) = ...` of the file."""
    import re
    defs = []
    for m in re.finditer(r"^(?:function\s+)?((?:Base\.)?[A-Za-z_][\w!]*)\(", src, flags=re.M):
        name = m.group(1).split(".")[-1]
        i, depth = m.end(), 1
        while depth and i < len(src):
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args = src[m.end():i - 1]
        rest = src[i:i + 200]
        is_def = m.group(0).startswith("function") or re.match(r"\s*(where\s*\{[^}]*\}\s*)?=(?!=)", rest)
        if not is_def:
            continue
        pos, _, kws = args.partition(";") if _top_level_semicolon(args) else (args, "", "")
        plist = [a for a in _split_top_level(pos) if a.strip()]
        klist = [a for a in _split_top_level(kws) if a.strip()]
        opt = sum(1 for a in plist if re.search(r"(?<![=<>!])=(?!=)", _strip_braces(a)))
        defs.append((name, len(plist), opt, {re.split(r"[:=]", a.strip())[0].strip().rstrip(".") for a in klist}))
    return defs


def _strip_braces(a):
    out, depth = [], 0
    for ch in a:
        depth += {"{": 1, "}": -1, "(": 1, ")": -1, "[": 1, "]": -1}.get(ch, 0)
        if depth == 0 and ch not in "})]":
            out.append(ch)
    return "".join(out)


def _top_level_semicolon(args):
    depth = 0
    for ch in args:
        depth += {"(": 1, ")": -1, "{": 1, "}": -1, "[": 1, "]": -1}.get(ch, 0)
        if ch == ";" and depth == 0:
            return True
    return False


def test_julia_layer_defines_the_reference_api():
    """Table-driven: every exported generic of the reference on the hot path is defined in SimSpreadB200.jl with the
    same number of positional parameters, the same optional ones, and (at least) the same keyword names."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    jl = open(os.path.join(root, "simspread.jl_b200", "julia", "SimSpreadB200.jl")).read()
    defs = _julia_definitions(jl)
    exported = set(re_findall_exports(jl))
    missing = []
    for name, npos, nopt, kws in _REFERENCE_API:
        ok = any(d[0] == name and d[1] == npos and d[2] >= nopt and kws <= d[3] for d in defs)
        if not ok:
            missing.append((name, npos, nopt, sorted(kws), [d for d in defs if d[0] == name]))
        if name != "split":
            assert name in exported, f"{name} is not exported"
    assert not missing, missing


def re_findall_exports(jl):
    import re
    m = re.search(r"^export (.*?)\n\n", jl, flags=re.S | re.M)
    return [x.strip() for x in m.group(1).replace("\n", " ").split(",") if x.strip()]
