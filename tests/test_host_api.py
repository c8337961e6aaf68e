"""Host-side mirror of the reference API (no GPU needed): option types, validation order/messages,
Result semantics, and that the C-ABI library loads and exports every symbol include/nmfb200.h declares."""
import os
import re
import subprocess
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound(NMF):
    hdr = open(os.path.join(ROOT, "include", "nmfb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nmfb200_[A-Za-z0-9_]+)\s*\(", hdr)) - {"nmfb200_trace_fn"}
    assert len(declared) >= 27
    NMF.build.build_library()
    lib = NMF._lib.load()  # resolves every bound symbol or raises
    assert declared == set(NMF._lib.SIGNATURES), (declared ^ set(NMF._lib.SIGNATURES))
    out = subprocess.check_output(["nm", "-D", "--defined-only", NMF._lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (nmfb200_[A-Za-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert lib.nmfb200_version() == 100
    assert lib.nmfb200_status_string(1) == b"invalid argument"
    assert lib.nmfb200_status_string(2) == b"dimension mismatch"


def test_result_struct_matches_header(NMF):
    import ctypes
    # int64, int32, int32, 4 doubles, 2 int64, double, int64, int64, double -> 96 bytes, no padding surprises
    assert ctypes.sizeof(NMF._lib.NmfResult) == 8 + 4 + 4 + 8 * 4 + 8 * 2 + 8 + 8 + 8 + 8


def test_library_built_for_sm100a_only(NMF):
    NMF.build.build_library()
    out = subprocess.check_output(["cuobjdump", "-lelf", NMF._lib.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_multupdate_ctor(NMF):
    a = NMF.MultUpdate(np.float32)
    assert (a.obj, a.maxiter, a.verbose, a.update_H) == ("mse", 100, False, True)
    assert a.tol == np.float32(np.cbrt(np.finfo(np.float32).eps))
    assert a.lambda_w == 0 and a.lambda_h == 0
    for bad in (dict(obj="kl"), dict(maxiter=1), dict(tol=0), dict(lambda_w=-1e-3), dict(lambda_h=-1e-3)):
        with pytest.raises(NMF.ArgumentError):
            NMF.MultUpdate(np.float64, **bad)
    d = NMF.MultUpdate(np.float64, obj="div", lambda_w=1e-3)
    assert d.lambda_w == 1e-3 and d.lambda_h == np.sqrt(np.finfo(np.float64).eps)
    with pytest.warns(UserWarning, match="deprecated"):
        l = NMF.MultUpdate(np.float64, lambda_=0.5, lambda_h=0.25)
    assert l.lambda_w == 0.5 and l.lambda_h == 0.25


def test_greedycd_ctor(NMF):
    a = NMF.GreedyCD(np.float64)
    assert (a.maxiter, a.verbose, a.update_H, a.lambda_w, a.lambda_h) == (100, False, True, 0, 0)
    assert a.tol == np.cbrt(np.finfo(np.float64).eps)
    for bad in (dict(maxiter=1), dict(tol=0), dict(lambda_w=-1), dict(lambda_h=-1)):
        with pytest.raises(NMF.ArgumentError):
            NMF.GreedyCD(np.float32, **bad)


def test_other_algorithm_types_exist(NMF):
    assert NMF.ProjectedALS(np.float32).lambda_w == np.float32(np.cbrt(np.finfo(np.float32).eps))
    assert NMF.ALSPGrad(np.float64).maxsubiter == 200
    assert NMF.CoordinateDescent(np.float64, alpha=1e-4, l1ratio=0.5, shuffle=True).shuffle
    assert NMF.SPA(np.float64).obj == "mse"
    with pytest.raises(NMF.ArgumentError):
        NMF.SPA(np.float64, obj="x")


def test_result_eq_hash(NMF):  # test/utils.jl:65-69
    W, H = np.ones((3, 2)), np.ones((2, 4))
    a, b = NMF.Result(W, H, 3, True, 0.5), NMF.Result(W.copy(), H.copy(), 3, True, 0.5)
    assert a == b and hash(a) == hash(b)
    assert a != NMF.Result(W, H, 4, True, 0.5)
    with pytest.raises(NMF.DimensionMismatch):
        NMF.Result(np.ones((3, 2)), np.ones((3, 4)), 1, False, 0.0)
    assert isinstance(NMF.Result(W.astype(np.float32), H.astype(np.float32), 1, False, 0.1).objvalue, np.float32)


def test_nnmf_validation_before_gpu(NMF):
    """interf.jl:15-36, :55, :74-79 -- all raised before any device work."""
    X = np.random.default_rng(0).random((6, 5))
    with pytest.raises(NMF.ArgumentError, match="non-negative"):
        NMF.nnmf(X - 1.0, 2, alg="multmse", init="random")
    with pytest.raises(NMF.ArgumentError, match="should not exceed"):
        NMF.nnmf(X, 6, alg="multmse", init="random")
    with pytest.raises(NMF.ArgumentError, match="replicates"):
        NMF.nnmf(X, 2, alg="multmse", init="random", replicates=0)
    with pytest.raises(NMF.ArgumentError, match="set W0 and H0"):
        NMF.nnmf(X, 2, alg="multmse", init="custom")
    W0, H0 = np.ones((6, 2)), np.ones((2, 5))
    with pytest.raises(NMF.ArgumentError, match="W0 must be non-negative"):
        NMF.nnmf(X, 2, alg="multmse", init="custom", W0=-W0, H0=H0)
    with pytest.raises(NMF.ArgumentError, match="Invalid size for W0"):
        NMF.nnmf(X, 2, alg="multmse", init="custom", W0=np.ones((5, 2)), H0=H0)
    with pytest.raises(NMF.ArgumentError, match="H0 must be non-negative"):
        NMF.nnmf(X, 2, alg="multmse", init="custom", W0=W0, H0=-H0)
    with pytest.raises(NMF.ArgumentError, match="Invalid size for H0"):
        NMF.nnmf(X, 2, alg="multmse", init="custom", W0=W0, H0=np.ones((2, 4)))
    with pytest.raises(NMF.ArgumentError, match="Invalid value for init"):
        NMF.nnmf(X, 2, alg="multmse", init="bogus")
    with pytest.raises(NMF.ArgumentError, match="Invalid algorithm"):
        NMF.nnmf(X, 2, alg="bogus", init="random")
    with pytest.raises(NMF.ArgumentError, match="use :spa instead"):
        NMF.nnmf(X, 2, alg="spa", init="random")
    with pytest.raises(NMF.ArgumentError, match="maxiter must be greater than 1"):
        NMF.nnmf(X, 2, alg="multmse", init="random", maxiter=1)
    with pytest.raises(NMF.ArgumentError, match="regularization"):
        NMF.CoordinateDescent(np.float64, regularization="bogus")
    with pytest.raises(NotImplementedError):
        NMF.nnmf(X, 2, init="spa")  # SPA initialisation is out of scope
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(NMF.ArgumentError):
            NMF.nnmf(X, 2, alg="bogus", init="random", W0=W0, update_H=False)
    msgs = " ".join(str(x.message) for x in w)
    assert "Only W will be updated." in msgs and "Ignore W0 and H0" in msgs


def test_randinit(NMF):  # test/initialization.jl:4-27
    W, H = NMF.randinit(10, 7, 3, np.float32, normalize=True, rng=np.random.default_rng(1))
    assert W.dtype == np.float32 and W.flags.f_contiguous and H.flags.f_contiguous
    np.testing.assert_allclose(W.sum(axis=0), 1.0, rtol=1e-6)
    W, H = NMF.randinit(10, 7, 3, np.float64, zeroh=True)
    assert (H == 0).all() and H.shape == (3, 7)


def test_nndsvd_with_initdata_matches_oracle_and_reference_properties(NMF, oracle):
    """test/initialization.jl:29-53 on the host mirror (initdata given => no device work) and equality with the oracle."""
    rng = np.random.default_rng(5678)
    X = rng.random((8, 12))
    U, s, Vt = np.linalg.svd(X, full_matrices=False)
    F = (U, s, Vt.T)
    for T in (np.float64, np.float32):
        Xt = X.astype(T)
        for variant in ("std", "a", "ar"):
            W, H = NMF.nndsvd(Xt, 5, variant=variant, initdata=F, rng=np.random.default_rng(3))
            Wo, Ho = oracle.nndsvd(Xt, 5, variant=variant, initdata=F, rng=np.random.default_rng(3))
            assert W.shape == (8, 5) and H.shape == (5, 12) and W.dtype == T and W.flags.f_contiguous and H.flags.f_contiguous
            assert (W >= 0).all() and (H >= 0).all()
            np.testing.assert_allclose(W, Wo, rtol=1e-6 if T == np.float32 else 1e-13, atol=0)
            np.testing.assert_allclose(H, Ho, rtol=1e-6 if T == np.float32 else 1e-13, atol=0)
        W2, H2 = NMF.nndsvd(Xt, 5, zeroh=True, initdata=F)
        W1, H1 = NMF.nndsvd(Xt, 5, initdata=F)
        assert (W2 == W1).all() and (H2 == 0).all()                      # test/initialization.jl:39-44
        U2, s2, Vt2 = np.linalg.svd(2 * X, full_matrices=False)
        Wb, Hb = NMF.nndsvd(2 * Xt, 5, initdata=(U2, s2, Vt2.T))
        np.testing.assert_allclose(Wb, np.sqrt(T(2)) * W1, rtol=1e-5)    # :46-49
        np.testing.assert_allclose(Hb, np.sqrt(T(2)) * H1, rtol=1e-5)
        War, _ = NMF.nndsvd(Xt, 5, variant="ar", initdata=F)
        assert (War > 0).all()                                           # :51-52
    with pytest.raises(NMF.ArgumentError, match="variant"):
        NMF.nndsvd(X, 5, variant="bogus", initdata=F)


def test_nndsvd_matches_golden_fixtures(NMF, oracle):
    """tests/golden/init/*.npz (oracle outputs for a stored SVD, generator: tests/golden/make_golden.py)."""
    import glob
    import os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "init", "*.npz")))
    assert len(files) >= 4
    for f in files:
        g = np.load(f)
        X, k = g["X"], int(g["k"])
        tol = 1e-6 if X.dtype == np.float32 else 1e-13
        for mod in (NMF, oracle):
            W, H = mod.nndsvd(X, k, zeroh=bool(g["zeroh"]), variant=str(g["variant"]), initdata=(g["U"], g["S"], g["V"]),
                              rng=np.random.default_rng(int(g["rng_seed"])))
            np.testing.assert_allclose(W, g["W"], rtol=tol, atol=0)
            np.testing.assert_allclose(H, g["H"], rtol=tol, atol=0)


def test_no_gpu_fails_loudly(NMF):
    """Without a CUDA device the product must raise, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X = np.random.default_rng(0).random((6, 5))
    with pytest.raises(NMF.NmfB200Error):
        NMF.nnmf(X, 2, alg="multmse", init="random")


# ---- include/nmfb200.h  <->  nmf.jl_b200/julia/NMFB200.jl: every ccall's (return type, argument-type tuple) ----------
_C2JL = {
    "nmfb200_handle**": "Ref{Ptr{Cvoid}}", "nmfb200_handle*": "Ptr{Cvoid}", "const nmfb200_handle*": "Ptr{Cvoid}",
    "void*": "Ptr{Cvoid}", "const void*": "Ptr{Cvoid}", "const char*": "Cstring", "int": "Cint", "int64_t": "Int64",
    "uint64_t": "UInt64", "float": "Float32", "double": "Float64", "float*": "Ptr{Float32}", "const float*": "Ptr{Float32}",
    "double*": "Ptr{Float64}", "const double*": "Ptr{Float64}", "nmfb200_result*": "Ref{CResult}", "nmfb200_trace_fn": "Ptr{Cvoid}",
    "int64_t*": "Ref{Int64}", "int32_t": "Int32", "const int64_t*": "Ptr{Int64}",
}


def _header_prototypes():
    hdr = open(os.path.join(ROOT, "include", "nmfb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|const char\*)\s+(nmfb200_[A-Za-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        types = []
        for a in [x.strip() for x in args.split(",")]:
            if a in ("void", ""):
                continue
            m = re.match(r"^(.*?)(\b[A-Za-z_][A-Za-z0-9_]*)$", a)          # strip the parameter name
            t = re.sub(r"\s*\*", "*", m.group(1).strip())
            types.append(_C2JL[t])
        protos[name] = (_C2JL[ret.replace(" *", "*")] if ret != "int" else "Cint", tuple(types))
    return protos


def _julia_ccalls():
    """Expand the two `for (T, sfx)` loops and the inner `for (alg, name)` loop of NMFB200.jl textually and return
    [(symbol, return type, (argument types...))] for every ccall in the file."""
    src = open(os.path.join(ROOT, "nmf.jl_b200", "julia", "NMFB200.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    pat = re.compile(r"ccall\(\(\s*([^,]+?)\s*,\s*libnmfb200\)\s*,\s*(\w+)\s*,\s*\(([^()]*(?:\{[^{}]*(?:\{[^{}]*\}[^{}]*)*\}[^()]*)*)\)", re.S)
    # variable -> name template, e.g. setx -> "nmfb200_set_X_{sfx}", cname -> "nmfb200_solve_{name}_{sfx}"
    templ = {}
    for var, parts in re.findall(r"(\w+)\s*=\s*Symbol\(([^)]*)\)", src):
        templ[var] = "".join(p.strip().strip('"') if p.strip().startswith('"') else "{" + p.strip() + "}" for p in parts.split(","))
    for var, lit, v2 in re.findall(r"(\w+)\s*=\s*\"(nmfb200_[A-Za-z0-9_]*)\"\s*\*\s*(\w+)", src):
        templ[var] = lit + "{" + v2 + "}"
    names = re.search(r"for \(alg, name\) in \(([^\n]*)\)\n", src).group(1)
    alg_names = re.findall(r'"(\w+)"', names)
    for sym, ret, args in pat.findall(src):
        sym = sym.strip()
        targs = [a.strip() for a in re.split(r",(?![^{]*\})", args) if a.strip()]
        if sym.startswith(":"):
            calls.append((sym[1:], ret, tuple(targs)))
            continue
        var = re.sub(r"[$()]|QuoteNode", "", sym)
        assert var in templ, f"cannot resolve ccall symbol {sym}"
        for T, sfx in (("Float32", "f32"), ("Float64", "f64")):
            for nm in (alg_names if "{name}" in templ[var] else [None]):
                full = templ[var].replace("{sfx}", sfx).replace("{name}", nm or "")
                calls.append((full, ret, tuple(a.replace("$T", T) for a in targs)))
    return calls


def test_julia_ccall_signatures_match_header():
    protos = _header_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 22
    seen = set()
    for name, ret, args in calls:
        assert name in protos, f"NMFB200.jl ccalls {name}, which include/nmfb200.h does not declare"
        assert (ret, args) == protos[name], f"{name}: NMFB200.jl has {ret} {args}, header says {protos[name]}"
        seen.add(name)
    # every solve / set_X / mul_X entry point of the header is bound by the Julia wrapper
    must = {n for n in protos if n.startswith(("nmfb200_solve_", "nmfb200_set_X_f", "nmfb200_mul_X_"))}
    assert must <= seen, must - seen


def test_julia_status_codes_match_header():
    hdr = open(os.path.join(ROOT, "include", "nmfb200.h")).read()
    codes = re.findall(r"(NMFB200_[A-Z]+)\s*=\s*(\d+)", hdr)
    jl = open(os.path.join(ROOT, "nmf.jl_b200", "julia", "NMFB200.jl")).read()
    m = re.search(r"const ([A-Z, ]+) = 0:(\d+)", jl)
    names = [x.strip() for x in m.group(1).split(",")]
    assert [c[0].replace("NMFB200_", "") for c in sorted(codes, key=lambda c: int(c[1]))] == names and int(m.group(2)) == len(names) - 1


# ---- solve_replicates (interf.jl:85-101): grouping into stacked solves, draw order, fallback -- host logic, no GPU -----------------
class _FakeReplicateSession:
    """Duck-typed Session: records the initial factors it is handed; the 'objective' of a solve is the sum of its initial W."""

    def __init__(self, p, n, batched_ok=True):
        self.shape, self.batched_ok, self.calls, self.seen = (p, n), batched_ok, [], []

    def _result(self, NMF, W, H, extra=None):
        self.seen.append(W.copy())
        return NMF.Result(W, H, 3, False, float(W.sum()), extra or {})

    def solve(self, alg, W, H):
        import nmf_jl_b200 as NMF
        self.calls.append(1)
        return self._result(NMF, W, H)

    def solve_batched(self, alg, Ws, Hs):
        import nmf_jl_b200 as NMF
        if not self.batched_ok:
            raise NotImplementedError("not covered")
        self.calls.append(len(Ws))
        return [self._result(NMF, W, H, {"batched": len(Ws)}) for W, H in zip(Ws, Hs)]


def test_solve_replicates_groups_draw_order_and_fallback(NMF):
    p, n = 40, 30
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=5)

    def run(k, replicates, batched, batched_ok=True, algo=alg):
        rng = np.random.default_rng(99)
        W, H = NMF.randinit(p, n, k, algo.T, normalize=True, rng=rng)
        s = _FakeReplicateSession(p, n, batched_ok)
        r = NMF.solve_replicates(algo, s, W, H, replicates=replicates, initH=True, rng=rng, batched=batched)
        return s, r

    one, r1 = run(100, 7, batched=False)
    grp, r2 = run(100, 7, batched=True)
    assert one.calls == [1] * 7 and grp.calls == [2, 2, 2, 1]               # groups of 256 // 100 replicates, the rest one by one
    assert all((a == b).all() for a, b in zip(one.seen, grp.seen))           # the restarts are drawn in the reference's order
    assert (r1.W == r2.W).all() and r1.objvalue == r2.objvalue               # the same replicate wins (first smallest objvalue)
    assert float(r1.objvalue) == min(float(w.sum(dtype=np.float32)) for w in one.seen)
    small, _ = run(8, 40, batched=True)
    assert small.calls == [32, 8]                                            # at most 32 replicates per stacked solve
    fb, r3 = run(100, 5, batched=True, batched_ok=False)                     # library says ENOTSUP: the same factors, one by one
    assert fb.calls == [1] * 5 and (r3.W == run(100, 5, batched=False)[1].W).all()
    for other in (NMF.MultUpdate(np.float32, obj="div", maxiter=5), NMF.MultUpdate(np.float64, maxiter=5), NMF.GreedyCD(np.float32, maxiter=5)):
        s, _ = run(8, 4, batched=True, algo=other)                           # only Float32 MultUpdate(:mse) is stacked
        assert s.calls == [1] * 4


def test_nnmf_sparse_validation_before_gpu(NMF):
    sp = pytest.importorskip("scipy.sparse")
    X = sp.random(12, 9, density=0.4, format="csr", dtype=np.float64, random_state=np.random.default_rng(0))
    assert NMF.api._is_sparse(X) and not NMF.api._is_sparse(X.toarray())
    bad = X.copy()
    bad.data[0] = -1.0
    with pytest.raises(NMF.ArgumentError, match="non-negative"):             # interf.jl:15 on the stored entries
        NMF.nnmf(bad, 3)
    with pytest.raises(NMF.ArgumentError, match="should not exceed"):        # interf.jl:18
        NMF.nnmf(X, 10)
    with pytest.raises(NMF.ArgumentError, match="eltype"):
        NMF.nnmf(X.astype(np.int32), 3)
