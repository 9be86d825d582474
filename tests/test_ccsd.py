"""pycc_b200 (fused residual, solve_cc, DIIS, (T), the reference-API building blocks) against the
reference's golden vectors (tests/golden/ref_*.npz, outputs of the reference's own code).

Every test runs twice: `emu` -- host-side logic through the numpy double of the C ABI (tests/emu.py;
`-m "not gpu"`, the build container has no GPU) -- and `cuda` -- the real kernels through
libb200cc.so on a B200 (`-m gpu`, the parity tests proper).  FP64 tolerances: 1e-10 Eh on energies,
1e-9 max-abs on amplitudes (north_star); most checks are far tighter."""
import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200 import cctriples
from pycc_b200.synthetic import full_eri
from pycc_b200.wavefunction import IntegralReference
from tests import emu


DEV = [torch.device("cpu")]


def T(x):
    return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True)).to(DEV[0])


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    """'emu': host logic on CPU tensors through the numpy double of the C ABI (no GPU in the build
    container).  'cuda': the real libb200cc.so kernels on a B200 -- the parity tests proper."""
    if request.param == "emu":
        DEV[0] = torch.device("cpu")
        with emu.install():
            yield DEV[0]
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        DEV[0] = torch.device("cuda:0")
        yield DEV[0]
        DEV[0] = torch.device("cpu")


def make_wfn(syn, model="CCSD", from_factor=True):
    ref = IntegralReference.from_synthetic(syn) if from_factor else \
        IntegralReference.from_arrays(syn.F, full_eri(syn), syn.no)
    return pycc_b200.ccwfn(ref, model=model, device='GPU', quiet=True)


def test_block_views_match_full_eri(golden, dev):
    g, syn = golden
    cc = make_wfn(syn)
    ERI = full_eri(syn)
    L = 2.0 * ERI - ERI.swapaxes(2, 3)
    o, v = cc.o, cc.v
    for pat in ("oooo", "ooov", "oovo", "ovoo", "vooo", "oovv", "vvoo", "ovov", "vovo", "ovvo", "voov",
                "ovvv", "vovv", "vvov", "vvvo", "vvvv"):
        key = tuple(o if c == "o" else v for c in pat)
        assert np.abs(cc.H.ERI[key].cpu().numpy() - ERI[key]).max() < 1e-13, pat
    for pat in ("oovv", "ovvv", "ooov", "ovvo", "oovo"):
        key = tuple(o if c == "o" else v for c in pat)
        assert np.abs(cc.H.L[key].cpu().numpy() - L[key]).max() < 1e-13, pat


@pytest.mark.parametrize("from_factor", [True, False])
def test_residuals_and_intermediates(golden, dev, from_factor):
    g, syn = golden
    cc = make_wfn(syn, from_factor=from_factor)
    o, v, H = cc.o, cc.v, cc.H
    t1, t2 = T(g["rand_t1"]), T(g["rand_t2"])
    F = H.F
    tol = 1e-12
    assert np.abs(cc.build_Fae(o, v, F, H.L, t1, t2).cpu().numpy() - g["rand_Fae"]).max() < tol
    assert np.abs(cc.build_Fmi(o, v, F, H.L, t1, t2).cpu().numpy() - g["rand_Fmi"]).max() < tol
    assert np.abs(cc.build_Fme(o, v, F, H.L, t1).cpu().numpy() - g["rand_Fme"]).max() < tol
    assert np.abs(cc.build_Wmnij(o, v, H.ERI, t1, t2).cpu().numpy() - g["rand_Wmnij"]).max() < tol
    assert np.abs(cc.build_Wmbej(o, v, H.ERI, H.L, t1, t2).cpu().numpy() - g["rand_Wmbej"]).max() < tol
    assert np.abs(cc.build_Wmbje(o, v, H.ERI, t1, t2).cpu().numpy() - g["rand_Wmbje"]).max() < tol
    assert np.abs(cc.build_Zmbij(o, v, H.ERI, t1, t2).cpu().numpy() - g["rand_Zmbij"]).max() < tol
    for (f1, f2) in ((1.0, 1.0), (1.0, 0.5), (0.5, 1.0)):
        assert np.abs(cc.build_tau(t1, t2, f1, f2).cpu().numpy() - g["rand_tau_%g_%g" % (f1, f2)]).max() < tol
    r1, r2 = cc.residuals(F, t1, t2)
    assert np.abs(r1.cpu().numpy() - g["rand_r1"]).max() < tol
    assert np.abs(r2.cpu().numpy() - g["rand_r2"]).max() < tol
    assert np.abs(cc.r_T1(o, v, F, H.ERI, H.L, t1, t2).cpu().numpy() - g["rand_r1"]).max() < tol
    assert np.abs(cc.r_T2(o, v, F, H.ERI, t1, t2).cpu().numpy() - g["rand_r2"]).max() < tol
    assert abs(float(cc.cc_energy(o, v, F, H.L, t1, t2)) - float(g["rand_ecc"])) < tol
    # the lazily built denominators
    eo, ev = syn.eps[:syn.no], syn.eps[syn.no:]
    assert np.abs(cc.Dia.cpu().numpy() - (eo[:, None] - ev)).max() < 1e-14
    D2 = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev
    assert np.abs(cc.Dijab.cpu().numpy() - D2).max() < 1e-14


def test_solve_cc_trace(golden, dev):
    g, syn = golden
    cc = make_wfn(syn, "CCSD(T)")
    ecc = cc.solve_cc(1e-12, 1e-12, 100)
    ref = g["trace_ecc_rms"]
    tr = np.array(cc.trace)
    assert len(tr) == len(ref)
    assert np.abs(tr[:, 0] - ref[:, 0]).max() < 1e-11
    assert np.all(np.abs(tr[:, 1] - ref[:, 1]) <= 1e-5 * np.abs(ref[:, 1]) + 1e-14)
    assert abs(float(ecc) - float(g["e_total_ccsd_t"])) < 1e-11
    assert np.abs(cc.t1.cpu().numpy() - g["conv_t1"]).max() < 1e-10
    assert np.abs(cc.t2.cpu().numpy() - g["conv_t2"]).max() < 1e-10
    assert float(cc.ecc) == float(ecc)


def test_diis_matches_reference(golden, dev):
    g, syn = golden
    d1, d2 = g["diis_in_t1"], g["diis_in_t2"]
    diis = pycc_b200.helper_diis(T(d1[0]), T(d2[0]), 4)
    x1, x2 = d1[0], d2[0]
    for n in range(1, 7):
        x1 = x1 + d1[n]
        x2 = x2 + d2[n]
        diis.add_error_vector(T(x1), T(x2))
        y1, y2 = diis.extrapolate(T(x1), T(x2))
        x1, x2 = y1.cpu().numpy().copy(), y2.cpu().numpy().copy()
        assert np.abs(x1 - g["diis_out_t1"][n - 1]).max() < 1e-10
        assert np.abs(x2 - g["diis_out_t2"][n - 1]).max() < 1e-10
    off = pycc_b200.helper_diis(T(d1[0]), T(d2[0]), 0)
    a, b = off.extrapolate(T(d1[1]), T(d2[1]))
    assert np.array_equal(a.cpu().numpy(), d1[1]) and np.array_equal(b.cpu().numpy(), d2[1])


def test_diis_history_longer_than_16(dev):
    """max_diis > 16, and start_diis > max_diis (the history then holds more vectors than max_diis: the reference trims
    one per extrapolate call, utils.py:317-320) -- both beyond the 16-vector limit of the multi_dot / multi_axpy kernels."""
    rng = np.random.default_rng(5)
    no, nv = 2, 3
    for max_diis, quiet_steps, steps in ((20, 0, 24), (8, 19, 24)):
        x1, x2 = rng.standard_normal((no, nv)), rng.standard_normal((no, no, nv, nv))
        ours = pycc_b200.helper_diis(T(x1), T(x2), max_diis)
        from oracle import ccsd_oracle as co
        ref = co.Diis(x1, x2, max_diis)
        for n in range(steps):
            x1 = x1 + 0.5 ** n * rng.standard_normal((no, nv))
            x2 = x2 + 0.5 ** n * rng.standard_normal((no, no, nv, nv))
            ours.add_error_vector(T(x1), T(x2))
            ref.add_error_vector(x1, x2)
            if n >= quiet_steps:
                y1, y2 = ours.extrapolate(T(x1), T(x2))
                r1, r2 = ref.extrapolate(x1, x2)
                assert np.abs(y1.cpu().numpy() - r1).max() < 1e-8 and np.abs(y2.cpu().numpy() - r2).max() < 1e-8
                x1, x2 = r1, r2
        assert ours.diis_size > 16


def test_triples(golden, dev):
    g, syn = golden
    cc = make_wfn(syn, "CCSD(T)")
    cc.t1, cc.t2 = T(g["conv_t1"]), T(g["conv_t2"])
    o, v, H = cc.o, cc.v, cc.H
    et = cctriples.t_tjl(cc)
    assert abs(float(et) - float(g["e_t_tjl"])) < 1e-12
    eng = cctriples.TriplesEngine(cc)
    for n, (i, j, k) in enumerate(g["triples"]):
        i, j, k = int(i), int(j), int(k)
        w3, d3 = eng.t3_parts(i, j, k, False)
        assert np.abs(w3.cpu().numpy() - g["W3"][n]).max() < 1e-12
        assert np.abs((w3 + d3).cpu().numpy() - g["V3"][n]).max() < 1e-12
        # reference-signature functions with the reference's own slicing of H.ERI
        t3c = cctriples.t3c_ijk(o, v, i, j, k, cc.t2, H.ERI[v, v, v, o], H.ERI[o, v, o, o], H.F, cc.contract, True)
        t3d = cctriples.t3d_ijk(o, v, i, j, k, cc.t1, cc.t2, H.ERI[o, o, v, v], H.F, cc.contract, True)
        assert np.abs(t3c.cpu().numpy() - g["t3c_denom"][n]).max() < 1e-12
        assert np.abs(t3d.cpu().numpy() - g["t3d_denom"][n]).max() < 1e-12
    # small batches exercise the chunking of the triple list
    eng2 = cctriples.TriplesEngine(cc, q_bytes=1)
    assert eng2.nb_max == 1
    tl = [t for t in cctriples.triples_list(cc.no) if not (t[0] == t[1] == t[2])]
    assert abs(float(eng2.energy(tl)[0]) - float(g["e_t_tjl"])) < 1e-12


@pytest.mark.parametrize("no,nv", [(4, 10), (6, 26)])
def test_paired_triples_equal_six_array_form(dev, no, nv):
    """t_tjl sums the six t3 products pairwise inside the GEMM (four K segments, three arrays per triple through HBM):
    same E(T) as the six-array form and as the numpy oracle, ragged 8-cubes included"""
    from pycc_b200.synthetic import make_synthetic, blocks_from_factor
    from oracle import triples_oracle as to
    syn = make_synthetic(no, nv, seed=7, fock_noise=0.01)
    cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True)
    cc.t1 = T(0.05 * np.random.default_rng(3).standard_normal((no, nv)))
    trip = [t for t in cctriples.triples_list(no) if not (t[0] == t[1] == t[2])]
    e6 = cctriples.TriplesEngine(cc, paired=False)
    e3 = cctriples.TriplesEngine(cc, paired=True)
    assert e3.paired and e3.nq == 3 and not e6.paired
    a, b = float(e6.energy(trip)[0]), float(e3.energy(trip)[0])
    bl = blocks_from_factor(syn, names=("ovvv", "ooov", "oovv"))
    want = to.t_tjl(cc.t1.cpu().numpy(), cc.t2.cpu().numpy(), syn.F, bl["ovvv"], bl["ooov"], bl["oovv"])
    assert abs(a - want) < 1e-12 and abs(b - want) < 1e-12
    w6, d6 = e6.t3_parts(no - 1, 1, 0, True)
    w3, d3 = e3.t3_parts(no - 1, 1, 0, True)
    assert float((w6 - w3).abs().max()) < 1e-13 and float((d6 - d3).abs().max()) == 0.0
    assert abs(float(cctriples.t_tjl(cc)) - want) < 1e-12


def test_vikings_cross_formulations(dev):
    # the reference's own cross-check (tests/test_005_ccsd_t_energy.py:30-36), on the smallest golden case
    from tests.conftest import load_golden, GOLDEN
    g, syn = load_golden([p for p in GOLDEN if "o3v7" in p][0])
    cc = make_wfn(syn, "CCSD(T)")
    cc.t1, cc.t2 = T(g["conv_t1"]), T(g["conv_t2"])
    assert abs(float(cctriples.t_vikings(cc)) - float(g["e_t_vikings"])) < 1e-12
    assert abs(float(cctriples.t_vikings_inverted(cc)) - float(g["e_t_vikings_inverted"])) < 1e-12


def test_ccd_is_ccsd_without_singles(dev):
    from tests.conftest import load_golden, GOLDEN
    from oracle import ccsd_oracle as co
    g, syn = load_golden(GOLDEN[0])
    cc = make_wfn(syn, "CCD")
    e = cc.solve_cc(1e-11, 1e-11)
    assert float(torch.abs(cc.t1).max()) == 0.0
    # oracle CCD: CCSD equations with t1 pinned to zero (reference ccwfn.py: the CCD branches drop every t1 term)
    P = co.Problem(co.blocks_from_full(full_eri(syn), syn.no), syn.F, syn.no)
    t1, t2 = P.guess()
    ecc = P.cc_energy(syn.F, t1, t2)
    diis = co.Diis(t1, t2, 8)
    for it in range(100):
        last = ecc
        r1, r2 = P.residuals(syn.F, t1, t2)
        t2 = t2 + r2 / P.Dijab
        rms = np.sqrt(np.sum((r2 / P.Dijab) ** 2))
        ecc = P.cc_energy(syn.F, t1, t2)
        if abs(ecc - last) < 1e-11 and rms < 1e-11:
            break
        diis.add_error_vector(t1, t2)
        t1, t2 = diis.extrapolate(t1, t2)
        t1 = np.zeros_like(t1)
    assert abs(float(e) - ecc) < 1e-10


@pytest.mark.parametrize("precision", ["MP", "SP"])
def test_mixed_precision_energy(golden, dev, precision):
    """precision='MP' (split-TF32 contractions on tcgen05, FP64 accumulation; 'SP' is served by the same path):
    converged E(CCSD) and E(CCSD(T)) within 1e-6 Eh of the reference's FP64 values (north_star tolerance), amplitudes
    within 1e-6; <ab|ef> is resident as TF32 planes only and constant operands are split once."""
    from pycc_b200 import kernels as K
    g, syn = golden
    keep = (K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles)
    K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = 0.0, 1, 1      # golden cases are tiny: force every K-major GEMM onto the mixed kernel
    try:
        cc = pycc_b200.ccwfn(IntegralReference.from_synthetic(syn), model="CCSD(T)", device='GPU',
                             precision=precision, quiet=True)
        assert cc.precision == precision and cc.mixed
        assert cc.H.vvvv_planes is not None and not cc.H.has("vvvv")
        before = dict(K.MIXED.stats)
        # the split contractions carry ~2^-22 relative jitter: rms stalls near 3e-9, so converge to the reference's
        # default thresholds (1e-7), not to 1e-10
        e = cc.solve_cc(1e-7, 1e-7, 100)
        assert e is not None
        assert K.MIXED.stats["gemm"] > before["gemm"] + 3 * len(cc.trace)
        assert K.MIXED.stats["split_cached"] > before["split_cached"]
        assert not K.MIXED.on                              # the switch is scoped to the residual evaluation
        assert abs(cc.trace[-1][0] - float(g["trace_ecc_rms"][-1, 0])) < 1e-6
        assert abs(float(e) - float(g["e_total_ccsd_t"])) < 1e-6
        assert np.abs(cc.t2.cpu().numpy() - g["conv_t2"]).max() < 1e-6
        assert np.abs(cc.t1.cpu().numpy() - g["conv_t1"]).max() < 1e-6
        # not silently the FP64 path
        assert np.abs(cc.t2.cpu().numpy() - g["conv_t2"]).max() > 1e-13
    finally:
        K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = keep


def test_mixed_precision_from_host_arrays_releases_fp64_ladder_block(dev):
    """from_arrays + 'MP': <ab|ef> is converted to planes, the FP64 block is released AND forgotten as a constant
    (its storage can be reused by per-iteration tensors, whose TF32 planes must never be cached)."""
    from pycc_b200 import kernels as K
    from pycc_b200.synthetic import make_synthetic
    syn = make_synthetic(4, 8, seed=3)
    ref = IntegralReference.from_arrays(syn.F, full_eri(syn), syn.no)
    keep = (K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles)
    K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = 0.0, 1, 1
    try:
        n0 = len(K._CONST)
        mp = pycc_b200.ccwfn(ref, model="CCSD", precision="MP", quiet=True)
        assert not mp.H.has("vvvv") and mp.H.vvvv_planes is not None
        live = {t.untyped_storage().data_ptr() for t in list(mp.H._blocks.values()) + list(mp.H._derived.values())}
        assert set(K._CONST) - live == set() or len(K._CONST) - n0 <= len(live)
        dp = pycc_b200.ccwfn(IntegralReference.from_arrays(syn.F, full_eri(syn), syn.no), model="CCSD", quiet=True)
        e_mp, e_dp = mp.solve_cc(1e-8, 1e-7), dp.solve_cc(1e-10, 1e-10)
        assert abs(float(e_mp) - float(e_dp)) < 1e-6
        for (a, _), (b, _) in zip(mp.trace, dp.trace):          # every iteration, not only the converged value
            assert abs(a - b) < 1e-6
    finally:
        K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = keep


def test_mixed_precision_keeps_callers_hamiltonian(dev):
    """A BlockHamiltonian handed in by the caller keeps its FP64 <ab|ef> (it may serve a 'DP' wavefunction too)."""
    from pycc_b200.hamiltonian import BlockHamiltonian
    from pycc_b200.synthetic import make_synthetic
    syn = make_synthetic(3, 6, seed=5)
    H = BlockHamiltonian.from_factor(syn, DEV[0])
    mp = pycc_b200.ccwfn(H, model="CCSD", precision="MP", quiet=True)
    dp = pycc_b200.ccwfn(H, model="CCSD", precision="DP", quiet=True)
    assert H.has("vvvv") and H.vvvv_planes is not None
    e_mp, e_dp = mp.solve_cc(1e-8, 1e-7), dp.solve_cc(1e-10, 1e-10)
    assert 1e-14 < abs(float(e_mp) - float(e_dp)) < 1e-6


def test_keyword_errors():
    from pycc_b200.exceptions import InvalidKeywordError, PyCCError
    from pycc_b200.synthetic import make_synthetic
    syn = make_synthetic(2, 3)
    with pytest.raises(InvalidKeywordError):
        pycc_b200.ccwfn(syn, model="CCSDT")
    with pytest.raises(ValueError):
        pycc_b200.ccwfn(syn, model="nope")
    with pytest.raises(PyCCError):
        pycc_b200.ccwfn(syn, model="CCSD", bogus=1)
    with pytest.raises(TypeError):
        pycc_b200.ccwfn(syn, frozen_core=True)
    with pytest.raises(InvalidKeywordError):
        pycc_b200.ccwfn(syn, precision="HP")
    with pytest.raises(PyCCError):
        pycc_b200.ccwfn(syn, device="CPU")
    with pytest.raises(TypeError):
        pycc_b200.ccwfn(object())


def test_no_cpu_fallback():
    """Without the test double the package must refuse CPU tensors loudly."""
    from pycc_b200 import kernels as K
    from pycc_b200._lib import B200ccError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B200ccError):
        K.axpbyz(1.0, torch.zeros(4, dtype=torch.float64), 0.0, None, torch.zeros(4, dtype=torch.float64))


@pytest.mark.gpu
@pytest.mark.parametrize("no,nv,seed,noise", [(8, 40, 0, 0.0), (7, 33, 1, 0.01), (10, 64, 2, 0.0)])
def test_medium_size_vs_oracle(no, nv, seed, noise):
    """CUDA path vs the numpy oracle on the same seeded synthetic inputs, at sizes the oracle does in
    seconds: E(CCSD), E(T) within 1e-10 Eh, amplitudes within 1e-9 (north_star tolerances)."""
    from oracle import ccsd_oracle as co, triples_oracle as to
    from pycc_b200.synthetic import make_synthetic, blocks_from_factor
    assert torch.cuda.is_available()
    DEV[0] = torch.device("cuda:0")
    try:
        syn = make_synthetic(no, nv, seed=seed, fock_noise=noise)
        b = blocks_from_factor(syn)
        P = co.Problem(b, syn.F, no)
        e_ref, t1_ref, t2_ref, trace = co.solve_cc(P, 1e-11, 1e-11, 100)
        et_ref = to.t_tjl(t1_ref, t2_ref, syn.F, b["ovvv"], b["ooov"], b["oovv"])
        cc = make_wfn(syn, "CCSD(T)")
        e = cc.solve_cc(1e-11, 1e-11, 100)
        assert len(cc.trace) == len(trace)
        e_ccsd = cc.trace[-1][0]
        assert abs(e_ccsd - e_ref) < 1e-10
        assert abs(float(e) - (e_ref + et_ref)) < 1e-10
        assert np.abs(cc.t1.cpu().numpy() - t1_ref).max() < 1e-9
        assert np.abs(cc.t2.cpu().numpy() - t2_ref).max() < 1e-9
        # first-iteration residuals (pre-DIIS), amplitude-independent check of every term
        r1_ref, r2_ref = P.residuals(syn.F, *P.guess())
        cc2 = make_wfn(syn, "CCSD")
        r1, r2 = cc2.residuals(cc2.H.F, cc2.t1, cc2.t2)
        assert np.abs(r1.cpu().numpy() - r1_ref).max() < 1e-11
        assert np.abs(r2.cpu().numpy() - r2_ref).max() < 1e-11
    finally:
        DEV[0] = torch.device("cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("no,nv,seed,force", [(8, 40, 0, True), (10, 64, 2, True), (12, 72, 3, False)])
def test_medium_size_mixed_vs_oracle(no, nv, seed, force):
    """precision='MP' on the GPU (tcgen05 split-TF32 GEMMs) vs the FP64 numpy oracle on the same inputs:
    E(CCSD) and E(CCSD(T)) within 1e-6 Eh (north_star tolerance for mixed precision), amplitudes within 1e-6.
    ``force``: every K-major GEMM goes through the mixed kernel regardless of size; otherwise the product
    thresholds apply (ladder always mixed)."""
    from oracle import ccsd_oracle as co, triples_oracle as to
    from pycc_b200 import kernels as K
    from pycc_b200.synthetic import make_synthetic, blocks_from_factor
    assert torch.cuda.is_available()
    DEV[0] = torch.device("cuda:0")
    keep = (K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles)
    if force:
        K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = 0.0, 1, 1
    try:
        syn = make_synthetic(no, nv, seed=seed)
        b = blocks_from_factor(syn)
        P = co.Problem(b, syn.F, no)
        e_ref, t1_ref, t2_ref, trace = co.solve_cc(P, 1e-10, 1e-10, 100)
        et_ref = to.t_tjl(t1_ref, t2_ref, syn.F, b["ovvv"], b["ooov"], b["oovv"])
        before = K.MIXED.stats["gemm"]
        cc = pycc_b200.ccwfn(IntegralReference.from_synthetic(syn), model="CCSD(T)", device='GPU', precision="MP",
                             quiet=True)
        e = cc.solve_cc(1e-7, 1e-7, 100)
        assert e is not None and K.MIXED.stats["gemm"] > before
        assert abs(cc.trace[-1][0] - e_ref) < 1e-6
        assert abs(float(e) - (e_ref + et_ref)) < 1e-6
        assert np.abs(cc.t1.cpu().numpy() - t1_ref).max() < 1e-6
        assert np.abs(cc.t2.cpu().numpy() - t2_ref).max() < 1e-6
    finally:
        K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = keep
        DEV[0] = torch.device("cpu")


def test_frozen_core_offsets(dev):
    """nfzc > 0: the reference keeps the full MO space and offsets the o/v slices (wavefunction.py:293-332);
    here the same F (full n x n) and a full-space ERI are given with nfzc = 2."""
    from oracle import ccsd_oracle as co, triples_oracle as to
    from pycc_b200.synthetic import make_synthetic
    nfzc, no, nv = 2, 3, 6
    syn = make_synthetic(nfzc + no, nv, seed=5, fock_noise=0.01)      # n = 11 orbitals, the first 2 frozen
    ERI = full_eri(syn)
    ref = IntegralReference.from_arrays(syn.F, ERI, no, nfzc, eref=-1.25)
    cc = pycc_b200.ccwfn(ref, model="CCSD(T)", device="GPU", quiet=True)
    assert (cc.nfzc, cc.no, cc.nv, cc.nmo) == (nfzc, no, nv, nfzc + no + nv)
    assert cc.o == slice(nfzc, nfzc + no) and cc.v == slice(nfzc + no, cc.nmo)
    e = cc.solve_cc(1e-11, 1e-11)
    P = co.Problem(co.blocks_from_full(ERI, no, nfzc), syn.F, no, nfzc)
    e_ref, t1, t2, trace = co.solve_cc(P, 1e-11, 1e-11)
    et = to.t_tjl(t1, t2, syn.F, P.ovvv, P.ooov, P.oovv, nfzc=nfzc)
    assert abs(float(e) - (e_ref + et)) < 1e-10
    assert np.abs(cc.t2.cpu().numpy() - t2).max() < 1e-9
    assert cc.eref == -1.25
    # the block views follow the offset slices
    assert np.abs(cc.H.ERI[cc.o, cc.v, cc.v, cc.o].cpu().numpy() - ERI[cc.o, cc.v, cc.v, cc.o]).max() < 1e-13


def test_not_converged_returns_none(dev):
    from pycc_b200.synthetic import make_synthetic
    cc = make_wfn(make_synthetic(3, 5, seed=1), "CCSD")
    assert cc.solve_cc(1e-12, 1e-12, maxiter=2) is None            # reference: falls off the loop (ccwfn.py:268-319)
    cc2 = make_wfn(make_synthetic(3, 5, seed=1), "CCSD")
    e_nodiis = cc2.solve_cc(1e-10, 1e-10, maxiter=200, max_diis=0)  # max_diis = 0 switches DIIS off (utils.py:311)
    cc3 = make_wfn(make_synthetic(3, 5, seed=1), "CCSD")
    e_diis = cc3.solve_cc(1e-10, 1e-10)
    assert abs(float(e_nodiis) - float(e_diis)) < 1e-9
    assert len(cc2.trace) > len(cc3.trace)


def test_tiny_and_ragged_shapes(dev):
    """o = 1 (no i>j>k triple survives), v = 1, and odd extents."""
    from oracle import ccsd_oracle as co, triples_oracle as to
    from pycc_b200.synthetic import make_synthetic, blocks_from_factor
    for (no, nv) in ((1, 3), (2, 1), (5, 9)):
        syn = make_synthetic(no, nv, seed=no + nv)
        b = blocks_from_factor(syn)
        P = co.Problem(b, syn.F, no)
        e_ref, t1, t2, _ = co.solve_cc(P, 1e-11, 1e-11)
        et = to.t_tjl(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
        cc = make_wfn(syn, "CCSD(T)")
        e = cc.solve_cc(1e-11, 1e-11)
        assert abs(float(e) - (e_ref + et)) < 1e-10, (no, nv)


def test_abc_batched_t3_and_helpers(golden, dev):
    """t3c_abc / t3d_abc (cctriples.py:75-105, 149-173) and the utils helpers keep the reference's results."""
    from pycc_b200 import utils
    g, syn = golden
    cc = make_wfn(syn, "CCSD(T)")
    cc.t1, cc.t2 = T(g["conv_t1"]), T(g["conv_t2"])
    o, v, H = cc.o, cc.v, cc.H
    a, b, c = 2, 1, 0
    t3c = cctriples.t3c_abc(o, v, a, b, c, cc.t2, H.ERI[v, v, v, o], H.ERI[o, v, o, o], H.F, cc.contract, True)
    t3d = cctriples.t3d_abc(o, v, a, b, c, cc.t1, cc.t2, H.ERI[o, o, v, v], H.F, cc.contract, True)
    assert np.abs(t3c.cpu().numpy() - g["t3c_abc_210"]).max() < 1e-12
    assert np.abs(t3d.cpu().numpy() - g["t3d_abc_210"]).max() < 1e-12
    x = utils.clone(cc.t2)
    assert x.data_ptr() != cc.t2.data_ptr() and torch.equal(x, cc.t2)
    d = float(utils.dot(cc.t2.reshape(-1), cc.t2.reshape(-1)))
    assert abs(d - float(np.sum(g["conv_t2"] ** 2))) < 1e-12
    assert float(utils.zeros_like(cc.t1).abs().max()) == 0.0


def test_cube_blocked_q_layout(golden, dev):
    """The optional 8x8x8-cube-blocked Q layout (GEMM epilogue out_cube_nv + blocked energy / assemble kernels)
    gives the same W3 and E(T) as the plain layout."""
    g, syn = golden
    if syn.no % 2 or syn.nv % 2:
        pytest.skip("TMA path needs even o, v")
    cc = make_wfn(syn, "CCSD(T)")
    cc.t1, cc.t2 = T(g["conv_t1"]), T(g["conv_t2"])
    eng = cctriples.TriplesEngine(cc, cube_q=True)
    assert eng.cube
    tl = [t for t in cctriples.triples_list(cc.no) if not (t[0] == t[1] == t[2])]
    assert abs(float(eng.energy(tl)[0]) - float(g["e_t_tjl"])) < 1e-12
    i, j, k = (int(x) for x in g["triples"][0])
    w3, _ = eng.t3_parts(i, j, k, False)
    assert np.abs(w3.cpu().numpy() - g["W3"][0]).max() < 1e-12


@pytest.mark.parametrize("no,nv", [(1, 1), (1, 2), (2, 1), (5, 1), (1, 5), (3, 2)])
def test_degenerate_shapes(dev, no, nv):
    """one occupied / one virtual orbital (H2 in a minimal basis has o = v = 1): every tile is ragged, the pair-packed
    ladder has a single pair, the (T) loop a single triple -- CCSD(T) and both (T) formulations against the oracle"""
    from pycc_b200.synthetic import make_synthetic, blocks_from_factor
    from oracle import ccsd_oracle as co, triples_oracle as to
    syn = make_synthetic(no, nv, seed=1, fock_noise=0.01)
    b = blocks_from_factor(syn)
    P = co.Problem(b, syn.F, no)
    e_ref, t1, t2, trace = co.solve_cc(P, 1e-11, 1e-11)
    et = to.t_tjl(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
    cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True)
    e = cc.solve_cc(1e-11, 1e-11)
    # (the iteration COUNT is not compared: with one or two amplitudes the DIIS matrix of 8 stored vectors is singular,
    #  and which of the equivalent extrapolations comes out depends on the last bits of the dots)
    assert e is not None and abs(float(e) - (e_ref + et)) < 1e-10
    assert np.abs(cc.t2.cpu().numpy() - t2).max() < 1e-9
    assert abs(float(cctriples.t_vikings(cc)) - float(cctriples.t_tjl(cc))) < 1e-12
    assert abs(float(cctriples.t_tjl(cc)) - et) < 1e-10
