"""The particle-particle ladder in symmetric / antisymmetric pair form (csrc/pairs.cu, ccwfn._ladder; reference
ccwfn.py:931 as written, the split itself as in ccwfn.py:1054-1120): the four packing kernels against numpy, the ladder
against the plain einsum on unsymmetric and pair-symmetric tau, and every consumer of <ab|ef> with the full FP64 block
NOT resident (HBAR / Lambda, CC2, CC3, H.ERI[v,v,v,v]) against the reference's golden vectors.  `emu` / `cuda` as in
test_ccsd.py."""
import glob
import os

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200 import kernels as K
from pycc_b200.hamiltonian import BlockHamiltonian
from pycc_b200.synthetic import Synthetic, blocks_from_factor, make_synthetic
from tests import emu
from tests.conftest import load_golden, GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = [torch.device("cpu")]


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        DEV[0] = torch.device("cpu")
        with emu.install():
            yield DEV[0]
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        DEV[0] = torch.device("cuda:0")
        yield DEV[0]
        DEV[0] = torch.device("cpu")


@pytest.fixture
def packed_only():
    """Release the full FP64 <ab|ef> block as soon as it is packed, whatever its size."""
    keep = BlockHamiltonian.keep_vvvv_bytes
    BlockHamiltonian.keep_vvvv_bytes = 0
    yield
    BlockHamiltonian.keep_vvvv_bytes = keep


def T(x):
    return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True)).to(DEV[0])


def sym_vvvv(nv, rng):
    """a random <ab|ef> with the symmetries the pair form relies on: <ab|ef> = <ba|fe> (and = <ef|ab>)"""
    X = rng.standard_normal((nv, nv, nv, nv))
    X = X + X.transpose(1, 0, 3, 2)
    return X + X.transpose(2, 3, 0, 1)


@pytest.mark.parametrize("nv,a0,a1", [(1, 0, 1), (5, 0, 5), (7, 2, 6), (33, 0, 33), (40, 31, 40)])
def test_pack_and_unpack_pairs(dev, nv, a0, a1):
    rng = np.random.default_rng(nv)
    X = sym_vvvv(nv, rng)
    nq, ldq = K.pair_count(nv), K.pair_ld(nv)
    npl = K.pair_count(a1) - K.pair_count(a0)
    V = torch.full((2, npl, ldq), 7.0, dtype=torch.float64, device=dev)
    # source handed over in [a, e, b, f] memory order through a strided view, as the generating GEMM leaves it
    src = T(X[a0:a1].transpose(0, 2, 1, 3)).permute(0, 2, 1, 3)
    K.pack_pairs(src, nv, a0, a1, V[0], V[1], ldq)
    Vn = V.cpu().numpy()
    e, f = np.tril_indices(nv)
    row = 0
    for a in range(a0, a1):
        for b in range(a + 1):
            plus = np.where(e == f, X[a, b][e, f], X[a, b][e, f] + X[a, b][f, e])
            minus = np.zeros(nq) if a == b else np.where(e == f, 0.0, X[a, b][e, f] - X[a, b][f, e])
            assert np.abs(Vn[0, row, :nq] - plus).max() < 1e-14 and np.abs(Vn[1, row, :nq] - minus).max() < 1e-14
            row += 1
    assert row == npl and np.all(Vn[:, :, nq:] == 0.0)
    back = K.unpack_pairs(V[0], V[1], ldq, nv, npl).cpu().numpy()
    row = 0
    for a in range(a0, a1):
        for b in range(a + 1):
            assert np.abs(back[row] - X[a, b]).max() < 1e-14
            row += 1


@pytest.mark.parametrize("no,nv", [(1, 1), (3, 5), (4, 33), (5, 40)])
@pytest.mark.parametrize("tri", [False, True])
def test_pack_tau_and_ladder_unpack(dev, no, nv, tri):
    rng = np.random.default_rng(no * nv)
    tau = rng.standard_normal((no, no, nv, nv))
    if tri:
        tau = tau + tau.transpose(1, 0, 3, 2)
    P = K.pack_tau(T(tau), tri).cpu().numpy()
    nq = K.pair_count(nv)
    rows = list(zip(*np.tril_indices(no))) if tri else [(i, j) for i in range(no) for j in range(no)]
    e, f = np.tril_indices(nv)
    assert P.shape[:2] == (2, len(rows)) and np.all(P[:, :, nq:] == 0.0)
    for m, (i, j) in enumerate(rows):
        x, y = tau[i, j][e, f], tau[i, j][f, e]
        assert np.abs(P[0, m, :nq] - np.where(e == f, x, 0.5 * (x + y))).max() < 1e-15
        assert np.abs(P[1, m, :nq] - np.where(e == f, 0.0, 0.5 * (x - y))).max() < 1e-15
    # scatter of S / A for a row range that does not start at 0 and is not 32-aligned
    a0, a1 = (nv // 3, nv) if nv > 2 else (0, nv)
    npl = K.pair_count(a1) - K.pair_count(a0)
    lds = npl + 3
    S = rng.standard_normal((len(rows), lds))
    A = rng.standard_normal((len(rows), lds))
    r0 = rng.standard_normal((no, no, nv, nv))
    r2 = T(r0)
    K.ladder_unpack(T(S), T(A), lds, no, nv, tri, a0, a1, 0.5, r2)
    want = r0.copy()
    for m, (i, j) in enumerate(rows):
        col = 0
        for a in range(a0, a1):
            for b in range(a + 1):
                s, d = S[m, col], A[m, col]
                want[i, j, a, b] += 0.5 * (s + d)
                if a != b:
                    want[i, j, b, a] += 0.5 * (s - d)
                if tri and i != j:
                    want[j, i, a, b] += 0.5 * (s - d)
                    if a != b:
                        want[j, i, b, a] += 0.5 * (s + d)
                col += 1
    assert np.abs(r2.cpu().numpy() - want).max() < 1e-14


@pytest.mark.parametrize("no,nv", [(1, 1), (3, 7), (2, 33), (5, 40)])
def test_ring_layouts_and_pair_rows(dev, no, nv):
    rng = np.random.default_rng(no + nv)
    t2 = rng.standard_normal((no, no, nv, nv))
    u, tb = K.ring_layouts(T(t2))
    assert np.array_equal(u.cpu().numpy(), (2.0 * t2 - t2.transpose(0, 1, 3, 2)).transpose(0, 2, 1, 3))
    assert np.array_equal(tb.cpu().numpy(), t2.transpose(0, 3, 1, 2))
    # X+- of integral rows and the (i,j) / (j,i) scatter of S +- A
    X = rng.standard_normal((no * nv, nv, nv))
    P = K.pack_rows(T(X), no * nv, nv).cpu().numpy()
    e, f = np.tril_indices(nv)
    nq = len(e)
    assert np.array_equal(P[0, :, :nq], np.where(e == f, X[:, e, f], X[:, e, f] + X[:, f, e]))
    assert np.array_equal(P[1, :, :nq], np.where(e == f, 0.0, X[:, e, f] - X[:, f, e])) and np.all(P[:, :, nq:] == 0.0)
    npair, ncols = no * (no + 1) // 2, 2 * nv + 1
    S, A = rng.standard_normal((npair, ncols + 2)), rng.standard_normal((npair, ncols + 2))
    out = torch.zeros((no, no, ncols), dtype=torch.float64, device=dev)
    K.pair_rows_unpack(T(S), T(A), ncols + 2, no, ncols, out, ncols)
    i, j = np.tril_indices(no)
    want = np.zeros((no, no, ncols))
    want[j, i] = (S - A)[:, :ncols]
    want[i, j] = (S + A)[:, :ncols]
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("no,nv", [(3, 7), (4, 10), (2, 33)])
def test_ladder_matches_einsum(dev, no, nv, packed_only):
    """general mode on an unsymmetric tau, tri mode on a pair-symmetric one -- both against 'ijef,abef->ijab'"""
    syn = make_synthetic(no, nv, seed=3)
    vvvv = blocks_from_factor(syn, names=("vvvv",))["vvvv"]
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    assert not cc.H.has("vvvv") and cc.H.vvvv_packed is not None
    rng = np.random.default_rng(1)
    tau = rng.standard_normal((no, no, nv, nv))
    r2 = T(np.zeros_like(tau))
    cc._ladder(T(tau), r2)
    assert np.abs(r2.cpu().numpy() - 0.5 * np.einsum("ijef,abef->ijab", tau, vvvv)).max() < 1e-12
    assert cc.ladder_flops == 2.0 * 2 * no * no * K.pair_count(nv) ** 2
    tau = tau + tau.transpose(1, 0, 3, 2)
    r2 = T(np.zeros_like(tau))
    cc._ladder(T(tau), r2, symmetric=True)
    assert np.abs(r2.cpu().numpy() - 0.5 * np.einsum("ijef,abef->ijab", tau, vvvv)).max() < 1e-12
    assert cc.ladder_flops == 2.0 * 2 * K.pair_count(no) * K.pair_count(nv) ** 2
    # H.ERI[v,v,v,v] on demand: the block is rebuilt from the packed form (BlockHamiltonian.materialize_vvvv)
    BlockHamiltonian.keep_vvvv_bytes = 1 << 30
    full = cc.H.ERI[cc.v, cc.v, cc.v, cc.v]
    assert np.abs(full.cpu().numpy() - vvvv).max() < 1e-14


@pytest.mark.parametrize("no,nv", [(3, 7), (4, 10), (5, 34)])
def test_z_in_pair_form_matches_dense(dev, no, nv):
    """Z_mbij = <mb|ef> tau_ijef (ccwfn.py:715) for pair-symmetric amplitudes: pair form (T+- of the ladder, packed
    <mb|ef>) against the dense o^3v^3 product and against numpy"""
    syn = make_synthetic(no, nv, seed=5, fock_noise=0.01)
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    rng = np.random.default_rng(2)
    t1 = 0.1 * rng.standard_normal((no, nv))
    t2 = rng.standard_normal((no, no, nv, nv))
    t2 = t2 + t2.transpose(1, 0, 3, 2)
    F = cc._check_F(cc.H.F)
    Zp = cc._intermediates(F, T(t1), T(t2), full=True, symmetric=True)["Zijmb"].cpu().numpy()
    Zd = cc._intermediates(F, T(t1), T(t2), full=True, symmetric=False)["Zijmb"].cpu().numpy()
    ovvv = blocks_from_factor(syn, names=("ovvv",))["ovvv"]
    tau = t2 + np.einsum("ia,jb->ijab", t1, t1)
    want = np.einsum("ijef,mbef->ijmb", tau, ovvv)
    assert np.abs(Zd - want).max() < 1e-12 and np.abs(Zp - want).max() < 1e-12
    assert ("ovvv_packed", 0, no) in cc.H._derived


@pytest.mark.parametrize("model", ["CCSD", "CCD"])
@pytest.mark.parametrize("no,nv", [(3, 7), (4, 10)])
def test_symmetric_path_equals_general_path(dev, model, no, nv):
    """the solve_cc path (pair-symmetric t2: ladder and Z on rows (i >= j), ring terms through D = W1 + W2/2 and W2 -- four
    o^3v^3 GEMMs instead of six) against the general path and the numpy oracle, on random pair-symmetric amplitudes far
    from convergence"""
    from oracle import ccsd_oracle as co
    syn = make_synthetic(no, nv, seed=9, fock_noise=0.02)
    cc = pycc_b200.ccwfn(syn, model=model, device="GPU", quiet=True)
    rng = np.random.default_rng(4)
    t1 = np.zeros((no, nv)) if model == "CCD" else 0.3 * rng.standard_normal((no, nv))
    t2 = rng.standard_normal((no, no, nv, nv))
    t2 = 0.3 * (t2 + t2.transpose(1, 0, 3, 2))
    F = cc.H.F
    r1s, hs = cc._residuals_half(F, T(t1), T(t2), symmetric=True)
    r1g, hg = cc._residuals_half(F, T(t1), T(t2), symmetric=False)
    K.symmetrize_r2(hs)
    K.symmetrize_r2(hg)
    P = co.Problem(blocks_from_factor(syn), syn.F, no)
    if model == "CCD":
        r2 = P.residuals_ccd(syn.F, t2) if hasattr(P, "residuals_ccd") else None
    else:
        r1, r2 = P.residuals(syn.F, t1, t2)
        assert np.abs(r1s.cpu().numpy() - r1).max() < 1e-11
    assert float((hs - hg).abs().max()) < 1e-11
    if r2 is not None:
        assert np.abs(hs.cpu().numpy() - r2).max() < 1e-11


def test_host_block_is_packed_and_large_blocks_released(dev, packed_only):
    """from_arrays / from_blocks: the FP64 block handed in is packed at construction and, beyond keep_vvvv_bytes,
    released; a caller's own BlockHamiltonian is left alone"""
    g, syn = load_golden(GOLDEN[0])
    b = blocks_from_factor(syn)
    from pycc_b200.wavefunction import IntegralReference
    cc = pycc_b200.ccwfn(IntegralReference.from_blocks(syn.F, b, syn.no), model="CCSD", device="GPU", quiet=True)
    assert not cc.H.has("vvvv") and cc.H.vvvv_packed is not None
    H = BlockHamiltonian(syn.F, {k: T(x) for k, x in b.items()}, syn.no, 0, DEV[0])
    cc2 = pycc_b200.ccwfn(H, model="CCSD", device="GPU", quiet=True)
    r1a, r2a = cc.residuals(cc.H.F, T(g["rand_t1"]), T(g["rand_t2"]))
    r1b, r2b = cc2.residuals(H.F, T(g["rand_t1"]), T(g["rand_t2"]))
    assert H.has("vvvv")
    for got in (r2a, r2b):
        assert np.abs(got.cpu().numpy() - g["rand_r2"]).max() < 1e-12


def _chain(path):
    tag = os.path.basename(path)[4:-4]
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag.split("_cc")[0])))
    return dict(np.load(path)), Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))


LAM = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "lam_*_ccsd.npz")))


@pytest.mark.parametrize("path", LAM, ids=[os.path.basename(p)[4:-4] for p in LAM])
def test_hbar_and_lambda_without_the_full_block(dev, packed_only, path):
    """CCSD -> HBAR -> Lambda with only the pair-packed <ab|ef> resident: Hvvvo's t_if<ab|ef> comes from pair chunks,
    the Lambda ladder from the packed GEMM; against the reference's goldens"""
    g, syn = _chain(path)
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    assert not cc.H.has("vvvv")
    cc.solve_cc(1e-12, 1e-12)
    hb = pycc_b200.cchbar(cc)
    assert not cc.H.has("vvvv")
    for k in ("Hvvvo", "Hovoo", "Hvv"):
        assert np.abs(getattr(hb, k).cpu().numpy() - g[k]).max() < 1e-11, k
    lam = pycc_b200.cclambda(cc, hb)
    e = lam.solve_lambda(1e-12, 1e-12)
    ref = g["trace_lecc_rms"]
    tr = np.array(lam.trace)
    assert len(tr) == len(ref) and np.abs(tr[:, 0] - ref[:, 0]).max() < 1e-11
    assert abs(float(e) - float(g["lecc"])) < 1e-11
    assert np.abs(lam.l2.cpu().numpy() - g["conv_l2"]).max() < 1e-10


@pytest.mark.parametrize("model", ["CC2", "CC3"])
def test_cc2_cc3_without_the_full_block(dev, packed_only, model):
    tag = "o3v7_s2"
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "%s_%s.npz" % (model.lower(), tag))))
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag)))
    syn = Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))
    cc = pycc_b200.ccwfn(syn, model=model, device="GPU", quiet=True)
    assert not cc.H.has("vvvv")
    e = cc.solve_cc(1e-12, 1e-12)
    assert not cc.H.has("vvvv")
    assert abs(float(e) - float(g["ecc"])) < 1e-11
