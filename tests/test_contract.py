"""Host-side contraction planner (pycc_b200/contract.py) against numpy.einsum, through the numpy
double of the C ABI (tests/emu.py).  Checks the pointer/stride/orientation logic, not the CUDA kernels."""
import numpy as np
import pytest
import torch

from pycc_b200.contract import Contractor
from tests import emu

DEV = [torch.device("cpu")]


def _T(x):
    return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True)).to(DEV[0])


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        DEV[0] = torch.device("cpu")
        with emu.install():
            yield DEV[0]
    else:
        assert torch.cuda.is_available()
        DEV[0] = torch.device("cuda:0")
        yield DEV[0]
        DEV[0] = torch.device("cpu")

O, V = 3, 5
DIMS = dict(i=O, j=O, k=O, l=O, m=O, n=O, a=V, b=V, c=V, d=V, e=V, f=V, P=7)

CASES = [
    "ijef,abef->ijab", "mnab,mnij->ijab", "ijef,mnef->mnij", "mbef,ijef->mbij", "mbef,jf->mbej",
    "jnfb,mnef->mbej", "njfb,mnef->mbej", "imae,mbej->ijab", "mjae,mbie->ijab", "ijae,be->ijab",
    "imab,mj->ijab", "ie,jabe->ijab", "ma,mbij->ijab", "mnaf,mnef->ae", "inef,mnef->mi", "nf,mnef->me",
    "ne,mnie->mi", "je,mnie->mnij", "ie,nmje->mnij", "ie,ae->ia", "mi,ma->ia", "imae,me->ia",
    "mnae,mnie->ia", "ia,ia->", "ijab,ijab->", "ia,jb->ijab", "Ppr,Pqs->pqrs", "me,ma->ae",
    "iajb,jb->ia", "bij,bjk->bik", "bij,bkj->bki", "ab,c->abc", "abe,ce->abc",
]


def rand(idx, rng):
    DIMS.update(p=4, q=5, r=3, s=6)
    return rng.standard_normal([DIMS[ch] for ch in idx])


@pytest.mark.parametrize("sub", CASES)
def test_planner_matches_einsum(sub, dev):
    rng = np.random.default_rng(len(sub))
    lhs, out = sub.split("->")
    ia, ib = lhs.split(",")
    A, B = rand(ia, rng), rand(ib, rng)
    ref = np.einsum(sub, A, B)
    if True:
        ct = Contractor()
        tA, tB = _T(A.copy()), _T(B.copy())
        got = ct(sub, tA, tB)
        assert np.abs(got.cpu().numpy() - ref).max() < 1e-12
        # accumulate form with alpha/beta into an existing (contiguous) output
        C0 = rng.standard_normal(ref.shape)
        tC = _T(C0.copy())
        ct(sub, tA, tB, out=tC, alpha=-0.5, beta=2.0)
        assert np.abs(tC.cpu().numpy() - (-0.5 * ref + 2.0 * C0)).max() < 1e-12
        # second call hits the plan cache
        got2 = ct(sub, tA, tB)
        assert np.abs(got2.cpu().numpy() - ref).max() < 1e-12


def test_strided_views_and_permuted_outputs(dev):
    rng = np.random.default_rng(7)
    if True:
        ct = Contractor()
        big = _T(rng.standard_normal((8, 8, 8, 8)))
        o, v = slice(0, 3), slice(3, 8)
        t2 = _T(rng.standard_normal((3, 3, 5, 5)))
        # sliced (non-contiguous) integral views, as pycc's ERI[o,o,v,v] produces
        got = ct("ijef,mnef->mnij", t2, big[o, o, v, v])
        ref = np.einsum("ijef,mnef->mnij", t2.cpu().numpy(), big[o, o, v, v].cpu().numpy())
        assert np.abs(got.cpu().numpy() - ref).max() < 1e-12
        # permuted operand views and a permuted output view
        out = torch.zeros((5, 3, 3, 5), dtype=torch.float64, device=DEV[0])
        ct("imae,mbej->ijab", t2.permute(1, 0, 3, 2), big[o, v, v, o], out=out.permute(1, 2, 0, 3))
        ref = np.einsum("imae,mbej->ijab", t2.permute(1, 0, 3, 2).cpu().numpy(), big[o, v, v, o].cpu().numpy())
        assert np.abs(out.permute(1, 2, 0, 3).cpu().numpy() - ref).max() < 1e-12


def test_three_operands_and_unary(dev):
    rng = np.random.default_rng(3)
    if True:
        ct = Contractor()
        t2 = _T(rng.standard_normal((3, 3, 5, 5)))
        t1 = _T(rng.standard_normal((3, 5)))
        f = _T(rng.standard_normal((3, 5)))
        got = ct("ijae,mb,me->ijab", t2, t1, f)
        ref = np.einsum("ijae,mb,me->ijab", t2.cpu().numpy(), t1.cpu().numpy(), f.cpu().numpy())
        assert np.abs(got.cpu().numpy() - ref).max() < 1e-12
        got = ct("ijab->jiba", t2)
        assert np.abs(got.cpu().numpy() - t2.cpu().numpy().transpose(1, 0, 3, 2)).max() == 0.0


def test_rejects_what_it_cannot_do(dev):
    from pycc_b200._lib import B200ccError
    if True:
        ct = Contractor()
        a = torch.zeros((3, 3), dtype=torch.float64, device=DEV[0])
        with pytest.raises(B200ccError):
            ct("ii,ij->j", a, a)
        with pytest.raises(B200ccError):
            ct("ij,jk->ik", a.float(), a.float())
