"""pycc_b200/planes.py: complex quantities as pairs of real planes contracted on the real kernels -- every branch of
``cterm`` / ``cprod`` (real x real, complex x real either way, complex x complex in the four- and the three-product
form, missing imaginary planes) and ``complex_tau`` against numpy's complex einsum.  Host logic through the numpy double
of the C ABI (the GPU lane exercises the same functions through tests/test_complex.py)."""
import numpy as np
import pytest
import torch

from pycc_b200 import planes
from pycc_b200.contract import Contractor
from pycc_b200.planes import Planes
from tests import emu


def P(z):
    z = np.asarray(z)
    if np.iscomplexobj(z):
        return Planes(torch.from_numpy(np.ascontiguousarray(z.real)), torch.from_numpy(np.ascontiguousarray(z.imag)))
    return Planes(torch.from_numpy(np.ascontiguousarray(z)), None)


def C(p):
    return p.re.numpy() + 1j * (0.0 if p.im is None else p.im.numpy())


@pytest.mark.parametrize("use3m", [False, True])
def test_cterm_and_cprod_match_complex_einsum(use3m):
    rng = np.random.default_rng(0)
    cz = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    rz = lambda *s: rng.standard_normal(s)
    sub = "imae,mbej->ijab"
    A_c, B_c, A_r, B_r = cz(3, 4, 5, 6), cz(4, 5, 6, 3), rz(3, 4, 5, 6), rz(4, 5, 6, 3)
    with emu.install():
        ct = Contractor()
        for A, B in ((A_c, B_c), (A_c, B_r), (A_r, B_c), (A_r, B_r)):
            want = np.einsum(sub, A, B)
            base = cz(3, 3, 5, 5)
            out = P(base)
            a = P(A) if np.iscomplexobj(A) else torch.from_numpy(A)            # real operands as plain tensors
            b = P(B) if np.iscomplexobj(B) else torch.from_numpy(B)
            planes.cterm(ct, -0.7, sub, a, b, out, use3m=use3m)
            assert np.abs(C(out) - (base - 0.7 * want)).max() < 1e-12
            got = planes.cprod(ct, 1.3, sub, a, b)
            assert np.abs(C(got) - 1.3 * want).max() < 1e-12
            assert (got.im is None) == (not np.iscomplexobj(want))
        # Planes with a missing imaginary plane behave as real operands
        out = P(cz(3, 3, 5, 5) * 0)
        planes.cterm(ct, 1.0, sub, P(A_r), P(B_c), out, use3m=use3m)
        assert np.abs(C(out) - np.einsum(sub, A_r, B_c)).max() < 1e-12


def test_complex_tau_and_copies():
    rng = np.random.default_rng(1)
    no, nv = 3, 5
    t1 = rng.standard_normal((no, nv)) + 1j * rng.standard_normal((no, nv))
    t2 = rng.standard_normal((no, no, nv, nv)) + 1j * rng.standard_normal((no, no, nv, nv))
    with emu.install():
        for a1, a2 in ((t1, t2), (t1.real, t2), (t1, t2.real), (t1.real, t2.real)):
            tau = planes.complex_tau(P(a1), P(a2))
            assert np.abs(C(tau) - (a2 + np.einsum("ia,jb->ijab", a1, a1))).max() < 1e-13
        F = rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8))
        sub = P(F).view(lambda x: x[:3, 3:])
        cp = planes.copy_planes(sub, 2.0)
        assert cp.re.is_contiguous() and np.abs(C(cp) - 2.0 * F[:3, 3:]).max() == 0.0
        cr = planes.copy_real(torch.from_numpy(F.real.copy())[:3, 3:], -1.0)
        assert np.abs(C(cr) + F.real[:3, 3:]).max() == 0.0 and float(cr.im.abs().max()) == 0.0
        assert P(F.real).full().im is not None
