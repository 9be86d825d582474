"""``IntegralReference.from_psi4`` without psi4: a stub module that answers the handful of calls the constructor makes
(wavefunction.py:304-315, hamiltonian.py:54-67) from the H2O/STO-3G fixture, and records that the n^4 MO array is never
requested (``mo_eri`` must not be called; ``ao_eri`` is)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200.wavefunction import IntegralReference, resolve_reference
from tests import emu
from tests.golden import gto

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Wfn:
    def __init__(self, g):
        self.g = g

    def Ca_subset(self, a, b):
        return self.g["C"]

    def Fa_subset(self, a):
        return self.g["F_ao"]

    def basisset(self):
        return "basis"

    def frzcpi(self):
        return [1]

    def doccpi(self):
        return [5]

    def energy(self):
        return -74.9


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        with emu.install():
            yield torch.device("cpu")
    else:
        yield torch.device("cuda:0")


def test_from_psi4_goes_through_the_ao_integrals(dev, monkeypatch):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "h2o_sto3g.npz")))
    calls = []

    class Mints:
        def __init__(self, basis):
            assert basis == "basis"

        def ao_eri(self):
            calls.append("ao_eri")
            return gto.unpack_eri(g["eri_packed"], 7)

        def mo_eri(self, *a):
            raise AssertionError("the n^4 MO array must never be requested")

    psi4 = types.ModuleType("psi4")
    psi4.core = types.SimpleNamespace(MintsHelper=Mints)
    monkeypatch.setitem(sys.modules, "psi4", psi4)
    ref = resolve_reference(_Wfn(g))                      # what CCwfn(scf_wfn) does with a psi4 wavefunction
    assert isinstance(ref, IntegralReference) and calls == ["ao_eri"] and ref.eref == -74.9
    e = pycc_b200.ccwfn(ref, model="CCSD(T)", device="GPU", quiet=True).solve_cc(1e-12, 1e-12, 75)
    assert abs(float(e) + 0.0707167876524093) < 1e-11     # pycc/tests/test_044_ccsd_t_gpu.py:37
