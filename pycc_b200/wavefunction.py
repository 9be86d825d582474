"""What the CC solver needs from an SCF reference (reference: pycc/wavefunction.py:111-188, 293-332).

pycc takes a psi4 ``Wavefunction`` and builds the full MO Hamiltonian through psi4's MintsHelper.
psi4 is an optional dependency here: anything that can provide (F, <pq|rs> or its six blocks, no,
nfzc, E_ref) works, through :class:`IntegralReference`:

* ``IntegralReference.from_arrays(F, ERI, no, nfzc, eref)`` -- host arrays (e.g. dumped from psi4);
* ``IntegralReference.from_synthetic(syn)``                 -- factorised synthetic integrals, contracted
                                                              into blocks on the device;
* ``IntegralReference.from_ao(F_ao, eri_ao, C, no, nfzc)``    -- AO-basis arrays; the AO -> MO transformation runs on
                                                              the device, block by block (hamiltonian.py:54-70);
* ``IntegralReference.from_psi4(scf_wfn)``                  -- same calls as pycc/hamiltonian.py:58-68.
A raw psi4 wavefunction or a ``Synthetic`` passed to ``CCwfn`` is converted automatically.
"""
from __future__ import annotations

import numpy as np

from .hamiltonian import BlockHamiltonian
from .synthetic import Synthetic


class IntegralReference:
    def __init__(self, eref=0.0):
        self.eref = float(eref)
        self._make = None
        self._owns = True          # the Hamiltonian is built for this wavefunction (False: a caller's object)

    def energy(self):
        return self.eref

    def hamiltonian(self, device, comm=None, mixed=False):
        """``mixed``: the Hamiltonian will be used by precision='MP' -- <ab|ef> is kept as TF32 planes only."""
        H = self._make(device, comm, mixed)
        H.owned = self._owns
        if mixed and H.vvvv_planes is None and (H.has("vvvv") or H.vvvv_packed is not None):
            H.to_mixed(drop=self._owns)
        elif not mixed and self._owns and H.has("vvvv"):
            H.packed()                     # pack now; a large FP64 block is released (BlockHamiltonian.keep_vvvv_bytes)
        return H

    @classmethod
    def from_arrays(cls, F, ERI, no, nfzc=0, eref=0.0):
        r = cls(eref)
        r._make = lambda device, comm, mixed: BlockHamiltonian.from_full(F, ERI, no, nfzc, device)
        return r

    @classmethod
    def from_blocks(cls, F, blocks, no, nfzc=0, eref=0.0):
        r = cls(eref)
        r._make = lambda device, comm, mixed: BlockHamiltonian(F, blocks, no, nfzc, device)
        return r

    @classmethod
    def from_synthetic(cls, syn):
        r = cls(syn.eref)

        def make(device, comm, mixed):
            a_range = None if comm is None else comm.a_range(syn.nv)
            return BlockHamiltonian.from_factor(syn, device, a_range=a_range, mixed=mixed)
        r._make = make
        return r

    @classmethod
    def from_ao(cls, F_ao, eri_ao, C, no, nfzc=0, eref=0.0, stream_ao=None):
        """AO-basis inputs (Fock matrix, chemist-order repulsion integrals, MO coefficients): the AO -> MO
        transformation runs on the device, straight into the six blocks (BlockHamiltonian.from_ao).  ``eri_ao`` may be a
        host array / np.memmap too large for the device: it is then streamed in slabs (``stream_ao``: None = decide
        from the free device memory)."""
        r = cls(eref)

        def make(device, comm, mixed):
            nv = np.asarray(C).shape[1] - no - nfzc
            a_range = None if comm is None else comm.a_range(nv)
            return BlockHamiltonian.from_ao(F_ao, eri_ao, C, no, nfzc, device, a_range=a_range, stream_ao=stream_ao)
        r._make = make
        return r

    @classmethod
    def from_psi4(cls, scf_wfn):
        """From a psi4 ``Wavefunction``.  The reference asks psi4 for the full n^4 MO array (``mints.mo_eri(C,C,C,C)``,
        hamiltonian.py:67) and forms ``ERI`` and ``L`` from it on the host -- 2 x 107 GB at o=40, v=300.  Here only what
        psi4 alone can provide is taken from it -- the AO Fock matrix, the AO repulsion integrals ``mints.ao_eri()``
        (chemist order) and the MO coefficients, exactly the inputs hamiltonian.py:54-67 starts from -- and the AO -> MO
        transformation runs on the device, block by block (``from_ao``): the n^4 MO array never exists."""
        import psi4                                   # optional dependency
        C = np.asarray(scf_wfn.Ca_subset("AO", "ALL"))
        F_ao = np.asarray(scf_wfn.Fa_subset("AO"))
        mints = psi4.core.MintsHelper(scf_wfn.basisset())
        eri_ao = np.asarray(mints.ao_eri())
        nfzc = int(sum(scf_wfn.frzcpi()))
        no = int(sum(scf_wfn.doccpi())) - nfzc
        return cls.from_ao(F_ao, eri_ao, C, no, nfzc, scf_wfn.energy())


def resolve_reference(x):
    if isinstance(x, IntegralReference):
        return x
    if isinstance(x, Synthetic):
        return IntegralReference.from_synthetic(x)
    if isinstance(x, BlockHamiltonian):
        r = IntegralReference(0.0)
        r._make = lambda device, comm, mixed: x
        r._owns = False
        return r
    if hasattr(x, "frzcpi") and hasattr(x, "Ca_subset"):
        return IntegralReference.from_psi4(x)
    raise TypeError("CCwfn needs a psi4 Wavefunction, an IntegralReference, a Synthetic or a BlockHamiltonian "
                    "(got %s)" % type(x).__name__)
