"""CPU ORACLE (test infrastructure, NOT a product path) for the AO -> MO integral staging -- SURVEY.md 8(f) "next" #3.

Plain-numpy restatement of CrawfordGroup/pycc ``Hamiltonian.__init__`` (hamiltonian.py:36-70):

    F_pq    = C_mu,p F_mu,nu C_nu,q                                   hamiltonian.py:55-56
    (pr|qs) = mo_eri(C, C, C, C)        (chemist)                     hamiltonian.py:67
    <pq|rs> = (pr|qs)                   (swapaxes(1,2))               hamiltonian.py:68
    L_pqrs  = 2 <pq|rs> - <pq|sr>                                     hamiltonian.py:70

PARITY UNPINNED against the reference run: ``mo_eri`` lives in psi4 (third-party, not under /root/reference,
unpinned -- CI installs conda-forge latest, .github/workflows/CI.yaml:64; not installable offline).  Its published
definition, also stated in the reference's own docstring (hamiltonian.py:37-43), is the four-index transformation
restated here; everything downstream of the MO integrals IS pinned (tests/golden/ref_*.npz).

Only ``tests/`` may import this module, as the checker.
"""
from __future__ import annotations

import numpy as np


def mo_hamiltonian(F_ao, eri_ao, C):
    """(F, ERI, L) in the MO basis from AO arrays; ``eri_ao`` in chemist order (mu lam|nu sig)."""
    F = C.T @ F_ao @ C
    chem = np.einsum("mp,lr,nq,st,mlns->prqt", C, C, C, C, eri_ao, optimize=True)     # (pr|qs)
    ERI = np.ascontiguousarray(chem.swapaxes(1, 2))                                     # <pq|rs>
    L = 2.0 * ERI - ERI.swapaxes(2, 3)
    return F, ERI, L
