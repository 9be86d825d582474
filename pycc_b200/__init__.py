"""pycc_b200 -- B200-native closed-shell RHF-CCSD / CCSD(T) behind the pycc API.

    from pycc_b200 import ccwfn
    cc = ccwfn(scf_wfn, model='CCSD(T)', device='GPU')
    ecc = cc.solve_cc(1e-10, 1e-10)

Importing the package never touches the GPU; the first kernel call loads ``libb200cc.so`` and
raises if it (or a CUDA device) is missing -- there is no CPU fallback.
"""
from .exceptions import PyCCError, PyCCWarning, InvalidKeywordError
from .ccwfn import CCwfn, ccwfn
from . import cctriples
from .cchbar import cchbar
from .cclambda import cclambda
from .utils import helper_diis
from .device import DeviceManager, ContractionBackend
from .wavefunction import IntegralReference
from .hamiltonian import BlockHamiltonian
from .synthetic import make_synthetic

__all__ = ["CCwfn", "ccwfn", "cctriples", "cchbar", "cclambda", "helper_diis", "DeviceManager", "ContractionBackend",
           "IntegralReference", "BlockHamiltonian", "make_synthetic", "PyCCError", "PyCCWarning",
           "InvalidKeywordError"]
