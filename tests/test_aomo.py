"""AO -> MO integral staging (reference hamiltonian.py:54-70): BlockHamiltonian.from_ao builds F and the
six Dirac blocks on the device from AO arrays.  Checked against the numpy oracle (oracle/aomo_oracle.py), through the
reference's own slicing syntax (H.ERI[o,v,v,o], H.L[...]), and end to end: a CCSD(T) solved from AO inputs must equal
the one solved from the MO arrays.  `emu` / `cuda` as in test_ccsd.py."""
import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200.hamiltonian import BlockHamiltonian
from pycc_b200.synthetic import make_synthetic, full_eri
from pycc_b200.wavefunction import IntegralReference
from oracle import aomo_oracle as ao
from tests import emu


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        with emu.install():
            yield torch.device("cpu")
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        yield torch.device("cuda:0")


def ao_problem(no, nv, nfzc, nbf, seed):
    """A synthetic AO problem: an 8-fold symmetric (mu lam|nu sig) and symmetric F_ao over nbf functions, and a
    rectangular C (nbf x nmo, nmo <= nbf as after removing linear dependencies) with orthonormal columns."""
    rng = np.random.default_rng(seed)
    nmo = nfzc + no + nv
    syn = make_synthetic(2, nbf - 2, seed=seed)                    # only its symmetric factor is used
    chem = np.einsum("Ppq,Prs->pqrs", syn.B, syn.B, optimize=True) * 0.002
    Q, _ = np.linalg.qr(rng.standard_normal((nbf, nbf)))
    C = np.ascontiguousarray(Q[:, :nmo])
    eps = np.concatenate((np.linspace(-3.0, -0.6, nfzc + no), np.linspace(0.5, 3.0, nv)))
    noise = rng.uniform(-0.01, 0.01, (nmo, nmo))
    Fmo = np.diag(eps) + 0.5 * (noise + noise.T) * (1.0 - np.eye(nmo))
    # an AO Fock matrix whose MO image is Fmo (plus a component outside span(C), which must not matter)
    R = rng.standard_normal((nbf, nbf))
    P = np.eye(nbf) - C @ C.T
    F_ao = C @ Fmo @ C.T + P @ (R + R.T) @ P
    return F_ao, chem, C


@pytest.mark.parametrize("no,nv,nfzc,nbf", [(3, 5, 0, 8), (2, 6, 1, 11), (4, 7, 2, 13)])
def test_blocks_match_oracle(dev, no, nv, nfzc, nbf):
    F_ao, chem, C = ao_problem(no, nv, nfzc, nbf, seed=nbf)
    F, ERI, L = ao.mo_hamiltonian(F_ao, chem, C)
    H = BlockHamiltonian.from_ao(F_ao, chem, C, no, nfzc, dev, chunk_bytes=8 * nbf ** 3 * 2)   # two a-rows per chunk
    assert (H.no, H.nv, H.nfzc) == (no, nv, nfzc)
    assert np.abs(H.F.cpu().numpy() - F).max() < 1e-12
    o, v = H.o, H.v
    for pat in ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv", "vvvo", "ovvo", "vovv", "ovoo"):
        key = tuple(o if c == "o" else v for c in pat)
        assert np.abs(H.ERI[key].cpu().numpy() - ERI[key]).max() < 1e-12, pat
    for pat in ("oovv", "ovvv", "ooov", "ovvo"):
        key = tuple(o if c == "o" else v for c in pat)
        assert np.abs(H.L[key].cpu().numpy() - L[key]).max() < 1e-12, pat


def test_a_sharded_ladder_block(dev):
    no, nv, nfzc, nbf = 3, 6, 1, 12
    F_ao, chem, C = ao_problem(no, nv, nfzc, nbf, seed=3)
    _, ERI, _ = ao.mo_hamiltonian(F_ao, chem, C)
    v = slice(nfzc + no, nfzc + no + nv)
    for a_range in ((0, 2), (2, 6), (5, 6)):
        H = BlockHamiltonian.from_ao(F_ao, chem, C, no, nfzc, dev, a_range=a_range)
        want = ERI[v, v, v, v][a_range[0]:a_range[1]]
        assert np.abs(H.block("vvvv").cpu().numpy() - want).max() < 1e-12


def test_ccsd_t_from_ao_equals_from_mo_arrays(dev):
    no, nv, nfzc, nbf = 3, 6, 1, 12
    F_ao, chem, C = ao_problem(no, nv, nfzc, nbf, seed=5)
    F, ERI, _ = ao.mo_hamiltonian(F_ao, chem, C)
    e_ao = pycc_b200.ccwfn(IntegralReference.from_ao(F_ao, chem, C, no, nfzc), model="CCSD(T)", device="GPU",
                           quiet=True).solve_cc(1e-11, 1e-11)
    e_mo = pycc_b200.ccwfn(IntegralReference.from_arrays(F, ERI, no, nfzc), model="CCSD(T)", device="GPU",
                           quiet=True).solve_cc(1e-11, 1e-11)
    assert e_ao is not None and abs(float(e_ao) - float(e_mo)) < 1e-11
    assert abs(float(e_ao)) > 1e-6


def test_shape_error(dev):
    from pycc_b200._lib import B200ccError
    F_ao, chem, C = ao_problem(2, 3, 0, 6, seed=1)
    with pytest.raises(B200ccError):
        BlockHamiltonian.from_ao(F_ao, chem[:5], C, 2, 0, dev)


@pytest.mark.parametrize("slab_rows", [1, 3, 100])
def test_host_streamed_ao_equals_resident(dev, slab_rows):
    """nbf too large for the AO array to sit in HBM: it stays on the host and every first quarter transformation sweeps
    it in slabs of the first AO index (ragged last slab; one sweep per <ab|ef> row chunk)."""
    no, nv, nfzc, nbf = 3, 6, 1, 11
    F_ao, chem, C = ao_problem(no, nv, nfzc, nbf, seed=7)
    Hr = BlockHamiltonian.from_ao(F_ao, chem, C, no, nfzc, dev, stream_ao=False)
    Hs = BlockHamiltonian.from_ao(F_ao, chem, C, no, nfzc, dev, stream_ao=True, slab_bytes=8 * nbf ** 3 * slab_rows,
                                  chunk_bytes=8 * nbf ** 3 * 4, a_range=(1, 6))
    for name in ("oooo", "ooov", "oovv", "ovov", "ovvv"):
        assert np.abs((Hs.block(name) - Hr.block(name)).cpu().numpy()).max() < 1e-13, name
    assert np.abs((Hs.block("vvvv") - Hr.block("vvvv")[1:6]).cpu().numpy()).max() < 1e-13
    assert np.abs((Hs.F - Hr.F).cpu().numpy()).max() == 0.0


def test_host_streamed_ao_from_memmap(dev, tmp_path):
    no, nv, nfzc, nbf = 2, 5, 0, 8
    F_ao, chem, C = ao_problem(no, nv, nfzc, nbf, seed=9)
    mm = np.memmap(tmp_path / "ao.bin", dtype=np.float64, mode="w+", shape=chem.shape)
    mm[:] = chem
    mm.flush()
    _, ERI, _ = ao.mo_hamiltonian(F_ao, chem, C)
    H = BlockHamiltonian.from_ao(F_ao, np.memmap(tmp_path / "ao.bin", dtype=np.float64, mode="r", shape=chem.shape), C,
                                 no, nfzc, dev, stream_ao=True, slab_bytes=8 * nbf ** 3 * 3)
    o, v = H.o, H.v
    for pat in ("oooo", "ovvv", "vvvv", "ovvo"):
        key = tuple(o if c == "o" else v for c in pat)
        assert np.abs(H.ERI[key].cpu().numpy() - ERI[key]).max() < 1e-12, pat


def test_ccsd_from_host_streamed_ao(dev):
    """the whole chain with the AO tensor kept on the host: same energy as with the resident tensor"""
    no, nv, nfzc, nbf = 3, 6, 1, 12
    F_ao, chem, C = ao_problem(no, nv, nfzc, nbf, seed=5)
    e_s = pycc_b200.ccwfn(IntegralReference.from_ao(F_ao, chem, C, no, nfzc, stream_ao=True), model="CCSD",
                          device="GPU", quiet=True).solve_cc(1e-11, 1e-11)
    e_r = pycc_b200.ccwfn(IntegralReference.from_ao(F_ao, chem, C, no, nfzc, stream_ao=False), model="CCSD",
                          device="GPU", quiet=True).solve_cc(1e-11, 1e-11)
    assert e_s is not None and abs(float(e_s) - float(e_r)) < 1e-13
