"""Pin the CPU oracle against outputs of the reference's own, unmodified code
(tests/golden/ref_*.npz, written by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import ccsd_oracle as co
from oracle import triples_oracle as to
from pycc_b200.synthetic import full_eri, blocks_from_factor

TOL = 1e-12


def problem(syn):
    ERI = full_eri(syn)
    blocks = co.blocks_from_full(ERI, syn.no)
    return co.Problem(blocks, syn.F, syn.no), blocks


def test_blocks_from_factor_match_full(golden):
    g, syn = golden
    _, blocks = problem(syn)
    fb = blocks_from_factor(syn)
    for k in blocks:
        assert np.abs(blocks[k] - fb[k]).max() < 1e-13, k


def test_intermediates_and_residuals(golden):
    g, syn = golden
    P, _ = problem(syn)
    t1, t2 = g["rand_t1"], g["rand_t2"]
    r1, r2, inter, _ = P.residuals(syn.F, t1, t2, parts=True)
    for k, val in inter.items():
        assert np.abs(val - g["rand_" + k]).max() < TOL, k
    for (f1, f2) in ((1.0, 1.0), (1.0, 0.5), (0.5, 1.0)):
        assert np.abs(P.tau(t1, t2, f1, f2) - g["rand_tau_%g_%g" % (f1, f2)]).max() < TOL
    assert np.abs(r1 - g["rand_r1"]).max() < TOL
    assert np.abs(r2 - g["rand_r2"]).max() < TOL
    assert abs(P.cc_energy(syn.F, t1, t2) - float(g["rand_ecc"])) < TOL


def test_solve_trace_and_energy(golden):
    g, syn = golden
    P, _ = problem(syn)
    ecc, t1, t2, trace = co.solve_cc(P, 1e-12, 1e-12, 100)
    ref = g["trace_ecc_rms"]
    assert len(trace) == len(ref)
    tr = np.array(trace)
    assert np.abs(tr[:, 0] - ref[:, 0]).max() < 1e-12
    # rms is printed with 6 significant digits by the reference
    assert np.all(np.abs(tr[:, 1] - ref[:, 1]) <= 1e-5 * np.abs(ref[:, 1]) + 1e-15)
    assert abs(ecc - float(g["e_ccsd"])) < TOL
    assert np.abs(t1 - g["conv_t1"]).max() < 1e-11
    assert np.abs(t2 - g["conv_t2"]).max() < 1e-11


def test_diis(golden):
    g, syn = golden
    d1, d2 = g["diis_in_t1"], g["diis_in_t2"]
    diis = co.Diis(d1[0], d2[0], 4)
    x1, x2 = d1[0], d2[0]
    for n in range(1, 7):
        x1 = x1 + d1[n]
        x2 = x2 + d2[n]
        diis.add_error_vector(x1, x2)
        x1, x2 = diis.extrapolate(x1, x2)
        assert np.abs(x1 - g["diis_out_t1"][n - 1]).max() < 1e-11
        assert np.abs(x2 - g["diis_out_t2"][n - 1]).max() < 1e-11


def test_t3_tiles(golden):
    g, syn = golden
    P, b = problem(syn)
    t1, t2 = g["conv_t1"], g["conv_t2"]
    for n, (i, j, k) in enumerate(g["triples"]):
        W = to.t3c_ijk(i, j, k, t2, b["ovvv"], b["ooov"])
        V = W + to.t3d_ijk(i, j, k, t1, t2, b["oovv"], syn.F)
        assert np.abs(W - g["W3"][n]).max() < TOL
        assert np.abs(V - g["V3"][n]).max() < TOL
        Wd = to.t3c_ijk(i, j, k, t2, b["ovvv"], b["ooov"], syn.F, True)
        Vd = to.t3d_ijk(i, j, k, t1, t2, b["oovv"], syn.F, True)
        assert np.abs(Wd - g["t3c_denom"][n]).max() < TOL
        assert np.abs(Vd - g["t3d_denom"][n]).max() < TOL


def test_t_energy(golden):
    g, syn = golden
    P, b = problem(syn)
    t1, t2 = g["conv_t1"], g["conv_t2"]
    et = to.t_tjl(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
    assert abs(et - float(g["e_t_tjl"])) < TOL
    # the reference's own cross-formulation check (tests/test_005_ccsd_t_energy.py:30-36)
    assert abs(float(g["e_t_vikings"]) - float(g["e_t_tjl"])) < 1e-11
    assert abs(float(g["e_t_vikings_inverted"]) - float(g["e_t_tjl"])) < 1e-11
    assert abs(float(g["e_t_from_solve"]) - float(g["e_t_tjl"])) < 1e-12
    if syn.nv <= 10:
        ev = to.t_vikings(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
        assert abs(ev - float(g["e_t_vikings"])) < TOL


def test_abc_tiles_and_exchanged_role_energy(golden):
    """t3c_abc / t3d_abc (cctriples.py:75-105, 149-173) against the reference's own outputs, and the Lee-Rendell bracket
    applied per (a,b,c) tile with exchanged roles -- the formulation of the fused CUDA kernel -- against the reference's
    E(T) (t_tjl == t_vikings == t_vikings_inverted, its tests/test_005_ccsd_t_energy.py:30-36)."""
    g, syn = golden
    b = blocks_from_factor(syn)
    t1, t2 = g["conv_t1"], g["conv_t2"]
    assert np.abs(to.t3c_abc(2, 1, 0, t2, b["ovvv"], b["ooov"], syn.F, True) - g["t3c_abc_210"]).max() < TOL
    assert np.abs(to.t3d_abc(2, 1, 0, t1, t2, b["oovv"], syn.F, True) - g["t3d_abc_210"]).max() < TOL
    assert abs(to.t_tjl_abc(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"]) - float(g["e_t_tjl"])) < TOL
