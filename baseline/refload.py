"""Load and drive the UNMODIFIED reference (CrawfordGroup/pycc) -- baseline / test infrastructure, never imported by
``pycc_b200``.

Used by ``bench.py --impl reference``, by the ``cpu_baseline`` leg of ``bench.py`` and by the boundary tests that run
the reference's own ``CCwfn.residuals`` with this package's contraction backend plugged into its ``contract`` seam.

The reference's source files are executed as they are, from ``$PYCC_REFERENCE``, else ``/root/reference`` (build
container), else ``baseline/_ref`` (the offline install that travels to the GPU box, see ``install_ref.py``).  Three
imports it needs are absent from this image and are shimmed at load time:

* ``psi4``        -- a stub module; never called on this path (integrals are handed in);
* ``opt_einsum``  -- ``contract(sub, *ops)`` = the exact ``numpy.einsum(..., optimize=True)`` (``torch.einsum`` for
                     tensors), i.e. einsum -> tensordot -> BLAS, which is what opt_einsum lowers these two-operand
                     contractions to (pycc/device.py:68,84);
* the ``pycc`` package object -- pre-registered empty so ``pycc/__init__.py`` (which pulls qcelemental through
                     ccresponse) is skipped and ``pycc.ccwfn`` is imported directly.

``reference_wfn`` fills a ``CCwfn`` exactly as pycc/ccwfn.py:195-211 and pycc/wavefunction.py:153-169 do, with
``H.ERI`` / ``H.L`` either the reference's full n^4 arrays or :class:`HostBlockIntegrals`, an object that serves the
reference's own slicing syntax (``ERI[o,v,v,o]``, ``L[o,o,v,v]``) from the six unique Dirac blocks so that sizes
whose 2 x n^4 doubles do not fit in host memory can still run the reference's code unmodified.
"""
import importlib
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STORED = ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")


def reference_root():
    """Directory that contains the reference's ``pycc`` package, or None."""
    for cand in (os.environ.get("PYCC_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref")):
        if cand and os.path.exists(os.path.join(cand, "pycc", "ccwfn.py")):
            return cand
    return None


_LOADED = {}


def load_reference(root=None):
    """Import the reference's hot-path modules; returns a namespace (ccwfn, cctriples, utils, device, root)."""
    root = root or reference_root()
    if root is None:
        raise ImportError("the reference is neither at $PYCC_REFERENCE, /root/reference nor baseline/_ref "
                          "(run baseline/install_ref.py in the build container)")
    if root in _LOADED:
        return _LOADED[root]
    for name in ("psi4", "psi4.core"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["psi4"].core = sys.modules["psi4.core"]
    if "opt_einsum" not in sys.modules:
        oe = types.ModuleType("opt_einsum")

        def contract(sub, *ops, **kw):
            if any(type(x).__module__.startswith("torch") for x in ops):
                import torch
                return torch.einsum(sub, *ops)
            return np.einsum(sub, *ops, optimize=True)
        oe.contract = contract
        sys.modules["opt_einsum"] = oe
    pkg = types.ModuleType("pycc")
    pkg.__path__ = [os.path.join(root, "pycc")]
    sys.modules["pycc"] = pkg
    ns = types.SimpleNamespace(root=root)
    for mod in ("ccwfn", "cctriples", "utils", "device"):
        setattr(ns, mod, importlib.import_module("pycc." + mod))
    _LOADED[root] = ns
    return ns


# ---- the eight index permutations that leave a real <pq|rs> invariant ------------------------------------------------
def _symmetries():
    sym, frontier = set(), [(0, 1, 2, 3)]
    gens = [(2, 1, 0, 3), (0, 3, 2, 1), (1, 0, 3, 2)]
    while frontier:
        g = frontier.pop()
        if g in sym:
            continue
        sym.add(g)
        frontier.extend(tuple(g[h[t]] for t in range(4)) for h in gens)
    return sorted(sym)


_SYM = _symmetries()


class HostBlockIntegrals:
    """``ERI`` / ``L`` of the reference's Hamiltonian (hamiltonian.py:67-70) served from the six unique Dirac blocks
    (host numpy arrays): ``X[o,v,v,o]`` with the wavefunction's own o / v slices returns the block as a (transposed)
    view; ``L`` blocks (2<pq|rs> - <pq|sr>) are formed on first use and cached."""

    def __init__(self, blocks, o, v, kind="ERI", eri=None):
        self.blocks, self.o, self.v, self.kind = blocks, o, v, kind
        self.eri = self if kind == "ERI" else eri
        self._cache = {}

    def _pattern(self, key):
        pat = ""
        for s in key:
            if s == self.o:
                pat += "o"
            elif s == self.v:
                pat += "v"
            else:
                raise KeyError("block integrals are indexed with the wavefunction's o / v slices (got %r)" % (s,))
        return pat

    def __getitem__(self, key):
        pat = self._pattern(key)
        if self.kind == "ERI":
            for g in _SYM:
                for name in STORED:
                    if name in self.blocks and all(pat[t] == name[g[t]] for t in range(4)):
                        blk = self.blocks[name]
                        return blk.transpose(g) if isinstance(blk, np.ndarray) else blk.permute(*g)   # numpy / torch
            raise KeyError("no stored block for pattern %r" % pat)
        if pat not in self._cache:
            a = self.eri[key]
            b = self.eri[(key[0], key[1], key[3], key[2])].swapaxes(2, 3)
            self._cache[pat] = 2.0 * a - b
        return self._cache[pat]


def host_blocks(syn, names=STORED, log=None):
    """The six Dirac blocks of a factorised synthetic problem as host arrays IN THE REFERENCE'S MEMORY LAYOUT: psi4's
    ``mo_eri`` returns the chemist array (pr|qs) C-contiguous and the reference takes ``.swapaxes(1, 2)`` of it as
    ``ERI`` (hamiltonian.py:67-68), so every ``ERI[...]`` slice it contracts is a strided view whose memory order is
    [p, r, q, s].  Each block is therefore built as the contiguous chemist product (pr|qs) = sum_P B[P,p,r] B[P,q,s]
    (one GEMM, no transposition) and returned as that same ``swapaxes(1, 2)`` view -- tensordot's transposition copies
    then cost here what they cost the reference."""
    B, no, n = syn.B, syn.no, syn.n
    sl = {"o": slice(0, no), "v": slice(no, n)}
    naux = B.shape[0]
    out = {}
    for name in names:
        p, q, r, s = (sl[c] for c in name)
        Bpr = np.ascontiguousarray(B[:, p, r]).reshape(naux, -1)
        Bqs = np.ascontiguousarray(B[:, q, s]).reshape(naux, -1)
        dims = (p.stop - p.start, r.stop - r.start, q.stop - q.start, s.stop - s.start)
        chem = np.empty((Bpr.shape[1], Bqs.shape[1]))
        step = max(1, (1 << 28) // max(Bqs.shape[1], 1))           # row panels of <= 2 GB
        for r0 in range(0, Bpr.shape[1], step):
            np.matmul(Bpr[:, r0:r0 + step].T * syn.scale, Bqs, out=chem[r0:r0 + step])
            if log is not None and name == "vvvv":
                log("  (ae|bf) rows %d/%d on the host" % (min(r0 + step, Bpr.shape[1]), Bpr.shape[1]))
        out[name] = chem.reshape(dims).swapaxes(1, 2)
    return out


def reference_wfn(ref, F, no, ERI, L=None, blocks=None, nfzc=0, model="CCSD", device="CPU", precision="DP",
                  contract=None):
    """A reference ``CCwfn`` with every attribute its constructor sets (ccwfn.py:195-211, wavefunction.py:153-169),
    without psi4.  ``ERI`` may be the full n^4 array (then ``L`` is formed as hamiltonian.py:70 does) or a
    :class:`HostBlockIntegrals`.  ``contract``: a replacement for the reference's contraction backend -- the seam of
    pycc/device.py:64 -- e.g. ``pycc_b200.ContractionBackend('GPU')``."""
    CCwfn = ref.ccwfn.CCwfn
    w = CCwfn.__new__(CCwfn)
    w.model = w.method = model
    w.make_t3_density = False
    w.store_triples = False
    w.local = None
    w.orbital_basis = "spatial"
    w.eref = 0.0
    n = F.shape[0]
    w.no, w.nfzc, w.nmo = int(no), int(nfzc), n
    w.nv = n - no - nfzc
    w.o, w.v = slice(nfzc, nfzc + no), slice(nfzc + no, n)
    mgr = ref.device.DeviceManager(device=device, precision=precision)
    w.device_manager = mgr
    w.device, w.device0, w.device1 = mgr.device, mgr.device0, mgr.device1
    w.precision = mgr.precision
    w.contract = mgr.contract if contract is None else contract
    H = types.SimpleNamespace()
    H.F = F.copy() if hasattr(F, "copy") else F.clone()
    H.eps = np.diagonal(np.asarray(F)).copy()
    if ERI is None:
        ERI = HostBlockIntegrals(blocks, w.o, w.v, "ERI")
    H.ERI = ERI
    if L is None:
        L = (HostBlockIntegrals(ERI.blocks, w.o, w.v, "L", eri=ERI) if isinstance(ERI, HostBlockIntegrals)
             else 2.0 * ERI - ERI.swapaxes(2, 3))
    H.L = L
    w.H = H
    if isinstance(L, HostBlockIntegrals):
        # the reference forms L once at construction (hamiltonian.py:70): do the same for the blocks `residuals` reads
        for pat in ("ovvv", "oovv", "ooov", "ovvo", "oovo"):
            L[tuple(w.o if c == "o" else w.v for c in pat)]
    eo, ev = H.eps[w.o], H.eps[w.v]
    w.Dijab = eo.reshape(-1, 1, 1, 1) + eo.reshape(-1, 1, 1) - ev.reshape(-1, 1) - ev
    w.Dia = eo.reshape(-1, 1) - ev
    w.t1 = np.zeros((w.no, w.nv))
    if device == "GPU" and isinstance(ERI, HostBlockIntegrals):
        # device-resident blocks (torch tensors on the compute device): the reference's device='GPU' path WITHOUT its
        # per-call host->device copies of the ERI slices (device.py:70-74) -- the library-formulation bar of bench.py
        H.F, H.eps = mgr.seed_compute(H.F), mgr.seed_compute(H.eps)
        w.Dijab, w.Dia, w.t1 = mgr.seed_compute(w.Dijab), mgr.seed_compute(w.Dia), mgr.seed_compute(w.t1)
        w.t2 = ERI[w.o, w.o, w.v, w.v] / w.Dijab
        return w
    w.t2 = np.array(ERI[w.o, w.o, w.v, w.v]) / w.Dijab
    if device == "GPU":
        # what wavefunction.py:165-169 and ccwfn.py:195-211 do for device='GPU': F / eps / denominators / amplitudes
        # become torch tensors on the compute device, the n^4 ERI / L torch tensors on the storage device (host)
        H.F, H.eps = mgr.seed_compute(H.F), mgr.seed_compute(H.eps)
        H.ERI, H.L = mgr.seed_store(np.ascontiguousarray(H.ERI)), mgr.seed_store(np.ascontiguousarray(H.L))
        w.Dijab, w.Dia = mgr.seed_compute(w.Dijab), mgr.seed_compute(w.Dia)
        w.t1, w.t2 = mgr.seed_compute(w.t1), mgr.seed_compute(w.t2)
    return w


class _IterationClock:
    """A stdout stand-in that time-stamps the reference's own per-iteration prints (``Iter n: CC Ecorr = ...``,
    ccwfn.py:264,288) -- the iteration times of ``solve_cc`` without touching its code."""

    def __init__(self, echo=None):
        self.stamps, self.lines, self.echo, self._buf = [], [], echo, ""

    def write(self, s):
        self._buf += s
        while "\n" in self._buf:
            line, self._buf = self._buf.split("\n", 1)
            if line.startswith("Iter"):
                self.stamps.append(time.perf_counter())
                self.lines.append(line)
                if self.echo is not None:
                    self.echo(line)
        return len(s)

    def flush(self):
        pass


def timed_solve_cc(w, iterations, max_diis=8, start_diis=1, echo=None):
    """Run the reference's own ``solve_cc`` (ccwfn.py:216-319) for exactly ``iterations`` iterations -- convergence
    thresholds that can never be met, ``maxiter = iterations`` -- and return (seconds of each iteration, energies).
    Iteration n is the interval between the reference's own ``Iter n-1`` and ``Iter n`` prints: residuals (321-372),
    Jacobi update + rms (281-284), energy (286) and, for n > 1, the DIIS add / extrapolate (317-319) that followed
    the previous print."""
    clock = _IterationClock(echo)
    keep = sys.stdout
    sys.stdout = clock
    try:
        w.solve_cc(-1.0, -1.0, int(iterations), max_diis, start_diis)
    finally:
        sys.stdout = keep
    st = clock.stamps                                     # Iter 0 (MP2) .. Iter n
    secs = [st[k] - st[k - 1] for k in range(1, len(st))]
    energies = []
    for line in clock.lines:
        try:
            energies.append(float(line.split("=")[1].split()[0]))
        except Exception:
            energies.append(float("nan"))
    return secs, energies
