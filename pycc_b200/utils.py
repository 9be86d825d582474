"""DIIS extrapolation on the device + the solver's text output (reference: pycc/utils.py:200-361).

``helper_diis`` keeps the reference's interface and Pulay algebra (utils.py:272-361) but not its cost:
each stored state is ONE flat device buffer (t1 then t2), the error-overlap matrix B is cached and
only its new row is computed -- one fused multi-dot pass over the newest error vector
(``b200cc_multi_dot``) instead of m(m+1)/2 separate dots -- and the extrapolant is one fused
multi-axpy pass (``b200cc_multi_axpy``).  The (m+1)x(m+1) Pulay system is solved on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K


# ---- tensor helpers with the reference's names (utils.py:14-191).  pycc dispatches these on numpy-vs-torch; here
# every amplitude is a CUDA tensor, so they are thin plumbing over torch allocation / the package's own kernels.
def zeros_like(a):
    return torch.zeros_like(a)


def zeros(shape, like):
    return torch.zeros(shape, dtype=like.dtype, device=like.device)


def real_zeros(shape, like):
    return torch.zeros(shape, dtype=like.real.dtype, device=like.device)


def clone(a, device=None):
    out = torch.empty(tuple(a.shape), dtype=a.dtype, device=a.device)
    K.strided_axpby(out, a, 1.0, 0.0)
    return out.to(device) if device is not None else out


def diag(a):
    return torch.diagonal(a) if a.dim() == 2 else torch.diag(a)


def dot(a, b):
    """1-D dot product on the device (b200cc_multi_dot); returns a 0-d tensor."""
    return K.multi_dot(a.contiguous().view(-1), [b.contiguous().view(-1)])[0]


def sqrt(a):
    return torch.sqrt(a) if isinstance(a, torch.Tensor) else np.sqrt(a)


def absolute(a):
    return torch.abs(a) if isinstance(a, torch.Tensor) else np.abs(a)


def conj(a):
    """complex conjugate (utils.py:157-161); a view flip for torch tensors, no arithmetic kernel involved"""
    return torch.conj(a) if isinstance(a, torch.Tensor) else np.conj(a)


def solve(A, b):
    """(m+1)x(m+1) Pulay systems: solved on the host (utils.py:348 uses torch/np.linalg.solve)."""
    if isinstance(A, torch.Tensor):
        return torch.from_numpy(np.linalg.solve(A.cpu().numpy(), b.cpu().numpy())).to(A.device)
    return np.linalg.solve(A, b)


def reshape(a, shape):
    return a.reshape(shape)


def concatenate(arrays):
    return torch.cat(arrays) if isinstance(arrays[0], torch.Tensor) else np.concatenate(arrays)


class helper_diis(object):
    """Pulay DIIS with the reference's interface (utils.py:257-361): ``helper_diis(t1, t2, max_diis, precision)``,
    ``add_error_vector(t1, t2)``, ``extrapolate(t1, t2) -> (t1, t2)``.  Iterates and error vectors are flat device
    buffers; B is cached and only its newest row is computed (one fused multi-dot pass), the extrapolant is one fused
    multi-axpy.

    ``comm`` (parallel.Comm, more than one rank): the history, the dots and the extrapolation are SHARDED over the rows
    i of t2 (``comm.occ_range``): a rank stores and touches only [t1 | t2[i0:i1]]; the B-matrix dots of the t2 rows are
    summed over the ranks (t1, replicated, is counted once), every rank solves the same (m+1) x (m+1) system, writes
    its rows of the extrapolant into the caller's t2 and the rows are all-gathered."""

    def __init__(self, t1, t2, max_diis, precision='DP', comm=None):
        self.max_diis = max_diis
        self.precision = precision
        self.comm = comm if (comm is not None and comm.size > 1) else None
        self.rows = self.comm.occ_range(t2.shape[0]) if self.comm is not None else (0, t2.shape[0])
        i0, i1 = self.rows
        self.shapes = (tuple(t1.shape), (i1 - i0,) + tuple(t2.shape[1:]))
        self.n1 = t1.numel()
        self.n2 = (i1 - i0) * (t2.numel() // max(t2.shape[0], 1))
        self.diis_size = 0
        self.last_coefficients = None
        if max_diis == 0:
            return
        self.old = self._flat(t1, t2)
        self.vals = [self._clone(self.old)]
        self.errors = []
        self._reserve(self.old)
        self._ids = []              # identity of each error vector (for the B cache)
        self._next_id = 0
        self._dots = {}

    # -- flat storage -----------------------------------------------------------------------------
    def _flat(self, t1, t2):
        i0, i1 = self.rows
        buf = torch.empty(self.n1 + self.n2, dtype=t2.dtype, device=t2.device)
        K.strided_axpby(buf[:self.n1].view(self.shapes[0]), t1, 1.0, 0.0)
        if self.n2:
            K.strided_axpby(buf[self.n1:].view(self.shapes[1]), t2[i0:i1], 1.0, 0.0)
        return buf

    def _reserve(self, like):
        """The history grows by two vectors per iteration until max_diis is reached; each would be a fresh cudaMalloc of
        an o^2v^2 buffer in the middle of the first iterations (measured at o=40,v=300: +60 ms per iteration while the
        history fills).  Allocate the final footprint once, here, and hand it to the caching allocator: the later
        requests are served from its pool.  Skipped when memory is short (the history then grows on demand)."""
        if like.device.type != "cuda" or self.max_diis <= 0:
            return
        need = (2 * min(self.max_diis, 16) + 1) * like.numel() * like.element_size()
        try:
            if torch.cuda.mem_get_info(like.device)[0] < 2 * need:
                return
            pool = [torch.empty_like(like) for _ in range(2 * min(self.max_diis, 16) + 1)]
            del pool
        except RuntimeError:
            pass

    @staticmethod
    def _clone(buf):
        out = torch.empty_like(buf)
        return K.axpbyz(1.0, buf, 0.0, None, out)

    def _split(self, buf):
        return buf[:self.n1].view(self.shapes[0]), buf[self.n1:].view(self.shapes[1])

    def _row_dots(self, err, others):
        """err . other for every stored error vector, summed over the ranks' row shares (t1 counted once)."""
        if self.comm is None:
            return K.multi_dot(err, others).tolist()
        m = len(others)
        d1 = K.multi_dot(err[:self.n1], [x[:self.n1] for x in others])
        d = torch.zeros(m, dtype=err.dtype, device=err.device)
        if self.n2:
            d = K.multi_dot(err[self.n1:], [x[self.n1:] for x in others])
        self.comm.all_reduce_sum(d)
        return K.axpbyz(1.0, d, 1.0, d1, d).tolist()

    # -- reference interface -----------------------------------------------------------------------
    def add_error_vector(self, t1, t2):
        """Store the new iterate and e = t - t_prev  (utils.py:284-295)."""
        if self.max_diis == 0:
            return
        val = self._flat(t1, t2)
        err = torch.empty_like(val)
        K.axpbyz(1.0, val, -1.0, self.old, err)
        self.vals.append(val)
        self.errors.append(err)
        self._ids.append(self._next_id)
        self._next_id += 1
        self.old = self._clone(val)
        # new row of B: one pass over the newest error vector per 16 stored ones (the kernel's limit; the history can
        # be longer -- max_diis > 16, or start_diis > max_diis, which the reference accepts, utils.py:317-320)
        new = self._ids[-1]
        for c0 in range(0, len(self.errors), 16):
            dots = self._row_dots(err, self.errors[c0:c0 + 16])
            for eid, d in zip(self._ids[c0:c0 + 16], dots):
                self._dots[(eid, new)] = self._dots[(new, eid)] = d

    def extrapolate(self, t1, t2):
        """Pulay extrapolation (utils.py:297-361); returns (t1, t2) unchanged when max_diis == 0.  Sharded: the rows of
        the extrapolant are written into the given ``t2`` and all-gathered (``t2`` is returned, complete on every rank)."""
        if self.max_diis == 0:
            return t1, t2
        if len(self.errors) > self.max_diis:
            dead = self._ids.pop(0)
            del self.vals[0]
            del self.errors[0]
            self._dots = {k: v for k, v in self._dots.items() if dead not in k}
        m = self.diis_size = len(self.errors)
        B = -np.ones((m + 1, m + 1))
        B[-1, -1] = 0.0
        for p in range(m):
            for q in range(m):
                B[p, q] = self._dots[(self._ids[p], self._ids[q])]
        B[:-1, :-1] /= np.abs(B[:-1, :-1]).max()
        rhs = np.zeros(m + 1)
        rhs[-1] = -1.0
        try:
            c = np.linalg.solve(B, rhs)
        except np.linalg.LinAlgError:          # exactly singular B (repeated iterates): minimum-norm solution
            c = np.linalg.lstsq(B, rhs, rcond=None)[0]
        self.last_coefficients = c[:m].copy()
        new = torch.empty_like(self.old)
        K.multi_axpy(c[:min(m, 16)], self.vals[1:min(m, 16) + 1], new)
        for c0 in range(16, m, 16):                      # more than 16 vectors: further chunks are added on
            part = torch.empty_like(new)
            K.multi_axpy(c[c0:min(m, c0 + 16)], self.vals[1 + c0:min(m, c0 + 16) + 1], part)
            K.axpbyz(1.0, new, 1.0, part, new)
        self.old = self._clone(new)
        if self.comm is None:
            return self._split(new)
        n1, n2 = self._split(new)
        i0, i1 = self.rows
        if self.n2:
            K.strided_axpby(t2[i0:i1], n2, 1.0, 0.0)
        self.comm.all_gather_rows(t2)
        return n1, t2


# ---- solver text output: same lines as the reference prints (utils.py:200-254) -------------------------
def title(text):
    return "\n" + text


def iteration(niter, energy=None, de=None, rms=None, e_label="Ecorr", note=None):
    cols = []
    if energy is not None:
        cols.append("%s = %.15f" % (e_label, energy))
    if de is not None:
        cols.append("dE = % .5E" % de)
    if rms is not None:
        cols.append("rms = % .5E" % rms)
    if note is not None:
        cols.append(note)
    return "Iter %3d: %s" % (niter, "  ".join(cols))


def converged(name, elapsed):
    return "%s converged in %.3f seconds." % (name, elapsed)


def timing(label, seconds):
    return "%s built in %.3f seconds." % (label, seconds)


def field(label, value, width=22):
    return "  %-*s : %s" % (width, label, value)


def solve_params(method, e_conv, r_conv, maxiter, max_diis, start_diis):
    rows = ["\n%s solve" % method,
            field("energy convergence", "%.2E" % e_conv),
            field("residual convergence", "%.2E" % r_conv),
            field("max iterations", "%d" % maxiter),
            field("DIIS max / start", "%d / %d" % (max_diis, start_diis))]
    return "\n".join(rows)


# ---- complex arguments through real evaluations (RT-CC right-hand sides) -------------------------------------------
# sample points s and the weights that turn {R(s)} into Re R(i), Im R(i) for a polynomial R of degree <= 4 in s
CPLX_PTS = (-2.0, -1.0, 0.0, 1.0, 2.0)
CPLX_RE = (1.0 / 12.0, -5.0 / 6.0, 2.5, -5.0 / 6.0, 1.0 / 12.0)
CPLX_IM = (1.0 / 6.0, -5.0 / 6.0, 0.0, 5.0 / 6.0, -1.0 / 6.0)
# degree <= 5 (the CC3 triples terms): six half-integer points; the weights solve sum_k w_k s_k^m = Re / Im (i^m), m <= 5
CPLX5_PTS = (-2.5, -1.5, -0.5, 0.5, 1.5, 2.5)
CPLX5_RE = (65.0 / 768.0, -145.0 / 256.0, 377.0 / 384.0, 377.0 / 384.0, -145.0 / 256.0, 65.0 / 768.0)
CPLX5_IM = (-13.0 / 384.0, 145.0 / 384.0, -377.0 / 192.0, 377.0 / 192.0, -145.0 / 384.0, 13.0 / 384.0)
CPLX_RULES = {4: (CPLX_PTS, CPLX_RE, CPLX_IM), 5: (CPLX5_PTS, CPLX5_RE, CPLX5_IM)}


def is_complex(x):
    return x.is_complex() if isinstance(x, torch.Tensor) else np.iscomplexobj(x)


def real_planes(x, device):
    """(re, im) contiguous float64 device copies of ``x`` (tensor or array); im is None for real input."""
    from . import kernels as K
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    x = x.to(device)
    ident = tuple(range(x.dim()))
    if x.is_complex():
        xr = torch.view_as_real(x.to(torch.complex128))
        return K.permuted(xr[..., 0], ident), K.permuted(xr[..., 1], ident)
    return K.permuted(x.to(torch.float64), ident), None


def planes_to_complex(re, im):
    """complex128 tensor from two equally shaped float64 planes (our strided copy kernel, no torch arithmetic)."""
    from . import kernels as K
    z = torch.empty(tuple(re.shape), dtype=torch.complex128, device=re.device)
    zr = torch.view_as_real(z)
    K.strided_axpby(zr[..., 0], re, 1.0, 0.0)
    K.strided_axpby(zr[..., 1], im, 1.0, 0.0)
    return z


def complex_from_real_samples(fn, args, device, degree=4, as_planes=False, planes=None):
    """Evaluate ``fn(*args) -> tuple of real tensors`` for COMPLEX ``args`` when fn is polynomial of total degree <= 4
    in its arguments (scaled together), using only real evaluations (``degree=5``: six samples, see CPLX5_*).

    For x = x_re + s x_im the map s -> fn(x(s)) is then a real polynomial of degree <= 4 in the real parameter s, and
    its value at s = i follows EXACTLY from the five samples s = -2..2:
        Re R(i) = 5/2 R(0) - 5/6 [R(1) + R(-1)] + 1/12 [R(2) + R(-2)]
        Im R(i) = 5/6 [R(1) - R(-1)] - 1/6 [R(2) - R(-2)]
    (weights sum to 4.3 in magnitude: less than one digit of round-off amplification).  The CCSD T residual
    (quartic in t1; every Fock term carries at most two amplitudes) and the Lambda residual with HBAR rebuilt from
    (F, t1, t2) (checked numerically with the oracle: exact at degree 4) both qualify.  Each sample runs the fused
    FP64 kernels of the energy path, so no complex kernel exists in the library.  Returns complex128 tensors, or with
    ``as_planes`` a list of (re, im) float64 pairs; ``planes``: the (re, im) pairs of ``args`` if the caller has them."""
    from . import kernels as K

    def at(re, im, s):
        if im is None or s == 0.0:
            return re
        return K.axpbyz(1.0, re, s, im, torch.empty_like(re))

    pts, w_re, w_im = CPLX_RULES[int(degree)]
    P = [real_planes(a, device) for a in args] if planes is None else planes
    samples = []
    for s in pts:
        out = fn(*[at(re, im, s) for re, im in P])
        samples.append([o.contiguous() for o in out])
    result = []
    for q in range(len(samples[0])):
        shape = tuple(samples[0][q].shape)
        flat = [smp[q].reshape(-1) for smp in samples]
        re = torch.empty(shape, dtype=torch.float64, device=device)
        im = torch.empty(shape, dtype=torch.float64, device=device)
        K.multi_axpy(w_re, flat, re.view(-1))
        K.multi_axpy([w for w in w_im if w != 0.0], [r for r, w in zip(flat, w_im) if w != 0.0], im.view(-1))
        for smp in samples:
            smp[q] = None                                            # release the samples of this output
        del flat
        result.append((re, im) if as_planes else planes_to_complex(re, im))
    return result
