"""Device / precision policy and the contraction backend -- the drop-in seam of the reference
(pycc/device.py:38-173), re-targeted: ``ContractionBackend.__call__(subscripts, *operands)`` keeps
its signature but runs on ``b200cc_dgemm`` / ``b200cc_permute`` instead of opt_einsum + cuBLAS, and
there is no CPU device: this package is the ``device='GPU'`` implementation only.
"""
from __future__ import annotations

import numpy as np
import torch

from .contract import Contractor
from .exceptions import InvalidKeywordError, PyCCError


def _current_device():
    """The reference pins 'cuda:0' (device.py:62,142); with one process per GPU the compute device is the
    process's CURRENT CUDA device (torch.cuda.set_device(LOCAL_RANK)), which is cuda:0 in the single-GPU case."""
    if torch.cuda.is_available():
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


class ContractionBackend(object):
    """``contract(subscripts, *operands)`` on the B200 kernels (reference: device.py:64-86).

    Operands that are not yet resident on the compute device (numpy arrays, CPU tensors -- the
    reference keeps ERI/L on the host) are uploaded first, exactly as the reference does per call;
    resident float64 CUDA tensors (views included) are used in place.
    """

    def __init__(self, device='GPU', device1=None):
        if device != 'GPU':
            raise PyCCError("pycc_b200 only implements device='GPU' (use pycc itself for the CPU path)")
        self.device = device
        self.device1 = device1 if device1 is not None else _current_device()
        self.engine = Contractor()

    def _resident(self, x):
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x))
        want = torch.complex128 if x.is_complex() else torch.float64
        if x.device != self.device1 or x.dtype != want:
            x = x.to(self.device1, dtype=want)
        return x

    def __call__(self, subscripts, *operands, **kw):
        ops = [self._resident(x) for x in operands]
        if any(x.is_complex() for x in ops):
            return self._complex(subscripts, ops, **kw)
        return self.engine(subscripts, *ops, **kw)

    # ---- complex operands (RT-CC: the reference upcasts real operands to complex and lets torch.einsum do complex
    # arithmetic, device.py:79-83).  Here a complex contraction is real GEMMs on the real / imaginary parts, which are
    # strided float64 views of the complex storage: 3 products for complex x complex (Karatsuba / "3M":
    # (Ar + Ai)(Br + Bi) - ArBr - AiBi is the imaginary part), 2 for complex x real.  Result: complex128.
    def _complex(self, subscripts, ops, out=None, alpha=1.0, beta=0.0):
        from . import kernels as K
        if out is not None or beta != 0.0:
            raise NotImplementedError("complex contractions return a new tensor (no in-place / accumulate form)")
        ins, res = subscripts.replace(" ", "").split("->")
        ins = ins.split(",")
        if len(ops) != len(ins):
            raise PyCCError("contract: %d operands for %r" % (len(ops), subscripts))

        def parts(x):
            return (x.real, x.imag) if x.is_complex() else (x, None)

        def cplx(re, im):
            z = torch.empty(tuple(re.shape), dtype=torch.complex128, device=re.device)
            K.strided_axpby(z.real, re, 1.0, 0.0)
            K.strided_axpby(z.imag, im, 1.0, 0.0)
            return z

        def pair(sub, A, B, scale=1.0):
            (ar, ai), (br, bi) = parts(A), parts(B)
            if ai is None:                                             # real x complex
                return cplx(self.engine(sub, ar, br, alpha=scale), self.engine(sub, ar, bi, alpha=scale))
            if bi is None:                                             # complex x real
                return cplx(self.engine(sub, ar, br, alpha=scale), self.engine(sub, ai, br, alpha=scale))
            p1 = self.engine(sub, ar, br, alpha=scale)
            p2 = self.engine(sub, ai, bi, alpha=scale)
            sa = K.strided_axpby(K.permuted(ar, tuple(range(ar.dim()))), ai, 1.0, 1.0)       # Ar + Ai
            sb = K.strided_axpby(K.permuted(br, tuple(range(br.dim()))), bi, 1.0, 1.0)       # Br + Bi
            im = self.engine(sub, sa, sb, alpha=scale)
            K.strided_axpby(im, p1, -1.0, 1.0)
            K.strided_axpby(im, p2, -1.0, 1.0)
            K.strided_axpby(p1, p2, -1.0, 1.0)                                               # ArBr - AiBi
            return cplx(p1, im)

        if len(ops) == 1:
            (xr, xi) = parts(ops[0])
            sub = "%s->%s" % (ins[0], res)
            return cplx(self.engine(sub, xr, alpha=alpha), self.engine(sub, xi, alpha=alpha))
        cur, cur_idx = ops[0], ins[0]
        for n in range(1, len(ops)):                    # left to right; intermediates keep every index still needed
            last = n == len(ops) - 1
            later = "".join(ins[n + 1:]) + res
            tgt = res if last else "".join(ch for ch in dict.fromkeys(cur_idx + ins[n]) if ch in later)
            sub = "%s,%s->%s" % (cur_idx, ins[n], tgt)
            scale = alpha if last else 1.0
            if cur.is_complex() or ops[n].is_complex():
                cur = pair(sub, cur, ops[n], scale)
            else:
                cur = self.engine(sub, cur, ops[n], alpha=scale)
            cur_idx = tgt
        return cur


class DeviceManager(object):
    """Validates ``device`` / ``precision`` and owns the contraction backend (reference: device.py:89-173).

    Differences, all deliberate: 'CPU' is rejected (there is no CPU path here), and both the "storage"
    and the "compute" handle are the GPU -- integrals live in HBM as blocks, nothing is staged on the host.
    """

    VALID_DEVICE = ['CPU', 'GPU']
    # 'DP': FP64 DMMA everywhere.  'MP' (new): every tensor stays FP64, the large contractions run as split-TF32
    # products on the tcgen05 tensor cores with FP64 accumulation (BASELINE configs[4]).  'SP': the reference computes
    # the whole calculation in float32 (device.py:147-151); here it is served by the 'MP' path, whose error is below
    # float32's (documented difference: tensors are returned as float64).
    VALID_PRECISION = ['SP', 'DP', 'MP']

    def __init__(self, device='GPU', precision='DP'):
        if precision.upper() not in self.VALID_PRECISION:
            raise InvalidKeywordError('precision', precision, self.VALID_PRECISION)
        self.precision = precision.upper()
        if device.upper() not in self.VALID_DEVICE:
            raise InvalidKeywordError('device', device, self.VALID_DEVICE)
        self.device = device.upper()
        if self.device != 'GPU':
            raise PyCCError("pycc_b200 only implements device='GPU' (use pycc itself for the CPU path)")
        self.mixed = self.precision in ('SP', 'MP')
        self.device1 = _current_device()
        self.device0 = self.device1
        self.real_dtype = np.float64
        self._torch_dtype = torch.float64
        self.contract = ContractionBackend(device=self.device, device1=self.device1)

    def seed_compute(self, a):
        return torch.as_tensor(np.asarray(a), dtype=self._torch_dtype).to(self.device1)

    seed_store = seed_compute
