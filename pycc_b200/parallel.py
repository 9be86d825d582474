"""One process per GPU: how the CCSD residual and (T) are split across ranks.

The reference has no distributed code at all.  Here ``torch.distributed`` (NCCL over NVLink/NVSwitch on
B200; gloo in the CPU tests) is plumbing only -- ONE all-reduce of the half-residual r2 (o^2 v^2
doubles) per CCSD iteration and ONE scalar all-reduce for E(T):

* <ab|ef> is held a-sharded in its pair-packed form (rows (a >= b), see csrc/pairs.cu): rank g owns the pairs of
  the rows a in ``a_range(nv)`` -- boundaries chosen for equal PAIR counts -- and computes the ladder contribution to
  r2[:, :, a_g, b <= a] and its (b, a) image from the (replicated) tau;
* every other o^3v^3 / o^4v^2 / o^2v^3 term of r2 is split over an occupied index -- rank g computes
  r2[i_g] rows (F/W_mnij/Z/t1-driven terms) and r2[:, j_g] columns (ring terms, whose W_mbej/W_mbje
  intermediates are built only for the local j_g, so they are never gathered);
* after the all-reduce every rank holds the same half residual; the HBM-bound tail of the iteration is then SHARDED
  over the rows i_g of t2: symmetrise + Jacobi update + rms, the energy, and DIIS (history, B-matrix dots and the
  extrapolation all on the rank's rows; the dots and the two scalars are summed with tiny all-reduces), and ONE
  all-gather of the t2 rows (o^2v^2 doubles in total, 1/N per rank) ends the iteration.  t1 is replicated.  The kernels
  are deterministic and every rank solves the same small DIIS system, so the gathered amplitudes are identical on all
  ranks (``CCwfn.shard_update = False`` restores the fully replicated tail);
* (T): the (i>=j>=k) triples are dealt round-robin (equal cost), energies summed with a scalar all-reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def split(n, size, rank):
    """Balanced contiguous partition of range(n): (lo, hi) of part ``rank``."""
    base, rem = divmod(n, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def split_pairs(n, size, rank):
    """Contiguous partition of range(n) into ``size`` parts with (nearly) equal numbers of PAIRS (a >= b): part
    boundaries at a_k ~ n sqrt(k / size).  <ab|ef> is held as its pair-packed rows (a >= b), so equal work per rank
    means equal pair counts, not equal row counts."""
    total = n * (n + 1) // 2

    def bound(k):
        if k <= 0:
            return 0
        if k >= size:
            return n
        target = total * k / size
        a = int((2.0 * target) ** 0.5)
        best = min(range(max(0, a - 2), min(n, a + 3) + 1), key=lambda x: abs(x * (x + 1) // 2 - target))
        return best
    return bound(rank), bound(rank + 1)


class Comm:
    """Rank / size and the two collectives the path needs."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def a_range(self, nv):
        return split_pairs(nv, self.size, self.rank)

    def occ_range(self, no):
        return split(no, self.size, self.rank)

    def occ_range_of(self, no, rank):
        return split(no, self.size, rank)

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_rows(self, t):
        """``t``: a contiguous tensor whose leading (occupied) dimension is split over the ranks by ``occ_range``; every
        rank holds valid data in its own rows.  Afterwards every rank holds all rows.  NCCL with equal shares: one
        all-gather; otherwise (gloo in the CPU tests, ragged shares) one broadcast per rank."""
        n = t.shape[0]
        bounds = [split(n, self.size, r) for r in range(self.size)]
        lo, hi = bounds[self.rank]
        if dist.get_backend(self.group) == "nccl" and len({b - a for a, b in bounds}) == 1 and hi > lo:
            mine = t[lo:hi].clone()
            dist.all_gather_into_tensor(t.view(-1), mine.view(-1), group=self.group)
            return t
        for r, (a, b) in enumerate(bounds):
            if b > a:
                dist.broadcast(t[a:b], src=dist.get_global_rank(self.group, r) if self.group is not None else r,
                               group=self.group)
        return t

    def exchange(self, send, recv):
        """All-to-all with preallocated buffers: ``send[d]`` goes to rank d, ``recv[r]`` receives rank r's piece."""
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all(recv, send, group=self.group)
            return recv
        for r in range(self.size):
            src = dist.get_global_rank(self.group, r) if self.group is not None else r
            for d in range(self.size):
                if r == self.rank:
                    buf = send[d]
                elif d == self.rank:
                    buf = recv[r]
                else:
                    continue
                if r == self.rank and d == self.rank:
                    recv[r].copy_(send[d])
                    continue
                if buf.numel() == 0:
                    continue
                # point-to-point as a two-member exchange: gloo supports send / recv
                if r == self.rank:
                    dist.send(buf, dst=dist.get_global_rank(self.group, d) if self.group is not None else d,
                              group=self.group)
                else:
                    dist.recv(buf, src=src, group=self.group)
        return recv

    def all_reduce_max_scalar(self, x):
        t = torch.tensor([float(x)], dtype=torch.float64)
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t[0])

    def barrier(self):
        dist.barrier(group=self.group)


class Serial:
    """The single-GPU stand-in: full ranges, no communication."""
    rank, size = 0, 1

    def a_range(self, nv):
        return 0, nv

    def occ_range(self, no):
        return 0, no

    def occ_range_of(self, no, rank):
        return 0, no

    def all_reduce_sum(self, t):
        return t

    def all_gather_rows(self, t):
        return t

    def barrier(self):
        pass
