"""
Slab decomposition of the grid along axis 0 over the GPUs of one node (one process per GPU,
torch.distributed; NCCL over NVLink on GPUs, gloo in the CPU tests).

The reference has no multi-device path at all; this module adds the only decomposition
that makes sense for the workload: every array of every multigrid level is split into `world`
contiguous slabs of planes.  A local array carries HALO = 2 extra planes on both sides:

    local[0:H]          halo planes owned by the lower neighbour (rank-1, periodic ring)
    local[H:H+n]        the n = N0/world planes owned by this rank
    local[H+n:H+n+H]    halo planes owned by the upper neighbour

Axis-0 planes are contiguous in C order, so halos are sent/received in place (no pack kernels).
Per epoch the engine does ONE batched exchange of all multigrid terms (width 2), synthesises every
level on its extended range, runs the fused stencil kernel on its owned planes, and exchanges the
gradient halo (width 1) once per multigrid level for the transposed interpolation.  Loss terms are
summed with one tiny all-reduce.
"""
import os

import torch

HALO = 2


class SlabInfo:

    def __init__(self, rank, world, group=None, halo=HALO):
        self.rank, self.world, self.group = int(rank), int(world), group
        self.halo = int(halo)

    @staticmethod
    def from_environment():
        """Active when torch.distributed is initialised with more than one rank (and ODIL_SLABS != 0)."""
        import torch.distributed as dist

        if int(os.environ.get("ODIL_SLABS", "1")) == 0:
            return None
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        # ODIL_HALO: halo planes per side; must be >= 2 * (largest |shift| along axis 0), default 2.
        return SlabInfo(dist.get_rank(), dist.get_world_size(), halo=int(os.environ.get("ODIL_HALO", HALO)))

    # -- geometry -----------------------------------------------------------------------------------
    def check(self, shape):
        if shape[0] % self.world != 0:
            raise ValueError(f"axis 0 of size {shape[0]} is not divisible by {self.world} slabs")
        n = shape[0] // self.world
        if n < self.halo:
            raise ValueError(f"slab of {n} planes is thinner than the halo ({self.halo})")
        return n

    def owned_range(self, shape):
        n = self.check(shape)
        return self.rank * n, n

    def local_shape(self, shape):
        return (self.check(shape) + 2 * self.halo,) + tuple(shape[1:])

    def owned(self, local):
        """View of the owned planes of a local array (contiguous)."""
        return local[self.halo: local.shape[0] - self.halo]

    # -- global <-> local ---------------------------------------------------------------------------
    def scatter(self, global_tensor):
        """Local array (with periodic halos filled) cut from a tensor every rank holds in full."""
        z0, n = self.owned_range(global_tensor.shape)
        N, H = global_tensor.shape[0], self.halo
        idx = torch.arange(z0 - H, z0 + n + H, device=global_tensor.device) % N
        return global_tensor.index_select(0, idx).contiguous()

    def gather(self, local):
        """Global tensor assembled from the owned planes of every rank (collective)."""
        own = self.owned(local).contiguous()
        if self.world == 1:
            return own.clone()
        import torch.distributed as dist

        parts = [torch.empty_like(own) for _ in range(self.world)]
        dist.all_gather(parts, own, group=self.group)
        return torch.cat(parts, dim=0)

    # -- halo exchange ------------------------------------------------------------------------------
    def exchange(self, locals_, width=None):
        """
        Fills the `width` innermost halo planes of every local array in `locals_` from the ring
        neighbours (rank 0's lower neighbour is rank world-1: periodic, which is what ctx.field's roll
        means; non-periodic problems multiply those planes by zero coefficients).  One batched
        send/recv group for all arrays.
        """
        import torch.distributed as dist

        width = self.halo if width is None else width
        if self.world == 1:
            for a in locals_:
                n = a.shape[0] - 2 * self.halo
                H = self.halo
                a[H - width:H].copy_(a[H + n - width:H + n])
                a[H + n:H + n + width].copy_(a[H:H + width])
            return
        lo = (self.rank - 1) % self.world
        hi = (self.rank + 1) % self.world
        ops = []
        for a in locals_:
            H = self.halo
            n = a.shape[0] - 2 * H
            send_lo = a[H:H + width]              # my first owned planes -> lower neighbour's upper halo
            send_hi = a[H + n - width:H + n]      # my last owned planes  -> upper neighbour's lower halo
            recv_lo = a[H - width:H]
            recv_hi = a[H + n:H + n + width]
            ops.append(dist.P2POp(dist.isend, send_hi, hi, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv_lo, lo, group=self.group))
            ops.append(dist.P2POp(dist.isend, send_lo, lo, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv_hi, hi, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def all_reduce_sum(self, t):
        import torch.distributed as dist

        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t
