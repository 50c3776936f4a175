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

Data plane on GPUs: the library's own communicator (csrc/comm.cu, `native.Comm`): boundary planes are stored
straight into the neighbours' staging areas over NVLink peer memory and flagged -- two small kernels per exchange,
one per all-reduce, all on the compute stream and capturable in a CUDA graph.  torch.distributed (NCCL) only
carries the 64-byte IPC handles at start-up.  ODIL_B200_COMM=nccl selects the round-1 arrangement (batched
ncclSend/ncclRecv groups + ncclAllReduce through torch.distributed); CPU tensors (the gloo tests of the host
logic) always take the torch.distributed route.
"""
import os

import torch

HALO = 2
_COMM = {}  # one peer-memory communicator per process: (rank, world) -> native.Comm


class SlabInfo:

    def __init__(self, rank, world, group=None, halo=HALO):
        self.rank, self.world, self.group = int(rank), int(world), group
        self.halo = int(halo)
        self.use_peer = os.environ.get("ODIL_B200_COMM", "peer") != "nccl"

    @property
    def comm(self):
        return _COMM.get((self.rank, self.world))

    # -- peer-memory communicator -------------------------------------------------------------------
    def _all_gather_bytes(self, raw):
        import torch.distributed as dist

        mine = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return [bytes(p.cpu().tolist()) for p in parts]

    def ensure_comm(self, nbytes):
        """Communicator whose staging holds `nbytes` per direction (collective when it has to be (re)created: every
        rank asks for the same size because every rank exchanges the same arrays)."""
        from . import native

        comm = self.comm
        if comm is None or comm.capacity < nbytes:
            if comm is not None:
                import torch.distributed as dist

                torch.cuda.synchronize()
                dist.barrier(group=self.group)  # nobody may still be writing into the block that is about to go
                comm.destroy()
            comm = native.Comm(self.rank, self.world, max(int(nbytes * 1.25), 1 << 20), self._all_gather_bytes)
            _COMM[(self.rank, self.world)] = comm
        return comm

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
        if not locals_:
            return
        if self.world == 1:
            for a in locals_:
                n = a.shape[0] - 2 * self.halo
                H = self.halo
                a[H - width:H].copy_(a[H + n - width:H + n])
                a[H + n:H + n + width].copy_(a[H:H + width])
            return
        if self.use_peer and locals_ and all(a.is_cuda for a in locals_):
            from . import native

            send_lo, send_hi, recv_lo, recv_hi = [], [], [], []
            for a in locals_:
                H = self.halo
                n = a.shape[0] - 2 * H
                send_lo.append(a[H:H + width])
                send_hi.append(a[H + n - width:H + n])
                recv_lo.append(a[H - width:H])
                recv_hi.append(a[H + n:H + n + width])
            need = native.Comm.bytes_needed([t.numel() * t.element_size() for t in send_lo])
            self.ensure_comm(need).halo_exchange(send_lo, send_hi, recv_lo, recv_hi)
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

    def accumulate(self, items):
        """
        Transpose of `exchange`: items = [(send_lo, send_hi, acc_lo, acc_hi)] of contiguous plane views.  send_lo /
        send_hi hold the partial sums this rank computed for the plane(s) just below / above its slab; the owners add
        them: acc_lo (this rank's first owned planes) += the lower neighbour's send_hi, acc_hi (last owned planes) +=
        the upper neighbour's send_lo -- in that order.  One call per epoch for all multigrid levels.
        """
        import torch.distributed as dist

        if not items:
            return
        if self.world == 1:
            for send_lo, send_hi, acc_lo, acc_hi in items:
                lo, hi = send_hi.clone(), send_lo.clone()  # ring of one: my own partials wrap around
                acc_lo.add_(lo)
                acc_hi.add_(hi)
            return
        if self.use_peer and all(t.is_cuda for it in items for t in it):
            from . import native

            cols = list(zip(*items))
            need = native.Comm.bytes_needed([t.numel() * t.element_size() for t in cols[0]])
            self.ensure_comm(need).halo_accumulate(list(cols[0]), list(cols[1]), list(cols[2]), list(cols[3]))
            return
        lo = (self.rank - 1) % self.world
        hi = (self.rank + 1) % self.world
        ops, bufs = [], []
        for send_lo, send_hi, acc_lo, acc_hi in items:
            from_lo, from_hi = torch.empty_like(acc_lo), torch.empty_like(acc_hi)
            bufs.append((from_lo, from_hi, acc_lo, acc_hi))
            ops.append(dist.P2POp(dist.isend, send_hi.contiguous(), hi, group=self.group))
            ops.append(dist.P2POp(dist.irecv, from_lo, lo, group=self.group))
            ops.append(dist.P2POp(dist.isend, send_lo.contiguous(), lo, group=self.group))
            ops.append(dist.P2POp(dist.irecv, from_hi, hi, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for from_lo, from_hi, acc_lo, acc_hi in bufs:
            acc_lo.add_(from_lo)
            acc_hi.add_(from_hi)

    def all_reduce_sum(self, t):
        import torch.distributed as dist

        if self.world > 1:
            if self.use_peer and t.is_cuda and t.dtype == torch.float64 and t.numel() <= 16:
                self.ensure_comm(0).allreduce_scalars(t)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t
