"""2-rank diagnostic: where does the slab evaluation deviate from the undecomposed one?"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
import odil
from tests import operators as ops
def rel(a, b): return float((a - b).abs().max() / b.abs().max())
cshape, nlvl, dt = (32 * world, 16, 24), 3, np.float64
problem, state = ops.make_poisson(cshape, nlvl, dt)
domain = problem.domain
rng = np.random.default_rng(7)
terms = [rng.standard_normal(tuple(cs)).astype(dt) for cs in domain.mg_cshapes]
arrays = [domain.slab.scatter(domain.mod.variable(t, dtype=dt)) for t in terms]
domain.arrays_to_state(arrays, state)
p1, s1 = ops.make_poisson(cshape, nlvl, dt)
d1 = p1.domain; d1.slab = None
st = odil.State(); st.fields["u"] = np.zeros(cshape, dtype=dt); s1 = d1.init_state(st)
d1.arrays_to_state([d1.mod.variable(t, dtype=dt) for t in terms], s1)
U1 = d1.field(s1, "u").full()
U = domain.field(state, "u").full()
if rank == 0: print("U relerr", rel(U, U1), flush=True)
eng = problem._engine(state)
loss1 = float(p1.eval_loss_grad(s1)[0])
plan = eng.outputs[0].blocks[0].plan
for variant in [0, 20, 12]:
    plan.tune(0, variant)
    loss, grads, *_ = problem.eval_loss_grad(state)
    g = domain.slab.gather(grads[0]); g1 = p1.eval_loss_grad(s1)[1][0]
    gs = [rel(domain.slab.gather(a), b) for a, b in zip(grads, p1.eval_loss_grad(s1)[1])]
    if rank == 0: print("variant", variant, "loss", float(loss), loss1, "grad relerr per level", gs, flush=True)
    # per-plane error of the finest gradient
    e = (g - g1).abs().amax(dim=(1, 2)) / g1.abs().max()
    if rank == 0: print("  planes with error > 1e-9:", torch.nonzero(e > 1e-9).reshape(-1).tolist(), flush=True)
dist.barrier(); dist.destroy_process_group()
