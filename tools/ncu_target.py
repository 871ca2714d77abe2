"""what `ncu --set full --profile-from-start off` should see of the bench workload (c2 by default): one values pass
and one captured batch of 8 Jacobi-PCG iterations, all of them live (no kernel returns early on the done flag)
    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/TAG python tools/ncu_target.py [nodes] [q|t]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fem_shell_b200 as fsb

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kind = sys.argv[2] if len(sys.argv) > 2 else "q"
# several ranks (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* in the environment, one strip per rank): ONE of them may run
# under ncu (tools/ncu_nvlink.sh) -- the rendezvous goes through a gloo group, the exchange through the library's own paths
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
nid = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    ids = [fsb.FemShell.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nid = ids[0]
m = fsb.meshgen(kind, nodes - 1, nodes * world - 1, 0.0, 0.0, 10.0, 10.0 * world, (1, 1, 1, 1), 300.0, 2, 1)
s = fsb.FemShell(device=lr, rank=rank, world=world, nccl_id=nid)
s.set_material(0.3, 1.0e7, 0.5)
s.set_mesh(m["xyz"], m["etype"], m["eptr"], m["enodes"], m["bc"])
s.set_nodal_loads(m["forces"])
for _ in range(3):
    s.assemble()
s.build_rhs(1.0)
kw = dict(rtol=1e-30, pc=fsb.PC_JACOBI, warm_start=False, allow_not_converged=True)
s.solve(max_its=64, check_every=64, **kw)      # capture + warm caches
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
s.assemble()
info = s.solve(max_its=8, check_every=8, **kw)
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("rank %d of %d profiled: 1 values pass (%s), %d iterations, comm %s" % (rank, world, s.assembly_path(), info.iterations, "peer" if world > 1 and s.comm_mode() == fsb.COMM_PEER else "-"))
if world > 1:
    dist.barrier()
s.close()
