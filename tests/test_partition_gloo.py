"""N>1 host logic on CPU: two gloo ranks derive their node-block partition plans from the replicated
mesh (fs_partition_plan, the code fs_set_mesh runs), exchange halo values exactly as the CUDA path
does (pack by send_idx -> send/recv into the contiguous halo segments) and reproduce the serial
oracle SpMV and a distributed dot product (allreduce) on their owned rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fem_shell_b200 as fsb
import meshes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import fso
        if case == "tri":
            m = fsb.meshgen("t", 11, 9, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 1)
        elif case == "quad":
            m = fsb.meshgen("q", 7, 14, 0, 0, 10, 10, (0, 1, 0, 1), 300.0, 2, 1)
        else:
            m = meshes.folded_cantilever(nx=9, ny=6, skew=0.2)
        n_nodes = m["xyz"].shape[0]
        om = fso.Mesh(np.asarray(m["xyz"], float), m["etype"], m["eptr"], m["enodes"], m["bc"])
        ref = fso.assemble(om, m["forces"], 0.3, 1e7, 0.5)          # serial, global numbering
        A = ref.scipy()
        plan = fsb.partition_plan(m["eptr"], m["enodes"], n_nodes, rank, world)
        l2g = plan["local_to_global"]
        ob, oe, olo = plan["own_begin"], plan["own_end"], plan["own_lo"]
        n_own = oe - ob
        assert np.all(np.diff(l2g) > 0) and np.array_equal(l2g[olo:olo + n_own], np.arange(ob, oe))
        # local elements: exactly those touching an owned node
        dn = ref.dofnode
        touch = [e for e in range(om.n_elem) if any(ob <= dn[n] < oe for n in m["enodes"][m["eptr"][e]:m["eptr"][e + 1]])]
        assert list(plan["loc_elems"]) == touch
        # halo completeness: every column of an owned row is a local node
        g2l = -np.ones(ref.n_dofnodes, np.int64)
        g2l[l2g] = np.arange(l2g.size)
        rows = A[6 * ob:6 * oe]
        assert np.all(g2l[np.unique(rows.indices // 6)] >= 0)
        # halo exchange as the device path does it
        xg = np.random.default_rng(5).standard_normal(6 * ref.n_dofnodes)
        xl = np.zeros((l2g.size, 6))
        xl[olo:olo + n_own] = xg.reshape(-1, 6)[ob:oe]
        ops, bufs = [], []
        for pr in plan["peers"]:
            if pr["send_count"]:
                idx = plan["send_idx"][pr["send_off"]:pr["send_off"] + pr["send_count"]]
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(xl[idx].copy()), pr["rank"]))
            if pr["recv_count"]:
                t = torch.empty((pr["recv_count"], 6), dtype=torch.float64)
                bufs.append((pr, t))
                ops.append(dist.P2POp(dist.irecv, t, pr["rank"]))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for pr, t in bufs:
            xl[pr["recv_off"]:pr["recv_off"] + pr["recv_count"]] = t.numpy()
        assert np.array_equal(xl, xg.reshape(-1, 6)[l2g]), "halo values differ from the global vector"
        # distributed SpMV on owned rows + allreduced dot
        y_own = rows @ xg
        y_ser = fso.spmv(ref, xg)[6 * ob:6 * oe]
        assert np.abs(y_own - y_ser).max() <= 1e-12 * np.abs(y_ser).max()
        d = torch.tensor([float(xg[6 * ob:6 * oe] @ y_own)], dtype=torch.float64)
        dist.all_reduce(d)
        assert abs(d.item() - xg @ fso.spmv(ref, xg)) <= 1e-10 * abs(d.item())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["tri", "quad", "mixed"])
def test_partition_and_halo_exchange(case, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_plans_are_mutually_consistent():
    m = fsb.meshgen("t", 13, 10, 0, 0, 10, 10, (1, 1, 1, 1), 300.0, 2, 0)
    n_nodes = m["xyz"].shape[0]
    W = 4
    plans = [fsb.partition_plan(m["eptr"], m["enodes"], n_nodes, r, W) for r in range(W)]
    assert plans[0]["own_begin"] == 0 and plans[-1]["own_end"] == plans[0]["n_global"]
    for a in range(W):
        assert a == 0 or plans[a]["own_begin"] == plans[a - 1]["own_end"]
        for pr in plans[a]["peers"]:
            b = pr["rank"]
            back = [p for p in plans[b]["peers"] if p["rank"] == a][0]
            sent = plans[a]["local_to_global"][plans[a]["send_idx"][pr["send_off"]:pr["send_off"] + pr["send_count"]]]
            recv = plans[b]["local_to_global"][back["recv_off"]:back["recv_off"] + back["recv_count"]]
            assert np.array_equal(sent, recv), (a, b)
