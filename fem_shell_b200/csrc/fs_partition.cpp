// fs_partition.cpp -- see fs_partition.hpp
#include "fs_partition.hpp"
#include "fs_host_par.hpp"

#include <algorithm>
#include <cstring>

#include "../../include/femshell_b200.h"

namespace fs {

int64_t compute_dof_order(int mode, int64_t n_nodes, int64_t n_elem, const int64_t *eptr, const int32_t *enodes,
                          std::vector<int32_t> &dofnode)
{
    dofnode.assign(n_nodes, -1);
    int64_t next = 0;
    const int64_t n_en = eptr[n_elem];
    if (mode == FS_DOF_FIRST_ENCOUNTER) {
        for (int64_t k = 0; k < n_en; k++)
            if (dofnode[enodes[k]] < 0) dofnode[enodes[k]] = (int32_t)next++;
    } else {
        for (int64_t k = 0; k < n_en; k++) dofnode[enodes[k]] = 0;
        for (int64_t i = 0; i < n_nodes; i++)
            if (dofnode[i] == 0) dofnode[i] = (int32_t)next++;
    }
    return next;
}

int owner_of(int64_t g, int64_t n_g, int world)
{
    int r = (int)((g * world) / n_g);
    while (r > 0 && g < (int64_t)r * n_g / world) r--;
    while (r < world - 1 && g >= (int64_t)(r + 1) * n_g / world) r++;
    return r;
}

int plan_partition(const std::vector<int32_t> &dofnode, int64_t n_g, int64_t n_elem, const int64_t *eptr,
                   const int32_t *enodes, int R, int W, PartitionPlan &p, int threads)
{
    if (W < 1 || R < 0 || R >= W || n_g < W) return FS_ERR_ARG;
    p = PartitionPlan();
    p.n_global = n_g;
    p.own_begin = (int64_t)R * n_g / W;
    p.own_end = (int64_t)(R + 1) * n_g / W;
    if (W == 1) {  // everything is local: no scan of the mesh
        p.loc_elems.resize(n_elem);
        p.local_to_global.resize(n_g);
        parallel_chunks(n_elem, threads, [&](int, int64_t e0, int64_t e1) { for (int64_t e = e0; e < e1; e++) p.loc_elems[e] = (int32_t)e; });
        parallel_chunks(n_g, threads, [&](int, int64_t g0, int64_t g1) { for (int64_t g = g0; g < g1; g++) p.local_to_global[g] = (int32_t)g; });
        p.own_lo = 0;
        return FS_OK;
    }

    // local elements = elements touching an owned node; local nodes = their nodes + owned nodes.  The scan over the
    // replicated mesh runs on a few threads, each with its own lists; concatenated in element order afterwards, so
    // the plan does not depend on the number of threads.
    std::vector<uint8_t> is_local(n_g, 0);
    for (int64_t g = p.own_begin; g < p.own_end; g++) is_local[g] = 1;
    const int T_max = std::max(1, threads);
    std::vector<std::vector<int32_t>> part_elems(T_max);
    std::vector<std::vector<std::vector<int32_t>>> part_send(T_max, std::vector<std::vector<int32_t>>(W));
    const int64_t ob = p.own_begin, oe = p.own_end;
    parallel_chunks(n_elem, threads, [&](int t, int64_t e0, int64_t e1) {
        std::vector<int32_t> &mine_elems = part_elems[t];
        std::vector<std::vector<int32_t>> &send = part_send[t];
        for (int64_t e = e0; e < e1; e++) {
            const int nen = (int)(eptr[e + 1] - eptr[e]);
            const int32_t *en = enodes + eptr[e];
            bool mine = false, all_mine = true;
            for (int k = 0; k < nen; k++) {
                const int64_t g = dofnode[en[k]];
                const bool own = g >= ob && g < oe;
                mine = mine || own;
                all_mine = all_mine && own;
            }
            if (!mine) continue;
            mine_elems.push_back((int32_t)e);
            if (all_mine) continue;  // interior element: nothing to mark, nothing to send
            int owners[4];
            for (int k = 0; k < nen; k++) owners[k] = owner_of(dofnode[en[k]], n_g, W);
            for (int k = 0; k < nen; k++) {
                const int32_t g = dofnode[en[k]];
                __atomic_store_n(&is_local[g], (uint8_t)1, __ATOMIC_RELAXED);  // several threads may mark the same node
                if (owners[k] == R)
                    for (int l = 0; l < nen; l++)
                        if (owners[l] != R) send[owners[l]].push_back(g);  // peer owners[l] needs my node g
            }
        }
    });
    std::vector<std::vector<int32_t>> send(W);
    for (int t = 0; t < T_max; t++) {
        p.loc_elems.insert(p.loc_elems.end(), part_elems[t].begin(), part_elems[t].end());
        for (int r = 0; r < W; r++) send[r].insert(send[r].end(), part_send[t][r].begin(), part_send[t][r].end());
    }
    for (int64_t g = 0; g < n_g; g++)
        if (is_local[g]) p.local_to_global.push_back((int32_t)g);
    const int64_t n_local = (int64_t)p.local_to_global.size();
    auto local_of = [&](int32_t g) {
        return (int32_t)(std::lower_bound(p.local_to_global.begin(), p.local_to_global.end(), g) - p.local_to_global.begin());
    };
    p.own_lo = local_of((int32_t)p.own_begin);

    // recv segments are contiguous per owner because local order == global order
    for (int r = 0; r < W; r++) {
        if (r == R) continue;
        PeerPlan pr;
        pr.rank = r;
        auto &s = send[r];
        std::sort(s.begin(), s.end());
        s.erase(std::unique(s.begin(), s.end()), s.end());
        pr.send_count = (int64_t)s.size();
        pr.send_off = (int64_t)p.send_idx.size();
        for (int32_t g : s) p.send_idx.push_back(local_of(g));
        const int64_t rb = (int64_t)r * n_g / W, re = (int64_t)(r + 1) * n_g / W;
        const int64_t first = std::lower_bound(p.local_to_global.begin(), p.local_to_global.end(), (int32_t)rb) - p.local_to_global.begin();
        const int64_t last = std::lower_bound(p.local_to_global.begin(), p.local_to_global.end(), (int32_t)re) - p.local_to_global.begin();
        pr.recv_count = last - first;
        pr.recv_off = first < n_local ? first : 0;
        if (pr.send_count || pr.recv_count) p.peers.push_back(pr);
    }
    return FS_OK;
}

}  // namespace fs

extern "C" int fs_partition_plan(int64_t n_nodes, int64_t n_elem, const int64_t *eptr, const int32_t *enodes, int dof_mode,
                                 int rank, int world, int64_t sizes[8], int32_t *local_to_global, int32_t *loc_elems,
                                 int32_t *send_idx, int64_t *peer_table)
{
    if (n_nodes <= 0 || n_elem <= 0 || !eptr || !enodes || !sizes) return FS_ERR_ARG;
    std::vector<int32_t> dofnode;
    const int64_t n_g = fs::compute_dof_order(dof_mode, n_nodes, n_elem, eptr, enodes, dofnode);
    fs::PartitionPlan p;
    int rc = fs::plan_partition(dofnode, n_g, n_elem, eptr, enodes, rank, world, p);
    if (rc) return rc;
    sizes[0] = p.n_global; sizes[1] = p.own_begin; sizes[2] = p.own_end; sizes[3] = p.own_lo;
    sizes[4] = (int64_t)p.local_to_global.size(); sizes[5] = (int64_t)p.loc_elems.size();
    sizes[6] = (int64_t)p.send_idx.size(); sizes[7] = (int64_t)p.peers.size();
    if (local_to_global) memcpy(local_to_global, p.local_to_global.data(), sizeof(int32_t) * p.local_to_global.size());
    if (loc_elems) memcpy(loc_elems, p.loc_elems.data(), sizeof(int32_t) * p.loc_elems.size());
    if (send_idx) memcpy(send_idx, p.send_idx.data(), sizeof(int32_t) * p.send_idx.size());
    if (peer_table)
        for (size_t i = 0; i < p.peers.size(); i++) {
            peer_table[5 * i + 0] = p.peers[i].rank;
            peer_table[5 * i + 1] = p.peers[i].send_count;
            peer_table[5 * i + 2] = p.peers[i].send_off;
            peer_table[5 * i + 3] = p.peers[i].recv_count;
            peer_table[5 * i + 4] = p.peers[i].recv_off;
        }
    return FS_OK;
}
