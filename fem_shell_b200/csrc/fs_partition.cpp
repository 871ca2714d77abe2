// fs_partition.cpp -- see fs_partition.hpp
#include "fs_partition.hpp"

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
                   const int32_t *enodes, int R, int W, PartitionPlan &p)
{
    if (W < 1 || R < 0 || R >= W || n_g < W) return FS_ERR_ARG;
    p = PartitionPlan();
    p.n_global = n_g;
    p.own_begin = (int64_t)R * n_g / W;
    p.own_end = (int64_t)(R + 1) * n_g / W;

    // local elements = elements touching an owned node; local nodes = their nodes + owned nodes
    std::vector<uint8_t> is_local(n_g, 0);
    for (int64_t g = p.own_begin; g < p.own_end; g++) is_local[g] = 1;
    std::vector<std::vector<int32_t>> send(W);
    for (int64_t e = 0; e < n_elem; e++) {
        bool mine = false;
        int owners[4];
        const int nen = (int)(eptr[e + 1] - eptr[e]);
        for (int k = 0; k < nen; k++) {
            const int64_t g = dofnode[enodes[eptr[e] + k]];
            owners[k] = (W == 1) ? 0 : owner_of(g, n_g, W);
            mine = mine || owners[k] == R;
        }
        if (!mine) continue;
        p.loc_elems.push_back((int32_t)e);
        for (int k = 0; k < nen; k++) {
            const int32_t g = dofnode[enodes[eptr[e] + k]];
            is_local[g] = 1;
            if (owners[k] == R)
                for (int l = 0; l < nen; l++)
                    if (owners[l] != R) send[owners[l]].push_back(g);  // peer owners[l] needs my node g
        }
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
