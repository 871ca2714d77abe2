// fs_partition.hpp -- host-only logic of mesh ingestion: DOF order and node-block partition plan.
// Kept free of CUDA so that the multi-rank logic can be exercised on CPU-only machines
// (tests/test_partition_gloo.py) through fs_partition_plan.
#pragma once
#include <cstdint>
#include <vector>

namespace fs {

struct PeerPlan {
    int rank = -1;
    int64_t send_count = 0, recv_count = 0;  // nodes
    int64_t send_off = 0;                    // offset into send_idx
    int64_t recv_off = 0;                    // first LOCAL node index of the contiguous recv segment
};

struct PartitionPlan {
    int64_t n_global = 0;                    // numbered dof-nodes
    int64_t own_begin = 0, own_end = 0;      // owned GLOBAL dof-node range
    int64_t own_lo = 0;                      // local index of the first owned node
    std::vector<int32_t> local_to_global;    // ascending: local order == global order
    std::vector<int32_t> loc_elems;          // ids of elements touching an owned node, ascending
    std::vector<PeerPlan> peers;
    std::vector<int32_t> send_idx;           // LOCAL node ids to pack, grouped by peer
};

// libMesh DofMap numbering (call sites fs.cpp:125,1205): mode 0 first-encounter, 1 node id.
// Returns the number of numbered nodes; dofnode[n] = -1 for nodes no element references.
int64_t compute_dof_order(int mode, int64_t n_nodes, int64_t n_elem, const int64_t *eptr, const int32_t *enodes,
                          std::vector<int32_t> &dofnode);

// rank r owns global dof-nodes [r*n/W, (r+1)*n/W)
int owner_of(int64_t g, int64_t n_g, int world);

// Every rank derives its own plan from the replicated mesh; the send list of rank a to rank b and the
// recv segment of rank b from rank a are the same set in the same (global) order by construction.
int plan_partition(const std::vector<int32_t> &dofnode, int64_t n_g, int64_t n_elem, const int64_t *eptr,
                   const int32_t *enodes, int rank, int world, PartitionPlan &plan, int threads = 1);

}  // namespace fs
