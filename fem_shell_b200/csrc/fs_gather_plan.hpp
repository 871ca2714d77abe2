// fs_gather_plan.hpp -- host-only schedule of the row-gather assembly pass (k_assemble_gather, fs_assembly.cu).
// Kept free of CUDA so that the packing and the emit-phase colouring can be exercised on CPU-only machines
// (tests/test_gather_plan.py) through fs_gather_plan.
#pragma once
#include <cstdint>
#include <vector>

namespace fs {

// one warp of the row-gather assembly: a run of block rows accumulated in shared memory
struct GatherChunk {
    long long val_off;   // first CSR value of the run (36 * nptr[row0])
    int val_count;       // 36 * (nptr[row1] - nptr[row0]) doubles staged in shared memory
    int n_rounds;        // emit phases of the chunk
    int n_threads;       // valid entries among the chunk's 32 thread-table slots
    int pad;
};

// Thread table: 32 entries per chunk, entry = one (element, node row I) incidence of an owned block row.
//   info[4*k+0] = I | is_quad << 2 | phase << 3 | valid << 8
//   info[4*k+1] = offset of the row inside the chunk's shared-memory slice (doubles) | blocks of the row << 16
//   info[4*k+2] = Dirichlet bits of the element's nodes, 8 bits each
//   info[4*k+3] = slot of node j in the row, 8 bits each
//   nodes[4*k+j] = local id of the element's j-th node
struct GatherPlan {
    std::vector<GatherChunk> chunks;
    std::vector<int32_t> info, nodes;
};

// tri/quad: connectivity in local node ids; *_gid: element ids (summation order); *_pos[e*n*n + I*n + j]: slot of
// node j in the block row of node I; nptr: block-row pointers of the owned rows; mask: Dirichlet bits per local node.
// Returns false when a block row does not fit one warp (more than 32 incident elements, more than warp_vals
// values, more than 255 blocks, or more than 32 mutually clashing elements): the caller uses the coloured pass.
bool plan_gather(int64_t n_own, int own_lo, int64_t nt, const int32_t *tri, const int32_t *tgid, const int32_t *tpos, int64_t nq,
                 const int32_t *quad, const int32_t *qgid, const int32_t *qpos, const int32_t *nptr, const uint8_t *mask,
                 int warp_vals, GatherPlan &plan);

}  // namespace fs
