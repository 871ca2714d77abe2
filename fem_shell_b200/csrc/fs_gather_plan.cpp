// fs_gather_plan.cpp -- see fs_gather_plan.hpp
#include "fs_gather_plan.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>

#include "../../include/femshell_b200.h"

namespace fs {

// fn(begin, end) on contiguous pieces of [0, n); the pieces are independent (rows of the matrix / chunks of the table)
template <class F>
static void parallel_ranges(int64_t n, F fn)
{
    const unsigned hw = std::thread::hardware_concurrency();
    int64_t T = std::max<int64_t>(1, std::min<int64_t>({(int64_t)(hw ? hw : 1), (int64_t)8, n / 65536}));
    if (const char *forced = getenv("FS_PLAN_THREADS"))  // tests: the plan must not depend on the number of threads
        T = std::max<int64_t>(1, std::min<int64_t>({(int64_t)atoi(forced), (int64_t)64, std::max<int64_t>(n, 1)}));
    if (T == 1) {
        fn((int64_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int64_t t = 0; t < T; t++) pool.emplace_back(fn, n * t / T, n * (t + 1) / T);
    for (std::thread &th : pool) th.join();
}

bool plan_gather(int64_t n_own, int own_lo, int64_t nt, const int32_t *tri, const int32_t *tgid, const int32_t *tpos, int64_t nq,
                 const int32_t *quad, const int32_t *qgid, const int32_t *qpos, const int32_t *nptr, const uint8_t *mask,
                 int warp_vals, GatherPlan &plan)
{
    struct Inc { int32_t row, gid, eidx; uint8_t type, I; };
    std::vector<int32_t> cnt(n_own + 1, 0);
    for (int64_t e = 0; e < nt; e++)
        for (int k = 0; k < 3; k++) { int p = tri[3 * e + k] - own_lo; if (p >= 0 && p < n_own) cnt[p + 1]++; }
    for (int64_t e = 0; e < nq; e++)
        for (int k = 0; k < 4; k++) { int p = quad[4 * e + k] - own_lo; if (p >= 0 && p < n_own) cnt[p + 1]++; }
    for (int64_t p = 0; p < n_own; p++) cnt[p + 1] += cnt[p];
    std::vector<Inc> inc(cnt[n_own]);
    std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
    for (int64_t e = 0; e < nt; e++)
        for (int k = 0; k < 3; k++) { int p = tri[3 * e + k] - own_lo; if (p >= 0 && p < n_own) inc[fill[p]++] = {(int32_t)p, tgid[e], (int32_t)e, 0, (uint8_t)k}; }
    for (int64_t e = 0; e < nq; e++)
        for (int k = 0; k < 4; k++) { int p = quad[4 * e + k] - own_lo; if (p >= 0 && p < n_own) inc[fill[p]++] = {(int32_t)p, qgid[e], (int32_t)e, 1, (uint8_t)k}; }
    // Phases.  All lanes emit block j (the element's j-th node column) in the same step, so two incidences of
    // one row may share a step iff they never meet in a slot at the same j, i.e. no node sits at the same local
    // index in both elements.  On a structured Quad-4 mesh the four elements around a node see every shared
    // neighbour under different local indices -> one phase; meshGen's triangle pairs share their hypotenuse end
    // node at the same index -> a few phases.  Greedy colouring in a fixed order (triangles first, then by
    // element id) keeps the summation order of every CSR value a function of the mesh alone.
    std::vector<uint8_t> phase(inc.size(), 0);
    std::atomic<bool> too_many(false);
    parallel_ranges(n_own, [&](int64_t p0, int64_t p1) {
        for (int64_t p = p0; p < p1; p++) {
            std::sort(inc.begin() + cnt[p], inc.begin() + cnt[p + 1], [](const Inc &a, const Inc &b) {
                return a.type != b.type ? a.type < b.type : a.gid < b.gid;
            });
            for (int k = cnt[p]; k < cnt[p + 1]; k++) {
                const int nen = inc[k].type ? 4 : 3;
                const int32_t *ek = inc[k].type ? &quad[4 * (int64_t)inc[k].eidx] : &tri[3 * (int64_t)inc[k].eidx];
                unsigned used = 0;
                for (int m = cnt[p]; m < k; m++) {
                    if (inc[m].type != inc[k].type) continue;  // quads and triangles are emitted one after the other
                    const int32_t *em = inc[m].type ? &quad[4 * (int64_t)inc[m].eidx] : &tri[3 * (int64_t)inc[m].eidx];
                    bool clash = false;
                    for (int j = 0; j < nen; j++) clash |= ek[j] == em[j];
                    if (clash) used |= 1u << phase[m];
                }
                int ph = 0;
                while (ph < 32 && ((used >> ph) & 1u)) ph++;
                if (ph > 31) {  // more than 32 mutually clashing elements at one node
                    too_many = true;
                    ph = 31;
                }
                phase[k] = (uint8_t)ph;
            }
        }
    });
    if (too_many) return false;

    // packing: a sequential scan over the rows (cheap); the table is then filled chunk by chunk in parallel
    std::vector<GatherChunk> &chunks = plan.chunks;
    std::vector<int32_t> &g_info = plan.info, &g_nodes = plan.nodes;  // 32 entries of 4 ints per chunk
    chunks.clear(); g_info.clear(); g_nodes.clear();
    std::vector<int64_t> first_row;
    int64_t row = 0;
    while (row < n_own) {
        int64_t r1 = row, vals = 0;
        int threads = 0, rounds = 0;
        while (r1 < n_own) {
            const int t2 = threads + (cnt[r1 + 1] - cnt[r1]);
            const int64_t v2 = vals + 36 * (int64_t)(nptr[r1 + 1] - nptr[r1]);
            if (t2 > 32 || v2 > warp_vals || nptr[r1 + 1] - nptr[r1] > 255) break;
            threads = t2;
            vals = v2;
            for (int k = cnt[r1]; k < cnt[r1 + 1]; k++) rounds = std::max(rounds, (int)phase[k] + 1);
            r1++;
        }
        if (r1 == row) return false;  // a single row does not fit a warp: leave this mesh to the coloured pass
        GatherChunk ch;
        ch.val_off = 36 * (long long)nptr[row];
        ch.n_threads = threads;
        ch.n_rounds = rounds;
        ch.val_count = (int)vals;
        ch.pad = 0;
        chunks.push_back(ch);
        first_row.push_back(row);
        row = r1;
    }
    first_row.push_back(n_own);
    g_info.assign(chunks.size() * 4 * 32, 0);
    g_nodes.assign(chunks.size() * 4 * 32, 0);
    parallel_ranges((int64_t)chunks.size(), [&](int64_t c0, int64_t c1) {
        for (int64_t ci = c0; ci < c1; ci++) {
            const int64_t row0 = first_row[ci], row1 = first_row[ci + 1];
            size_t at = (size_t)ci * 4 * 32;
            // quads first so that a mixed chunk splits into at most two divergent halves
            for (int pass = 1; pass >= 0; pass--)
                for (int64_t p = row0; p < row1; p++)
                    for (int k = cnt[p]; k < cnt[p + 1]; k++)
                        if (inc[k].type == pass) {
                            const int nen = pass ? 4 : 3, I = inc[k].I;
                            const int32_t *en = pass ? &quad[4 * (int64_t)inc[k].eidx] : &tri[3 * (int64_t)inc[k].eidx];
                            const int32_t *ps = pass ? &qpos[16 * (int64_t)inc[k].eidx + 4 * I] : &tpos[9 * (int64_t)inc[k].eidx + 3 * I];
                            unsigned slots = 0, mbits = 0;
                            for (int j = 0; j < nen; j++) {
                                g_nodes[at + j] = en[j];
                                slots |= (unsigned)(ps[j] & 0xff) << (8 * j);
                                mbits |= (unsigned)(mask[en[j]] & 0x3f) << (8 * j);
                            }
                            const unsigned soff = (unsigned)(36 * (nptr[p] - nptr[row0]));
                            const unsigned deg = (unsigned)(nptr[p + 1] - nptr[p]);
                            g_info[at] = I | (inc[k].type << 2) | ((int)phase[k] << 3) | (1 << 8);
                            g_info[at + 1] = (int)(soff | (deg << 16));
                            g_info[at + 2] = (int)mbits;
                            g_info[at + 3] = (int)slots;
                            at += 4;
                        }
        }
    });
    return true;
}

}  // namespace fs

// ---------------------------------------------------------------------------------------------
// C-ABI test hook (declared in include/femshell_b200.h): the schedule fs_assemble would use for a single-rank
// mesh given in dof-node numbering.  Builds the node-block pattern on the host the way the device pattern pass
// does (sorted unique neighbours per node) and runs plan_gather on it.
// ---------------------------------------------------------------------------------------------
extern "C" int fs_gather_plan(int64_t n_nodes, int64_t n_elem, const int32_t *etype, const int64_t *eptr, const int32_t *enodes,
                              const uint8_t *mask, int warp_vals, int64_t sizes[3], int64_t *chunks, int32_t *info, int32_t *nodes,
                              int32_t *nptr_out, int32_t *nadj_out)
{
    if (n_nodes <= 0 || n_elem <= 0 || !etype || !eptr || !enodes || !sizes || warp_vals <= 0) return FS_ERR_ARG;
    std::vector<std::vector<int32_t>> adj(n_nodes);
    std::vector<int32_t> tri, quad, tgid, qgid;
    for (int64_t e = 0; e < n_elem; e++) {
        const int nen = (int)(eptr[e + 1] - eptr[e]);
        if ((etype[e] == FS_TRI3 && nen != 3) || (etype[e] == FS_QUAD4 && nen != 4) || (etype[e] != FS_TRI3 && etype[e] != FS_QUAD4)) return FS_ERR_ARG;
        const int32_t *en = enodes + eptr[e];
        for (int i = 0; i < nen; i++) {
            if (en[i] < 0 || en[i] >= n_nodes) return FS_ERR_ARG;
            for (int j = 0; j < nen; j++) adj[en[i]].push_back(en[j]);
            (nen == 3 ? tri : quad).push_back(en[i]);
        }
        (nen == 3 ? tgid : qgid).push_back((int32_t)e);
    }
    std::vector<int32_t> nptr(n_nodes + 1, 0), nadj;
    for (int64_t p = 0; p < n_nodes; p++) {
        std::sort(adj[p].begin(), adj[p].end());
        adj[p].erase(std::unique(adj[p].begin(), adj[p].end()), adj[p].end());
        nptr[p + 1] = nptr[p] + (int32_t)adj[p].size();
        nadj.insert(nadj.end(), adj[p].begin(), adj[p].end());
    }
    auto positions = [&](const std::vector<int32_t> &conn, int nen) {
        std::vector<int32_t> pos(conn.size() * nen);
        for (size_t e = 0; e < conn.size() / nen; e++)
            for (int i = 0; i < nen; i++)
                for (int j = 0; j < nen; j++) {
                    const std::vector<int32_t> &a = adj[conn[e * nen + i]];
                    pos[(e * nen + i) * nen + j] = (int32_t)(std::lower_bound(a.begin(), a.end(), conn[e * nen + j]) - a.begin());
                }
        return pos;
    };
    const std::vector<int32_t> tpos = positions(tri, 3), qpos = positions(quad, 4);
    std::vector<uint8_t> zero_mask;
    if (!mask) {
        zero_mask.assign(n_nodes, 0);
        mask = zero_mask.data();
    }
    fs::GatherPlan plan;
    const bool ok = fs::plan_gather(n_nodes, 0, (int64_t)tgid.size(), tri.data(), tgid.data(), tpos.data(), (int64_t)qgid.size(), quad.data(),
                                    qgid.data(), qpos.data(), nptr.data(), mask, warp_vals, plan);
    sizes[0] = ok ? (int64_t)plan.chunks.size() : 0;
    sizes[1] = ok ? 1 : 0;
    sizes[2] = (int64_t)nadj.size();
    if (nptr_out) std::copy(nptr.begin(), nptr.end(), nptr_out);
    if (nadj_out) std::copy(nadj.begin(), nadj.end(), nadj_out);
    if (!ok) return FS_OK;
    if (chunks)
        for (size_t i = 0; i < plan.chunks.size(); i++) {
            chunks[4 * i + 0] = plan.chunks[i].val_off;
            chunks[4 * i + 1] = plan.chunks[i].val_count;
            chunks[4 * i + 2] = plan.chunks[i].n_rounds;
            chunks[4 * i + 3] = plan.chunks[i].n_threads;
        }
    if (info) std::copy(plan.info.begin(), plan.info.end(), info);
    if (nodes) std::copy(plan.nodes.begin(), plan.nodes.end(), nodes);
    return FS_OK;
}
