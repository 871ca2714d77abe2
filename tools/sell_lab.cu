// sell_lab.cu -- stand-alone micro-benchmark used to choose the launch shape of k_spmv_sell (not product code).
// Synthetic SELL-32 matrix of a structured nx x ny Quad-4 / Tri-3 plate with the planar 14-of-36 pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I fem_shell_b200/csrc tools/sell_lab.cu -o tools/bin/sell_lab
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fs_sell.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
using namespace fs;

template <class F>
static float time_it(F f, int reps)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / reps;
}

int main(int argc, char **argv)
{
    const int nx = argc > 1 ? atoi(argv[1]) : 1000, ny = argc > 2 ? atoi(argv[2]) : 1000;
    const bool tri = argc > 3 && atoi(argv[3]);
    const int n = nx * ny, n_slices = (n + 31) / 32;
    constexpr int NZ = sell_popcount(SELL_MASK_XY);
    std::vector<std::vector<int>> nb(n);
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    if (tri && dx * dy == 1) continue;
                    int xx = x + dx, yy = y + dy;
                    if (xx < 0 || yy < 0 || xx >= nx || yy >= ny) continue;
                    nb[y * nx + x].push_back(yy * nx + xx);
                }
    std::vector<int> sptr(n_slices + 1, 0);
    for (int s = 0; s < n_slices; s++) {
        int d = 0;
        for (int l = 0; l < 32 && 32 * s + l < n; l++) d = std::max(d, (int)nb[32 * s + l].size());
        sptr[s + 1] = sptr[s] + d;
    }
    const size_t slots = sptr[n_slices];
    std::vector<int> adj(32 * slots);
    std::vector<double> vals(32 * NZ * slots), hx(6 * (size_t)n);
    unsigned r = 12345;
    auto rnd = [&] { r = r * 1664525u + 1013904223u; return (double)(r >> 8) / (1 << 24) - 0.5; };
    size_t real_blocks = 0;
    for (int s = 0; s < n_slices; s++)
        for (int l = 0; l < 32; l++) {
            const int p = 32 * s + l;
            const int deg = p < n ? (int)nb[p].size() : 0;
            real_blocks += deg;
            for (int k = 0; k < sptr[s + 1] - sptr[s]; k++) {
                adj[32 * (size_t)(sptr[s] + k) + l] = k < deg ? nb[p][k] : std::min(p, n - 1);
                for (int i = 0; i < NZ; i++) vals[32 * ((size_t)NZ * (sptr[s] + k) + i) + l] = k < deg ? rnd() : 0.0;
            }
        }
    for (auto &v : hx) v = rnd();
    int *d_sptr, *d_adj;
    double *d_vals, *d_x, *d_y, *d_y0;
    CK(cudaMalloc(&d_sptr, sizeof(int) * sptr.size()));
    CK(cudaMalloc(&d_adj, sizeof(int) * adj.size()));
    CK(cudaMalloc(&d_vals, sizeof(double) * vals.size()));
    CK(cudaMalloc(&d_x, sizeof(double) * 6 * n));
    CK(cudaMalloc(&d_y, sizeof(double) * 6 * n));
    CK(cudaMalloc(&d_y0, sizeof(double) * 6 * n));
    CK(cudaMemcpy(d_sptr, sptr.data(), sizeof(int) * sptr.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_adj, adj.data(), sizeof(int) * adj.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_vals, vals.data(), sizeof(double) * vals.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_x, hx.data(), sizeof(double) * 6 * n, cudaMemcpyHostToDevice));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sm = prop.multiProcessorCount;
    const double bytes = 8.0 * vals.size() + 4.0 * adj.size() + 4.0 * sptr.size() + 16.0 * 6 * n;
    printf("grid %dx%d %s: n=%d blocks=%zu padded=%zu values %.3f GB, bytes/launch %.3f GB, SMs %d\n", nx, ny, tri ? "tri" : "quad", n,
           real_blocks, 32 * slots, 8e-9 * vals.size(), 1e-9 * bytes, sm);
    // reference on the host for the first rows
    std::vector<double> y0(6 * (size_t)n), y1(6 * (size_t)n);
    k_spmv_sell<SELL_MASK_XY, false, 128, 4, 1><<<sm * 4, 128>>>(n, n_slices, d_sptr, d_adj, d_vals, d_x, d_y0, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(y0.data(), d_y0, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost));
    {
        double md = 0;
        for (int p = 0; p < std::min(n, 5000); p++) {
            const int s = p / 32, l = p % 32;
            double acc[6] = {0, 0, 0, 0, 0, 0};
            for (int k = 0; k < (int)nb[p].size(); k++) {
                int item = 0;
                for (int a = 0; a < 6; a++)
                    for (int b = 0; b < 6; b++)
                        if (SELL_MASK_XY & sell_bit(a, b)) {
                            acc[a] += vals[32 * ((size_t)NZ * (sptr[s] + k) + item) + l] * hx[6 * (size_t)nb[p][k] + b];
                            item++;
                        }
            }
            for (int a = 0; a < 6; a++) md = std::max(md, fabs(acc[a] - y0[6 * (size_t)p + a]));
        }
        printf("host check of the first rows: maxdiff %.2e\n", md);
    }
#define RUN(name, B, M, U, gridmul)                                                                                       \
    {                                                                                                                     \
        auto kern = k_spmv_sell<SELL_MASK_XY, false, B, M, U>;                                                            \
        int nbk = 0;                                                                                                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, kern, B, 0);                                                  \
        const int grid = std::min((n_slices + B / 32 - 1) / (B / 32), sm * (gridmul > 0 ? gridmul : nbk));               \
        CK(cudaMemset(d_y, 0, sizeof(double) * 6 * n));                                                                   \
        float ms = time_it([&] { kern<<<grid, B>>>(n, n_slices, d_sptr, d_adj, d_vals, d_x, d_y, nullptr, nullptr, nullptr, nullptr, nullptr, 0); }, 30); \
        CK(cudaMemcpy(y1.data(), d_y, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost));                                   \
        double md = 0;                                                                                                    \
        for (size_t i = 0; i < y1.size(); i++) md = std::max(md, fabs(y1[i] - y0[i]));                                    \
        cudaFuncAttributes fa;                                                                                            \
        cudaFuncGetAttributes(&fa, kern);                                                                                 \
        printf("%-34s %8.4f ms  %8.1f GB/s  regs %3d occ %2d blk/SM grid %5d maxdiff %.1e\n", name, ms, 1e-9 * bytes / (ms * 1e-3), fa.numRegs, nbk, grid, md); \
    }
    RUN("B128 minb4 u1", 128, 4, 1, 0)
    RUN("B128 minb4 u2", 128, 4, 2, 0)
    RUN("B128 minb4 u3", 128, 4, 3, 0)
    RUN("B128 minb6 u1", 128, 6, 1, 0)
    RUN("B128 minb6 u2", 128, 6, 2, 0)
    RUN("B128 minb8 u1", 128, 8, 1, 0)
    RUN("B128 minb8 u2", 128, 8, 2, 0)
    RUN("B128 minb12 u1", 128, 12, 1, 0)
    RUN("B128 minb16 u1", 128, 16, 1, 0)
    RUN("B64 minb8 u2", 64, 8, 2, 0)
    RUN("B64 minb16 u1", 64, 16, 1, 0)
    RUN("B64 minb16 u2", 64, 16, 2, 0)
    RUN("B256 minb2 u2", 256, 2, 2, 0)
    RUN("B256 minb4 u1", 256, 4, 1, 0)
    RUN("B256 minb4 u2", 256, 4, 2, 0)
    RUN("B128 minb8 u2 grid=all", 128, 8, 2, 100000)
    RUN("B128 minb4 u2 grid=all", 128, 4, 2, 100000)
    return 0;
}
