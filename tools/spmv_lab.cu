// spmv_lab.cu -- stand-alone micro-benchmark used to choose the SpMV kernel shape (not product code).
// Builds the block-CSR pattern of a structured nx x ny Quad-4 or Tri-3 plate with synthetic values and
// times kernel variants with CUDA events.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

enum LoadKind { LD_PLAIN = 0, LD_CS = 1, LD_NC = 2, LD_LU = 3 };

template <int LK>
__device__ __forceinline__ double2 ldv(const double2 *p)
{
    if (LK == LD_CS) return __ldcs(p);
    if (LK == LD_NC) return __ldg(p);
    if (LK == LD_LU) return __ldlu(p);
    return *p;
}

__device__ __forceinline__ void reduce6_store(double acc[6], int lane, double *y, size_t p)
{
    double t3[4];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double send = hi ? acc[k] : acc[k + 3];
            double keep = hi ? acc[k + 3] : acc[k];
            t3[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        t3[3] = 0.0;
    }
    double t2[2];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            double send = hi ? t3[k] : t3[k + 2];
            double keep = hi ? t3[k + 2] : t3[k];
            t2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    double t1;
    {
        const bool hi = lane & 4;
        double send = hi ? t2[0] : t2[1];
        double keep = hi ? t2[1] : t2[0];
        t1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
    t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
    const int row = ((lane & 16) ? 3 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
    const bool writer = ((lane & 3) == 0) && (((lane >> 2) & 3) != 3);
    if (writer) y[6 * p + row] = t1;
}

// variant A: one warp per block row, grid-stride (the round-1 first-pass kernel), UNROLL nodes in flight
template <int LK, int BLOCK, int MINB, int UNROLL>
__global__ void __launch_bounds__(BLOCK, MINB)
spmv_a(int n, const int *__restrict__ nptr, const int *__restrict__ nadj, const double *__restrict__ vals,
       const double *__restrict__ x, double *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    const int nw = gridDim.x * (BLOCK / 32);
    for (int p0 = gw * UNROLL; p0 < n; p0 += nw * UNROLL) {
        double acc[UNROLL][6];
        double2 v[UNROLL][6];
        double2 xv[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int p = p0 + u;
            ok[u] = false;
#pragma unroll
            for (int a = 0; a < 6; a++) acc[u][a] = 0.0;
            if (p < n) {
                const int b0 = nptr[p], deg = nptr[p + 1] - b0;
                const int L2 = 3 * deg;
                if (lane < L2) {
                    ok[u] = true;
                    const double2 *base = reinterpret_cast<const double2 *>(vals + (size_t)36 * b0);
                    const int j = lane / 3, h = lane - 3 * j;
                    const int col = nadj[b0 + j];
                    xv[u] = *reinterpret_cast<const double2 *>(x + 6 * (size_t)col + 2 * h);
#pragma unroll
                    for (int a = 0; a < 6; a++) v[u][a] = ldv<LK>(base + (size_t)a * L2 + lane);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (ok[u]) {
#pragma unroll
                for (int a = 0; a < 6; a++) acc[u][a] = v[u][a].x * xv[u].x + v[u][a].y * xv[u].y;
            }
            const int p = p0 + u;
            if (p < n) {
                // rows longer than 32 double2 (deg > 10): remaining chunks
                const int b0 = nptr[p], deg = nptr[p + 1] - b0;
                const int L2 = 3 * deg;
                const double2 *base = reinterpret_cast<const double2 *>(vals + (size_t)36 * b0);
                for (int l = lane + 32; l < L2; l += 32) {
                    const int j = l / 3, h = l - 3 * j;
                    const int col = nadj[b0 + j];
                    const double2 xx = *reinterpret_cast<const double2 *>(x + 6 * (size_t)col + 2 * h);
#pragma unroll
                    for (int a = 0; a < 6; a++) {
                        double2 vv = ldv<LK>(base + (size_t)a * L2 + l);
                        acc[u][a] += vv.x * xx.x + vv.y * xx.y;
                    }
                }
                reduce6_store(acc[u], lane, y, (size_t)p);
            }
        }
    }
}

// pure streaming read of the value array: the ceiling any SpMV can reach on this access pattern
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) stream_read(const double2 *__restrict__ v, size_t n2, double *out)
{
    double s = 0.0;
    for (size_t i = blockIdx.x * (size_t)BLOCK + threadIdx.x; i < n2; i += (size_t)gridDim.x * BLOCK) {
        double2 t = __ldcs(v + i);
        s += t.x + t.y;
    }
    if (s == 1.2345e-300) out[0] = s;
}

struct Case {
    int n;
    std::vector<int> nptr, nadj;
};

static Case make_grid(int nx, int ny, bool tri)
{
    Case c;
    c.n = nx * ny;
    c.nptr.assign(c.n + 1, 0);
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            std::vector<int> nb;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    if (tri && dx * dy == 1) continue;  // ul_lr diagonal: (x+1,y+1) and (x-1,y-1) are not neighbours
                    int xx = x + dx, yy = y + dy;
                    if (xx < 0 || yy < 0 || xx >= nx || yy >= ny) continue;
                    nb.push_back(yy * nx + xx);
                }
            std::sort(nb.begin(), nb.end());
            for (int v : nb) c.nadj.push_back(v);
            c.nptr[y * nx + x + 1] = (int)c.nadj.size();
        }
    return c;
}

template <class F>
static float time_it(F f, int reps)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    f();
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
    int nx = argc > 1 ? atoi(argv[1]) : 1000, ny = argc > 2 ? atoi(argv[2]) : 1000;
    bool tri = argc > 3 && atoi(argv[3]);
    Case c = make_grid(nx, ny, tri);
    const size_t nb = c.nadj.size(), nv = 36 * nb;
    int *d_nptr, *d_nadj;
    double *d_vals, *d_x, *d_y, *d_y0;
    CK(cudaMalloc(&d_nptr, sizeof(int) * (c.n + 1)));
    CK(cudaMalloc(&d_nadj, sizeof(int) * nb));
    CK(cudaMalloc(&d_vals, sizeof(double) * nv));
    CK(cudaMalloc(&d_x, sizeof(double) * 6 * c.n));
    CK(cudaMalloc(&d_y, sizeof(double) * 6 * c.n));
    CK(cudaMalloc(&d_y0, sizeof(double) * 6 * c.n));
    CK(cudaMemcpy(d_nptr, c.nptr.data(), sizeof(int) * (c.n + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_nadj, c.nadj.data(), sizeof(int) * nb, cudaMemcpyHostToDevice));
    std::vector<double> hv(nv), hx(6 * (size_t)c.n);
    unsigned s = 12345;
    for (auto &v : hv) { s = s * 1664525u + 1013904223u; v = (double)(s >> 8) / (1 << 24) - 0.5; }
    for (auto &v : hx) { s = s * 1664525u + 1013904223u; v = (double)(s >> 8) / (1 << 24) - 0.5; }
    CK(cudaMemcpy(d_vals, hv.data(), sizeof(double) * nv, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_x, hx.data(), sizeof(double) * 6 * c.n, cudaMemcpyHostToDevice));
    int sm = 148;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    sm = prop.multiProcessorCount;
    const double bytes = 8.0 * nv + 4.0 * nb + 4.0 * (c.n + 1) + 16.0 * 6 * c.n;
    printf("grid %dx%d %s: n=%d blocks=%zu values %.3f GB, bytes/launch %.3f GB, SMs %d\n", nx, ny, tri ? "tri" : "quad", c.n, nb,
           8e-9 * nv, 1e-9 * bytes, sm);

    {
        float ms = time_it([&] { stream_read<256><<<sm * 8, 256>>>((const double2 *)d_vals, nv / 2, d_y); }, 20);
        printf("%-44s %8.3f ms  %8.1f GB/s (values only)\n", "stream_read 256x(8/SM) ldcs", ms, 8e-9 * nv / (ms * 1e-3));
        ms = time_it([&] { stream_read<512><<<sm * 4, 512>>>((const double2 *)d_vals, nv / 2, d_y); }, 20);
        printf("%-44s %8.3f ms  %8.1f GB/s (values only)\n", "stream_read 512x(4/SM) ldcs", ms, 8e-9 * nv / (ms * 1e-3));
    }
    // reference result from the baseline variant
    spmv_a<LD_CS, 256, 1, 1><<<sm * 3, 256>>>(c.n, d_nptr, d_nadj, d_vals, d_x, d_y0);
    CK(cudaDeviceSynchronize());
    std::vector<double> y0(6 * (size_t)c.n), y1(6 * (size_t)c.n);
    CK(cudaMemcpy(y0.data(), d_y0, sizeof(double) * 6 * c.n, cudaMemcpyDeviceToHost));

#define RUN(name, kern, grid, block)                                                                         \
    {                                                                                                        \
        CK(cudaMemset(d_y, 0, sizeof(double) * 6 * c.n));                                                    \
        float ms = time_it([&] { kern<<<(grid), (block)>>>(c.n, d_nptr, d_nadj, d_vals, d_x, d_y); }, 20);   \
        CK(cudaMemcpy(y1.data(), d_y, sizeof(double) * 6 * c.n, cudaMemcpyDeviceToHost));                    \
        double md = 0;                                                                                       \
        for (size_t i = 0; i < y1.size(); i++) md = std::max(md, fabs(y1[i] - y0[i]));                       \
        int nbk = 0;                                                                                         \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, kern, block, 0);                                 \
        printf("%-44s %8.3f ms  %8.1f GB/s  occ %d blk/SM  maxdiff %.1e\n", name, ms, 1e-9 * bytes / (ms * 1e-3), nbk, md); \
    }
    RUN("A cs  256 minb1 u1 grid 3/SM", (spmv_a<LD_CS, 256, 1, 1>), sm * 3, 256)
    RUN("A cs  256 minb4 u1 grid 4/SM", (spmv_a<LD_CS, 256, 4, 1>), sm * 4, 256)
    RUN("A cs  256 minb4 u1 grid 8/SM", (spmv_a<LD_CS, 256, 4, 1>), sm * 8, 256)
    RUN("A cs  256 minb4 u1 grid n/8", (spmv_a<LD_CS, 256, 4, 1>), (c.n + 7) / 8, 256)
    RUN("A cs  128 minb8 u1 grid 8/SM", (spmv_a<LD_CS, 128, 8, 1>), sm * 8, 128)
    RUN("A cs  512 minb2 u1 grid 2/SM", (spmv_a<LD_CS, 512, 2, 1>), sm * 2, 512)
    RUN("A cs  256 minb2 u2 grid 2/SM", (spmv_a<LD_CS, 256, 2, 2>), sm * 2, 256)
    RUN("A cs  256 minb2 u2 grid 4/SM", (spmv_a<LD_CS, 256, 2, 2>), sm * 4, 256)
    RUN("A cs  256 minb3 u2 grid 3/SM", (spmv_a<LD_CS, 256, 3, 2>), sm * 3, 256)
    RUN("A cs  256 minb1 u4 grid 2/SM", (spmv_a<LD_CS, 256, 1, 4>), sm * 2, 256)
    RUN("A pl  256 minb4 u1 grid 4/SM", (spmv_a<LD_PLAIN, 256, 4, 1>), sm * 4, 256)
    RUN("A nc  256 minb4 u1 grid 4/SM", (spmv_a<LD_NC, 256, 4, 1>), sm * 4, 256)
    RUN("A lu  256 minb4 u1 grid 4/SM", (spmv_a<LD_LU, 256, 4, 1>), sm * 4, 256)
    RUN("A nc  256 minb2 u2 grid 4/SM", (spmv_a<LD_NC, 256, 2, 2>), sm * 4, 256)
    RUN("A cs  1024 minb1 u1 grid 1/SM", (spmv_a<LD_CS, 1024, 1, 1>), sm * 1, 1024)
    return 0;
}
