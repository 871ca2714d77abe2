// fs_bench.cu -- measurement helpers exported for bench.py: the FP64 FMA peak of the device, which is the
// roofline of the element kernels (MEASURED_PEAKS.json only carries HBM and bf16 figures; SURVEY.md 8d asks
// for an FMA micro-benchmark).
#include <algorithm>
#include <cmath>
#include <vector>

#include "fs_context.hpp"

namespace fs {

// 8 independent dependent-FMA chains per thread, 4096 FMAs each: pure DFMA issue
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, double a, double b, int iters)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------------------------------------
// FP64 DMMA against the FMA pipe on the contraction of the plate kernel (north_star (a): "use FP64 DMMA only if ncu
// shows the batched B^T D B contraction beats the FMA pipe").  Both kernels form, for every element of a synthetic
// batch, the DKQ plate matrix Kp = sum over 4 Gauss points of (Dp B)^T B with B 3 x 12 (fs.cpp:664-681) from the
// same closed-form B entries and reduce it to one checksum per element, so that neither is bound by HBM.
//   DFMA  thread = (element, node row I): 3 x 12 accumulators, 108 FMAs per Gauss point -- the production layout
//   DMMA  warp = element: mma.sync.m8n8k4.f64, K = (Gauss point, strain) = 12 = 3 k-steps, the 12 x 12 result padded
//         to 2 x 2 tiles of 8 x 8 (12 DMMAs, 56 % of their MACs useful); fragments are generated in place, i.e. the
//         lane-to-lane shuffles a real element kernel would need to bring B columns into fragment layout are NOT
//         charged to this variant
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double bench_b(double base, int gp, int r, int c) { return fma(base, 0.015625 * (double)(gp * 36 + r * 12 + c + 1), 0.25 * (double)(r + 1)); }
__device__ __forceinline__ double bench_w(int m, int n) { return 1.0 + 0.0078125 * (double)(m * 12 + n); }

__global__ void __launch_bounds__(128) k_contract_dfma(int64_t n_elem, double d11, double d12, double d33, double *__restrict__ out)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t e = t >> 2;
    const int I = (int)(t & 3);
    if (e >= n_elem) return;
    const double base = 1.0 + 1e-6 * (double)(e & 1023);
    double acc[3][12];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 12; c++) acc[r][c] = 0.0;
#pragma unroll
    for (int gp = 0; gp < 4; gp++) {
        double B[3][12];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 12; c++) B[r][c] = bench_b(base, gp, r, c);
        double E[3][3];  // Dp * B_I, columns 3I..3I+2 picked with selects like the production kernel's run-time node row
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double b0 = I == 0 ? B[0][c] : (I == 1 ? B[0][3 + c] : (I == 2 ? B[0][6 + c] : B[0][9 + c]));
            const double b1 = I == 0 ? B[1][c] : (I == 1 ? B[1][3 + c] : (I == 2 ? B[1][6 + c] : B[1][9 + c]));
            const double b2 = I == 0 ? B[2][c] : (I == 1 ? B[2][3 + c] : (I == 2 ? B[2][6 + c] : B[2][9 + c]));
            E[0][c] = d11 * b0 + d12 * b1;
            E[1][c] = d12 * b0 + d11 * b1;
            E[2][c] = d33 * b2;
        }
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 12; c++) acc[r][c] += E[0][r] * B[0][c] + E[1][r] * B[1][c] + E[2][r] * B[2][c];
    }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 12; c++) {
            const double w0 = bench_w(r, c), w1 = bench_w(3 + r, c), w2 = bench_w(6 + r, c), w3 = bench_w(9 + r, c);
            s += acc[r][c] * (I == 0 ? w0 : (I == 1 ? w1 : (I == 2 ? w2 : w3)));
        }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (I == 0) out[e] = s;
}

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) k_contract_dmma(int64_t n_elem, double d11, double d12, double d33, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int am = lane >> 2, ak = lane & 3;   // A fragment: row am, k ak;  B fragment: k ak, column am
    for (int64_t e = warp; e < n_elem; e += n_warps) {
        const double base = 1.0 + 1e-6 * (double)(e & 1023);
        double c[2][2][2];
#pragma unroll
        for (int i = 0; i < 8; i++) (&c[0][0][0])[i] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 3; ks++) {
            const int k = 4 * ks + ak, gp = k / 3, r = k - 3 * gp;   // K index = (Gauss point, strain row)
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int m = 8 * mi + am;                           // row of Kp = column of B
                double a = 0.0;
                if (m < 12) {                                         // A[m][k] = (Dp B_gp)[r][m]
                    const double b0 = bench_b(base, gp, 0, m), b1 = bench_b(base, gp, 1, m), b2 = bench_b(base, gp, 2, m);
                    a = r == 0 ? d11 * b0 + d12 * b1 : (r == 1 ? d12 * b0 + d11 * b1 : d33 * b2);
                }
#pragma unroll
                for (int ni = 0; ni < 2; ni++) {
                    const int n = 8 * ni + am;
                    const double b = n < 12 ? bench_b(base, gp, r, n) : 0.0;
                    dmma_m8n8k4(c[mi][ni][0], c[mi][ni][1], a, b);
                }
            }
        }
        double s = 0.0;
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int m = 8 * mi + (lane >> 2), n = 8 * ni + 2 * (lane & 3) + h;
                    if (m < 12 && n < 12) s += c[mi][ni][h] * bench_w(m, n);
                }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[e] = s;
    }
}

// 8 independent accumulator pairs per warp: pure DMMA issue
__global__ void __launch_bounds__(256) k_dmma_peak(double *out, double a, double b, int iters)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
#pragma unroll
            for (int i = 0; i < 8; i++) dmma_m8n8k4(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

}  // namespace fs

extern "C" int fs_bench_contraction(fs_context *c, int64_t n_elem, int reps, double out[6])
{
    using namespace fs;
    if (!c || !out || n_elem <= 0 || reps <= 0) return FS_ERR_ARG;
    FS_CUDA(c, cudaSetDevice(c->device));
    DevBuf<double> ra, rb;
    FS_CUDA(c, ra.alloc(n_elem));
    FS_CUDA(c, rb.alloc(n_elem));
    const double d11 = 1.1e5, d12 = 0.33e5, d33 = 0.385e5;
    const unsigned ga = (unsigned)((4 * n_elem + 127) / 128), gb = (unsigned)(c->sm_count * 16);
    float ms[2] = {0.f, 0.f};
    for (int v = 0; v < 2; v++) {
        for (int rep = -2; rep < reps; rep++) {  // two warm-up launches
            if (rep == 0) FS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
            if (v == 0) k_contract_dfma<<<ga, 128, 0, c->stream>>>(n_elem, d11, d12, d33, ra.p);
            else k_contract_dmma<<<gb, 128, 0, c->stream>>>(n_elem, d11, d12, d33, rb.p);
        }
        FS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        FS_CUDA(c, cudaGetLastError());
        FS_CUDA(c, cudaEventElapsedTime(&ms[v], c->ev0, c->ev1));
        ms[v] /= reps;
    }
    std::vector<double> ha(n_elem), hb(n_elem);
    FS_CUDA(c, cudaMemcpy(ha.data(), ra.p, sizeof(double) * n_elem, cudaMemcpyDeviceToHost));
    FS_CUDA(c, cudaMemcpy(hb.data(), rb.p, sizeof(double) * n_elem, cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (int64_t i = 0; i < n_elem; i++) worst = std::max(worst, std::fabs(ha[i] - hb[i]) / std::fabs(ha[i]));
    // pure DMMA issue rate
    const int iters = 256, blocks = c->sm_count * 8;
    k_dmma_peak<<<blocks, 256, 0, c->stream>>>((double *)c->d_state.p, 0.999999, 1e-9, 4);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        FS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        k_dmma_peak<<<blocks, 256, 0, c->stream>>>((double *)c->d_state.p, 0.999999, 1e-9, iters);
        FS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        float t = 0.f;
        FS_CUDA(c, cudaEventElapsedTime(&t, c->ev0, c->ev1));
        best = t < best ? t : best;
    }
    FS_CUDA(c, cudaGetLastError());
    const double useful = 2.0 * 12 * 12 * 12 * (double)n_elem;  // flops of the contraction itself
    out[0] = ms[0];
    out[1] = ms[1];
    out[2] = worst;
    out[3] = 2.0 * 256.0 * 8 * 8 * (double)iters * (256 / 32) * blocks / (best * 1e-3) / 1e12;  // DMMA peak, TFLOP/s
    out[4] = useful / (ms[0] * 1e-3) / 1e12;
    out[5] = useful / (ms[1] * 1e-3) / 1e12;
    return FS_OK;
}

extern "C" int fs_bench_fp64_peak(fs_context *c, double *tflops)
{
    using namespace fs;
    if (!c || !tflops) return FS_ERR_ARG;
    FS_CUDA(c, cudaSetDevice(c->device));
    const int iters = 256, blocks = c->sm_count * 8;
    k_fp64_peak<<<blocks, 256, 0, c->stream>>>((double *)c->d_state.p, 0.999999, 1e-9, 8);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        FS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        k_fp64_peak<<<blocks, 256, 0, c->stream>>>((double *)c->d_state.p, 0.999999, 1e-9, iters);
        FS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        FS_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        best = ms < best ? ms : best;
    }
    FS_CUDA(c, cudaGetLastError());
    const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return FS_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel-to-kernel latency inside a CUDA graph, with and without programmatic dependent launch: the coarse levels of
// the multilevel cycle and the three kernels of a CG iteration are chains of short dependent kernels
// ---------------------------------------------------------------------------------------------
namespace fs {
template <bool PDL>
__global__ void __launch_bounds__(128) k_chain_link(double *v, int n)
{
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next link may be scheduled behind this grid
        asm volatile("griddepcontrol.wait;" ::: "memory");                // ... and this one touches memory only after its predecessor is done
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i] * 1.0000001 + 1.0;
}
}  // namespace fs

extern "C" int fs_bench_launch_chain(fs_context *c, int links, int n, int reps, double out_us[2])
{
    using namespace fs;
    if (!c || !out_us || links <= 0 || n <= 0 || reps <= 0) return FS_ERR_ARG;
    FS_CUDA(c, cudaSetDevice(c->device));
    DevBuf<double> v;
    FS_CUDA(c, v.alloc((size_t)n));
    FS_CUDA(c, cudaMemsetAsync(v.p, 0, sizeof(double) * n, c->stream));
    const unsigned grid = (unsigned)((n + 127) / 128);
    for (int variant = 0; variant < 2; variant++) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        FS_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < links; k++) {
            if (variant == 0) k_chain_link<false><<<grid, 128, 0, c->stream>>>(v.p, n);
            else {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(128);
                cfg.stream = c->stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, k_chain_link<true>, v.p, n);
            }
        }
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        FS_CUDA(c, ce);
        FS_CUDA(c, cudaGraphInstantiate(&exec, graph, 0));
        FS_CUDA(c, cudaGraphLaunch(exec, c->stream));
        FS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        for (int r = 0; r < reps; r++) FS_CUDA(c, cudaGraphLaunch(exec, c->stream));
        FS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        FS_CUDA(c, cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        FS_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        out_us[variant] = 1e3 * ms / ((double)reps * links);
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
    }
    return FS_OK;
}

