// fs_bench.cu -- measurement helpers exported for bench.py: the FP64 FMA peak of the device, which is the
// roofline of the element kernels (MEASURED_PEAKS.json only carries HBM and bf16 figures; SURVEY.md 8d asks
// for an FMA micro-benchmark).
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

}  // namespace fs

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
