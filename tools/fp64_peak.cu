// tools/fp64_peak.cu — measures B200's FP64 vector pipe (the roofline that binds WENO-5 in Float64; MEASURED_PEAKS.json has
// no FP64 figure).  (a) independent DFMA chains; (b) the same with one integer instruction issued per DFMA (does non-FP64 work
// co-issue for free?); (c) dependent DFMA latency.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MIXED>
__global__ void dfma_kernel(double* out, int iters, double a, double b, int* iout) {
    double x[ILP];
    int n[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-3 + i; n[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            x[i] = fma(x[i], a, b);
            if (MIXED) n[i] = n[i] * 3 + it;      // one IMAD per DFMA
        }
    }
    double s = 0;
    int m = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s += x[i]; m += n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (MIXED) iout[blockIdx.x * blockDim.x + threadIdx.x] = m;
}

template <int ILP, int MIXED>
double run(int threads, int blocks_per_sm, int sms, int iters) {
    double* out; int* iout;
    int nb = blocks_per_sm * sms;
    cudaMalloc(&out, sizeof(double) * nb * threads);
    cudaMalloc(&iout, sizeof(int) * nb * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<ILP, MIXED><<<nb, threads>>>(out, iters, 1.0000001, 1e-9, iout);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma_kernel<ILP, MIXED><<<nb, threads>>>(out, iters, 1.0000001, 1e-9, iout);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(out); cudaFree(iout);
    return (double)nb * threads * iters * ILP / (ms * 1e-3);     // DFMA lane-ops per second
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double peak = 64.0 * p.multiProcessorCount * clk_khz * 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f, \"nominal_fp64_lane_ops_per_s\": %.4g,\n", p.name, p.multiProcessorCount, clk_khz / 1e3, peak);
    double r;
    r = run<8, 0>(256, 4, p.multiProcessorCount, 20000); printf(" \"dfma_ilp8_1024thr\": %.4g, \"frac\": %.3f,\n", r, r / peak);
    r = run<4, 0>(288, 3, p.multiProcessorCount, 20000); printf(" \"dfma_ilp4_864thr\": %.4g, \"frac\": %.3f,\n", r, r / peak);
    r = run<2, 0>(288, 3, p.multiProcessorCount, 20000); printf(" \"dfma_ilp2_864thr\": %.4g, \"frac\": %.3f,\n", r, r / peak);
    r = run<1, 0>(288, 3, p.multiProcessorCount, 20000); printf(" \"dfma_ilp1_864thr\": %.4g, \"frac\": %.3f,\n", r, r / peak);
    r = run<8, 1>(256, 4, p.multiProcessorCount, 20000); printf(" \"dfma_plus_imad_ilp8_1024thr\": %.4g, \"frac\": %.3f,\n", r, r / peak);
    r = run<2, 1>(288, 3, p.multiProcessorCount, 20000); printf(" \"dfma_plus_imad_ilp2_864thr\": %.4g, \"frac\": %.3f,\n", r, r / peak);
    r = run<1, 0>(32, 1, p.multiProcessorCount, 200000);
    printf(" \"dependent_dfma_latency_cycles\": %.2f}\n", 32.0 * p.multiProcessorCount * clk_khz * 1e3 / r);
    return 0;
}
