// dfma_probe.cu — FP64 FMA issue rate on one SM partition as a function of independent chains per
// thread (ILP) and resident warps: tells how much instruction-level parallelism the fused pass
// needs to keep the FP64 pipe busy.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_probe tools/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void probe(double *out, double a, double b, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
/* the fused pass's operand pattern: every FMA reads three distinct register pairs (amplitude,
 * matrix entry, accumulator) instead of two loop constants */
template <int ILP>
__global__ void probe3(double *out, const double *in, int iters) {
    double acc[ILP], x[ILP], m[4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i, x[i] = in[(threadIdx.x + i) & 63];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = in[64 + i];
    for (int it = 0; it < iters; it += 2) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(x[i], m[i & 3], acc[i]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(x[(i + 1) % ILP], m[(i + 1) & 3], acc[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run3(int warps_per_sm, double *d, const double *in) {
    const int iters = 4096, nsm = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe3<ILP><<<nsm, warps_per_sm * 32>>>(d, in, 16);
    cudaEventRecord(e0);
    probe3<ILP><<<nsm, warps_per_sm * 32>>>(d, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double dfma_per_smsp = (double)iters * ILP * warps_per_sm / 4.0;
    printf("3-operand ILP %2d warps/SM %2d: %.2f cycles per warp-DFMA per SMSP\n", ILP, warps_per_sm, cycles / dfma_per_smsp);
}
template <int ILP>
void run(int warps_per_sm, double *d) {
    const int iters = 4096, nsm = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<ILP><<<nsm, warps_per_sm * 32>>>(d, 1.0000001, 1e-9, 16);
    cudaEventRecord(e0);
    probe<ILP><<<nsm, warps_per_sm * 32>>>(d, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double dfma_per_smsp = (double)iters * ILP * warps_per_sm / 4.0;
    printf("ILP %2d warps/SM %2d: %.2f cycles per warp-DFMA per SMSP (%.1f%% of 1 per 2 cycles), %.1f cycles between dependent DFMAs\n", ILP,
           warps_per_sm, cycles / dfma_per_smsp, 200. * dfma_per_smsp / cycles, cycles / iters);
}
int main() {
    double *d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    double *in; cudaMalloc(&in, 128 * sizeof(double)); cudaMemset(in, 0, 128 * sizeof(double));
    for (int w : {4, 8, 12, 16, 24, 32}) {
        run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<16>(w, d);
    }
    for (int w : {4, 8, 12, 16, 32}) {
        run3<8>(w, d, in); run3<16>(w, d, in);
    }
    return 0;
}
