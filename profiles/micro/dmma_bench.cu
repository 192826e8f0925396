// Micro-benchmark: peak rate of the FP64 tensor instruction shapes on this GPU (no memory traffic):
// every warp keeps NACC independent accumulator fragments and issues mma.sync back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu && ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(512) k(double* out, int iters, double a0, double b0) {
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
void run(const char* name, double flop_per_mma, int threads) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, (size_t)sms * 4 * 1024 * 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<SHAPE, NACC><<<sms, threads>>>(out, 100, 1.0, 2.0);
    cudaEventRecord(e0);
    k<SHAPE, NACC><<<sms, threads>>>(out, iters, 1.0, 2.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)sms * (threads / 32) * iters * NACC;
    printf("%-12s acc %2d warps/SM %2d : %8.2f TFLOP/s  (%.3f ms)  %s\n", name, NACC, threads / 32, mmas * flop_per_mma / ms / 1e9, ms,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, 512);
    run<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, 256);
    run<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, 128);
    run<0, 4>("m8n8k4", 2.0 * 8 * 8 * 4, 512);
    run<1, 8>("m16n8k4", 2.0 * 16 * 8 * 4, 512);
    run<2, 8>("m16n8k8", 2.0 * 16 * 8 * 8, 512);
    run<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, 512);
    run<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, 128);
    run<3, 2>("m16n8k16", 2.0 * 16 * 8 * 16, 512);
    return 0;
}
