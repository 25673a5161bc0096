// Developer probe: FP64 FMA / FP64 MMA throughput and latency on the target GPU (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: DFMA, CHAINS independent chains per thread; 1: m8n8k4; 2: m16n8k8; 3: m16n8k16; 4: mixed dfma+m8n8k4
template <int MODE, int CHAINS>
__global__ void probe(double *out, int iters, double seed)
{
    double acc[CHAINS][4];
    for (int i = 0; i < CHAINS; i++) for (int q = 0; q < 4; q++) acc[i][q] = seed * (i + q + threadIdx.x);
    double a8[8], b4[4];
    for (int i = 0; i < 8; i++) a8[i] = seed + i;
    for (int i = 0; i < 4; i++) b4[i] = seed - i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) acc[i][0] = fma(acc[i][0], a8[0], b4[0]);
            if (MODE == 1) dmma884(acc[i][0], acc[i][1], a8[0], b4[0]);
            if (MODE == 2) { double a4[4] = {a8[0], a8[1], a8[2], a8[3]}; double b2[2] = {b4[0], b4[1]}; dmma1688(acc[i], a4, b2); }
            if (MODE == 3) dmma16816(acc[i], a8, b4);
            if (MODE == 4) { dmma884(acc[i][0], acc[i][1], a8[0], b4[0]); acc[i][2] = fma(acc[i][2], a8[1], b4[1]); acc[i][3] = fma(acc[i][3], a8[2], b4[2]); }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; i++) for (int q = 0; q < 4; q++) s += acc[i][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = double(t1 - t0);
}

template <int MODE, int CHAINS>
void run(const char *name, int warps_per_sm, double flops_per_inst_warp, double *d_out)
{
    int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int threads = 32 * warps_per_sm;
    probe<MODE, CHAINS><<<148, threads>>>(d_out, 64, 1e-9);
    cudaEventRecord(e0);
    probe<MODE, CHAINS><<<148, threads>>>(d_out, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d_out + 148 * threads, 8, cudaMemcpyDeviceToHost);
    double inst_per_sm = double(iters) * CHAINS * warps_per_sm * (MODE == 4 ? 1 : 1);
    printf("%-10s warps/SM=%2d chains=%d : %8.1f cycles/iter  -> %6.2f cycles per warp-inst per SM, %7.2f TFLOP/s (%s)\n", name,
           warps_per_sm, CHAINS, cyc / iters, cyc / inst_per_sm,
           148.0 * inst_per_sm * flops_per_inst_warp / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double *d; cudaMalloc(&d, (148 * 1024 + 8) * 8);
    // latency: 1 warp, 1 chain
    run<0, 1>("dfma", 1, 64, d);
    run<1, 1>("m8n8k4", 1, 512, d);
    run<2, 1>("m16n8k8", 1, 2048, d);
    run<3, 1>("m16n8k16", 1, 4096, d);
    // throughput
    for (int w : {4, 8, 16, 32}) {
        run<0, 8>("dfma", w, 64, d);
        run<1, 8>("m8n8k4", w, 512, d);
        run<2, 8>("m16n8k8", w, 2048, d);
        run<3, 8>("m16n8k16", w, 4096, d);
        run<4, 4>("mix1+2", w, 512 + 128, d);
    }
    return 0;
}
