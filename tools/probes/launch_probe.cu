// Microbenchmark: period of back-to-back dependent kernels vs kernel duration, for plain stream launches,
// CUDA-graph replay and programmatic dependent launch (PDL). Build: nvcc -arch=sm_100a -O3 -o launch_probe launch_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void spin_kernel(long long cycles, int use_pdl, float* sink) {
    if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (use_pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (sink && threadIdx.x == 0 && blockIdx.x == 0 && cycles < 0) *sink = 1.f;
}

static float run_stream(cudaStream_t st, int n, int blocks, int threads, long long cycles, int smem, bool pdl) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < n; ++i) {
        if (!pdl) {
            spin_kernel<<<blocks, threads, smem, st>>>(cycles, 0, nullptr);
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, spin_kernel, cycles, 1, (float*)nullptr);
        }
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / n;
}

static float run_graph(cudaStream_t st, int n, int blocks, int threads, long long cycles, int smem, bool pdl) {
    cudaStream_t cs; cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed);
    for (int i = 0; i < n; ++i) {
        if (!pdl) spin_kernel<<<blocks, threads, smem, cs>>>(cycles, 0, nullptr);
        else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = cs;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, spin_kernel, cycles, 1, (float*)nullptr);
        }
    }
    if (cudaStreamEndCapture(cs, &g) != cudaSuccess) { printf("capture failed\n"); return -1; }
    if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(cs);
    return ms * 1000.f / n;
}

int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock (attr) %d kHz\n", clk);
    // warm up clocks
    for (int i = 0; i < 200; ++i) spin_kernel<<<148, 256, 0, st>>>(200000, 0, nullptr);
    cudaStreamSynchronize(st);
    const double us_to_cycles = 1965.0;
    const int n = 200;
    printf("%8s %8s %6s | %10s %10s %10s %10s   (period per kernel, us)\n", "spin_us", "blocks", "smemKB", "stream", "graph", "stream+pdl", "graph+pdl");
    for (int smem : {0, 150 * 1024}) {
        for (int blocks : {100}) {
            for (double us : {0.0, 0.5, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 8.0}) {
                long long cyc = (long long)(us * us_to_cycles);
                float a = run_stream(st, n, blocks, 128, cyc, smem, false);
                float b = run_graph(st, n, blocks, 128, cyc, smem, false);
                float c = run_stream(st, n, blocks, 128, cyc, smem, true);
                float d = run_graph(st, n, blocks, 128, cyc, smem, true);
                printf("%8.1f %8d %6d | %10.2f %10.2f %10.2f %10.2f\n", us, blocks, smem / 1024, a, b, c, d);
            }
        }
    }
    return 0;
}
