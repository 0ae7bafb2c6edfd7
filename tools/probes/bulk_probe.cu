// Probe: per-SM throughput / latency of cp.async.bulk (global -> shared, L2-resident source) as a function of copy size
// and copies in flight, for 1 / 104 / 148 CTAs, same or different source per CTA.  nvcc -arch=sm_100a -o bulk_probe bulk_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ int g_use_test;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = clock64();
    const int use_test = g_use_test;
    while (!done) {
        if (use_test)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void bulk_probe(const unsigned char* src, size_t region, int chunk, int depth, int n, int same, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = (uint64_t*)sm;
    unsigned char* buf = sm + 128;
    if (threadIdx.x == 0) {
        for (int d = 0; d < depth; ++d) mbar_init(smem_u32(bars + d), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const size_t base = same ? 0 : ((size_t)blockIdx.x * 1315423911ull) % (region / chunk);
        const long long t0 = clock64();
        for (int i = 0; i < n + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) mbar_wait(smem_u32(bars + s), ((i / depth) - 1) & 1);
            if (i < n) {
                const size_t off = ((base + i) % (region / chunk)) * (size_t)chunk;
                mbar_expect(smem_u32(bars + s), chunk);
                bulk(smem_u32(buf + (size_t)s * chunk), src + off, chunk, smem_u32(bars + s));
            }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const size_t region = 16u << 20;
    unsigned char* src;
    long long* out;
    cudaMalloc(&src, region);
    cudaMemset(src, 1, region);
    cudaMalloc(&out, 1024 * sizeof(long long));
    cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int chunks[] = {4096, 20480, 37376};
    const int depths[] = {1, 2, 3, 5, 8};
    const int grids[] = {1, 104};
    printf("wait grid same chunk depth | cycles/chunk (median CTA)  GB/s per SM  aggregate TB/s\n");
    for (int ut = 0; ut < 2; ++ut)
    for (int g : grids)
        for (int same = 1; same < 2; ++same)
            for (int c : chunks)
                for (int d : depths) {
                    if ((size_t)c * d + 128 > 200 * 1024) continue;
                    const int n = 128;
                    cudaMemcpyToSymbol(g_use_test, &ut, sizeof(int));
                    for (int rep = 0; rep < 2; ++rep)
                        bulk_probe<<<g, 32, (size_t)c * d + 128>>>(src, region, c, d, n, same, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    std::vector<long long> h(g);
                    cudaMemcpy(h.data(), out, g * sizeof(long long), cudaMemcpyDeviceToHost);
                    std::sort(h.begin(), h.end());
                    const double cyc = (double)h[g / 2] / n;
                    const double gbs = c / cyc * 1.965;
                    printf("%s %4d %4d %6d %5d | %10.0f %12.1f %10.2f\n", ut ? "test" : "try ", g, same, c, d, cyc, gbs, gbs * g / 1000);
                }
    return 0;
}
