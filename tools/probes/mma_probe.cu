// Probe: issue rate of tcgen05.mma kind::f16 (bf16, fp32 accumulate) in SS mode as a function of the shared-memory
// layout (no swizzle vs 32/64/128-byte swizzle), M and N. Operand contents are irrelevant for timing.
// nvcc -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity, int use_test) {
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        if (use_test)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// layout: 0 none, 6 = 32B, 4 = 64B, 2 = 128B swizzle
__device__ int g_fmt_f16;  // 1: a/b format F16 instead of BF16
__global__ void mma_probe(int M, int N, int layout, int n_mma, int use_test, long long* out, int n_issuers) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), n_issuers); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if ((threadIdx.x & 31) == 0 && warp < n_issuers) {
        const uint32_t a0 = smem_u32(sm) + warp * 1024, b0 = smem_u32(sm) + 96 * 1024 + warp * 2048;
        // K-major operands. no swizzle: core matrix 8 rows x 16 B, SBO = 128, LBO = rows*16.
        // swizzled: rows of 32/64/128 B, 8-row atoms; SBO = 8 * row bytes; successive K16 steps advance the start by 32 B.
        const uint32_t rowb = layout == 2 ? 128 : layout == 4 ? 64 : layout == 6 ? 32 : 16;
        uint64_t dA, dB;
        if (layout == 0) { dA = desc(a0, 132 * 16, 128, 0); dB = desc(b0, N * 16, 128, 0); }
        else { dA = desc(a0, 16, 8 * rowb, layout); dB = desc(b0, 16, 8 * rowb, layout); }
        const uint32_t id = g_fmt_f16 ? (idesc(M, N) & ~((1u << 7) | (1u << 10))) : idesc(M, N);
        const int ksteps = layout == 0 ? 1 : rowb / 32;  // K16 steps inside one swizzle row
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            // vary the start address like a real main loop: tap shift (rows) / k step
            uint64_t aofs, bofs;
            if (layout == 0) { aofs = (uint64_t)(((i % 5) * 16) >> 4); bofs = (uint64_t)(((i % 5) * 2048) >> 4); }
            else { aofs = (uint64_t)(((i % ksteps) * 32 + (i % 5) * rowb * 0) >> 4); bofs = aofs; }
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + warp * 64), "l"(dA + aofs), "l"(dB + bofs), "r"(id), "r"(i ? 1u : 0u) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait_spin(smem_u32(&bar), 0, use_test);
        const long long t2 = clock64();
        out[2 * warp + 0] = t1 - t0;
        out[2 * warp + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
    long long* out;
    cudaMalloc(&out, 128);
    cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int layouts[] = {0, 6, 4, 2};
    const char* names[] = {"none", "sw32", "sw64", "sw128"};
    const int Ms[] = {128, 64};
    const int Ns[] = {32, 64, 128, 256};
    printf("layout    M    N  wait | issue cyc/mma   total cyc/mma  (tensor floor M*N/256)\n");
    for (int use_test = 0; use_test < 2; ++use_test)
        for (int li = 0; li < 4; ++li)
            for (int M : Ms)
                for (int N : Ns) {
                    if (use_test && !(M == 128 && N == 64)) continue;
                    const int n = 400;
                    long long h[2];
                    for (int rep = 0; rep < 2; ++rep) mma_probe<<<1, 128, 160 * 1024>>>(M, N, layouts[li], n, use_test, out, 1);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("%s M=%d N=%d: error %s\n", names[li], M, N, cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                    printf("%-6s %4d %4d  %s | %10.1f %14.1f   (%d)\n", names[li], M, N, use_test ? "test" : "try ", (double)h[0] / n, (double)h[1] / n,
                           (M < 128 ? 128 : M) * N / 256);
                }
    for (int fmt = 0; fmt < 2; ++fmt) {
    cudaMemcpyToSymbol(g_fmt_f16, &fmt, sizeof(int));
    printf("operand format %s\n", fmt ? "F16" : "BF16");
    printf("issuers  N | issue cyc/mma (per issuer)  total cyc per mma (all issuers)\n");
    for (int ni = 1; ni <= 4; ++ni)
        for (int N : {32, 64}) {
            const int n = 400;
            long long h[8];
            for (int rep = 0; rep < 2; ++rep) mma_probe<<<1, 128, 160 * 1024>>>(128, N, 0, n, 0, out, ni);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("issuers=%d: error %s\n", ni, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int k = 0; k < ni; ++k) mx = h[2 * k + 1] > mx ? h[2 * k + 1] : mx;
            printf("%7d %3d | %10.1f %24.1f\n", ni, N, (double)h[0] / n, (double)mx / (n * ni));
        }
    }
    return 0;
}
