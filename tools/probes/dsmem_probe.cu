// Probe: pushing a tile to the 7 peers of an 8-CTA cluster: (a) cp.async.bulk shared::cta -> shared::cluster with
// complete_tx on the peer's mbarrier, (b) per-thread st.shared::cluster.v4 + fence + mbarrier arrive (512 threads),
// (c) per-thread st.async.v4 with complete_tx on the peer's mbarrier (no release fence, no remote arrive), (d) = (b) without
// the producer-side proxy fence (what the cluster kernel does: the consumer fences).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t cta) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(8, 1, 1) dsmem_probe(int bytes, int mode, int iters, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = (uint64_t*)sm;               // one barrier
    unsigned char* src = sm + 128;               // my tile
    unsigned char* dst = sm + 128 + 16384;       // 8 slots of 16 KB
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(smem_u32(bar), (mode == 0 || mode == 2) ? 1 : 7 * 16 + 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 16384 / 4; i += blockDim.x) ((uint32_t*)src)[i] = i + rank;
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    cluster_sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(7 * bytes) : "memory");
                for (uint32_t p = 1; p < 8; ++p) {
                    const uint32_t peer = (rank + p) & 7;
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(smem_u32(dst + rank * 16384), peer)),
                                 "r"(smem_u32(src)), "r"(bytes), "r"(mapa(smem_u32(bar), peer)) : "memory");
                }
            }
        } else if (mode == 2) {
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(7 * bytes) : "memory");
            for (int o = tid * 16; o < bytes; o += blockDim.x * 16) {
                uint4 v = *reinterpret_cast<uint4*>(src + o);
                for (uint32_t p = 1; p < 8; ++p) {
                    const uint32_t peer = (rank + p) & 7;
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(mapa(smem_u32(dst + rank * 16384 + o), peer)),
                                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mapa(smem_u32(bar), peer)) : "memory");
                }
            }
        } else {
            // 512 threads: each stores its 16-byte pieces to the 7 peers, fence, one arrive per warp per peer
            for (int o = tid * 16; o < bytes; o += blockDim.x * 16) {
                uint4 v = *reinterpret_cast<uint4*>(src + o);
                for (uint32_t p = 1; p < 8; ++p) {
                    const uint32_t peer = (rank + p) & 7;
                    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(mapa(smem_u32(dst + rank * 16384 + o), peer)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                }
            }
            if (mode == 1) asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
            const int lane = tid & 31;
            if (lane >= 1 && lane < 8) {
                const uint32_t peer = (rank + lane) & 7;
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(smem_u32(bar), peer)) : "memory");
            }
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(smem_u32(bar), it & 1);
        cluster_sync();  // keep iterations separated (its cost is included in both modes)
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    cluster_sync();
}

int main() {
    long long* out; cudaMalloc(&out, 1024);
    cudaFuncSetAttribute(dsmem_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("mode(0=bulk push,1=st.shared::cluster+proxy fence,2=st.async complete_tx,3=st.shared::cluster) bytes | cycles per round (incl. one cluster barrier ~400)\n");
    for (int mode = 0; mode < 4; ++mode)
        for (int bytes : {1024, 4096, 8192, 12288, 16384}) {
            const int iters = 50;
            dsmem_probe<<<8 * 13, 512, 128 + 16384 * 9>>>(bytes, mode, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d bytes %d: error %s\n", mode, bytes, cudaGetErrorString(e)); return 1; }
            long long h[104]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < 104; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%d %6d | %8.0f   (%.1f B/cyc pushed per SM)\n", mode, bytes, (double)mx / iters, 7.0 * bytes / ((double)mx / iters));
        }
    return 0;
}
