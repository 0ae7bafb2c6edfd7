// Shared helpers of the mpdb200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace mpdb {

void set_error(const std::string& msg);
extern std::atomic<long long> g_launch_count;

#define MPDB_CHECK_CUDA(expr)                                                                       \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            mpdb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ +  \
                            ":" + std::to_string(__LINE__) + ")");                                  \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

#define MPDB_REQUIRE(cond, msg)                  \
    do {                                         \
        if (!(cond)) {                           \
            mpdb::set_error(std::string(msg));   \
            return 2;                            \
        }                                        \
    } while (0)

// Every extern "C" entry runs on its handle's device and restores the caller's current device on return (a call on a
// cuda:1 model must not change torch's current device behind the caller's back).
struct DeviceGuard {
    int prev = -1, target = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) : target(dev) {
        ok = cudaGetDevice(&prev) == cudaSuccess && (prev == dev || cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (ok && prev != target) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define MPDB_ENTER_DEVICE(dev)       \
    mpdb::DeviceGuard _mpdb_dg(dev); \
    MPDB_REQUIRE(_mpdb_dg.ok, "cudaSetDevice(" + std::to_string(dev) + ") failed")

// cudaFuncSetAttribute is per DEVICE: true the first time the calling site runs on the current device (one bit per ordinal).
inline bool first_use_on_device(unsigned long long& mask) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d > 63) return true;
    if ((mask >> d) & 1ull) return false;
    mask |= 1ull << d;
    return true;
}

#define MPDB_LAUNCH_CHECK()                      \
    do {                                         \
        mpdb::g_launch_count.fetch_add(1);       \
        MPDB_CHECK_CUDA(cudaGetLastError());     \
    } while (0)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Programmatic dependent launch (PDL): a kernel launched with the stream-serialization attribute may start while its
// predecessor is still running; it must not touch the predecessor's outputs before pdl_wait(). Both are no-ops for
// ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern bool g_use_pdl;     // MPDB_PDL=1: programmatic dependent launch for EVERY kernel (off by default, see engine.cu)
extern bool g_pdl_layers;  // MPDB_PDL_LAYERS (default 1): for the persistent per-layer tensor-core kernels only (engine.cu)
extern bool g_pdl_loop;    // MPDB_PDL_LOOP (default 1): cluster-kernel forwards and the single-wave guide launches between them

// Launch with (pdl = true) or without the PDL attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// Launch with the PDL attribute when it is enabled globally (falls back to a plain launch otherwise).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
    return launch_kernel_pdl(kernel, grid, block, smem, stream, g_use_pdl, args...);
}

// Same, with a thread-block-cluster shape (cluster_y CTAs along grid y).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         unsigned cluster_y, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = cluster_y;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_use_pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Mish(x) = x * tanh(softplus(x)) (aten: x * tanh(log1p(exp(x)))). With e = exp(x):
// tanh(log(1 + e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = n / (n + 2), n = e * (e + 2)  — one expf, one division,
// no cancellation for x << 0; for x > 20 tanh(softplus(x)) == 1 in fp32.
__device__ __forceinline__ float mishf(float x) {
    // branch-free: for x >= 20, e = exp(20), n = e(e+2) ~ 2.4e17 and n / (n + 2) == 1.0f exactly, so the clamp alone
    // reproduces "tanh(softplus(x)) == 1" and independent evaluations can interleave
    const float e = expf(fminf(x, 20.f));
    const float n = e * (e + 2.f);
    return x * __fdiv_rn(n, n + 2.f);
}
// fast variant for the tensor-core epilogue (ex2.approx + rcp.approx, ~1e-6 relative; the fp16-split MMA itself is
// ~1e-5): the epilogue is instruction-issue bound, so the ~40-instruction precise version would dominate it.
__device__ __forceinline__ float mishf_fast(float x) {
    const float e = __expf(fminf(x, 20.f));
    const float n = e * (e + 2.f);
    return x * __fdividef(n, n + 2.f);  // n + 2 <= 2.4e17 < 2^126: inside __fdividef's valid range
}
// Two at a time with ONE reciprocal: 1/(a b) gives 1/a = b/(a b) and 1/b = a/(a b). The tensor-core epilogues run 8 of these
// per thread on warps whose lane quarter ties them to one scheduler (TMEM access rule), so the quarter-rate MUFU pipe of two or
// three schedulers does all of it: 12 MUFU per thread instead of 16. (n + 2)^2 <= 5.6e34 stays inside fp32 and rcp.approx's range.
__device__ __forceinline__ void mishf_fast2(float& x0, float& x1) {
    const float e0 = __expf(fminf(x0, 20.f)), e1 = __expf(fminf(x1, 20.f));
    const float n0 = e0 * (e0 + 2.f), n1 = e1 * (e1 + 2.f);
    const float d0 = n0 + 2.f, d1 = n1 + 2.f;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
    x0 = x0 * (n0 * (r * d1));
    x1 = x1 * (n1 * (r * d0));
}
// reference-composition variant (used once per load for the time tables, where cost does not matter)
__device__ __forceinline__ float mishf_ref(float x) { return x * tanhf(log1pf(expf(x))); }

}  // namespace mpdb
