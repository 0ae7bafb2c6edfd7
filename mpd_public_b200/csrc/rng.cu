// Standard-normal noise for the reverse loop, drawn on the device in ONE launch for all its steps, reproducing bit for bit
// what the reference's eager calls would have drawn from torch's CUDA generator:
//   x = torch.randn(shape, device=device)            diffusion_model_base.py:165
//   noise = torch.randn_like(x)   (every step)       sample_functions.py:51
// i.e. n_draws consecutive `normal_()` calls on contiguous fp32 tensors of `numel` elements.
//
// What such a call does (ATen/native/cuda/DistributionTemplates.h, restated from its published behaviour, not copied):
//   * launch geometry: 256 threads, grid = min(#SM * (max threads per SM / 256), ceil(numel / 256)); S = 256 * grid threads;
//   * thread idx owns Philox4x32-10 subsequence idx of the generator's (seed, offset); its j-th engine call yields four
//     normals (Box-Muller on the four 32-bit outputs, curand_normal4) that go to elements idx + (4 j + ii) S, ii = 0..3;
//   * the generator's offset then advances by 4 * ceil(numel / (4 S)) (four 32-bit outputs per engine call).
// Philox makes every value addressable: element li of draw k = component ii of philox(counter = (offset_k / 4 + j, idx),
// key = seed), so one kernel fills all draws. The host advances the torch generator by exactly what the eager calls would
// have consumed (Python binding), so later draws from the same generator are unchanged, too.
#include "common.cuh"
#include "internal.h"

namespace mpdb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        if (r < 9) { k.x += W0; k.y += W1; }
    }
    return c;
}

// curand's Box-Muller on two 32-bit outputs (curand_normal.h, device branch): same expressions, same intrinsics
__device__ __forceinline__ float2 box_muller(unsigned x, unsigned y) {
    constexpr float k2pow32_inv = 2.3283064e-10f, k2pow32_inv_2pi = 2.3283064e-10f * 6.2831855f;
    const float u = x * k2pow32_inv + (k2pow32_inv / 2);
    const float v = y * k2pow32_inv_2pi + (k2pow32_inv_2pi / 2);
    const float s = sqrtf(-2.0f * logf(u));
    float2 r;
    __sincosf(v, &r.x, &r.y);
    r.x *= s;
    r.y *= s;
    return r;
}

// state: device int64[2] = {seed, offset of the first draw}; draw k uses offset + k * offset_stride
__global__ void __launch_bounds__(256) normal_fill_kernel(float* __restrict__ out, long long numel, int n_draws,
                                                          const long long* __restrict__ state, long long offset_stride,
                                                          unsigned S, unsigned J) {
    const unsigned long long seed = (unsigned long long)state[0];
    const unsigned long long offset0 = (unsigned long long)state[1];
    const unsigned long long total = (unsigned long long)S * J * (unsigned)n_draws;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned idx = (unsigned)(w % S);
        const unsigned long long rest = w / S;
        const unsigned j = (unsigned)(rest % J);
        const int k = (int)(rest / J);
        const unsigned long long ctr_lo = (offset0 + (unsigned long long)k * (unsigned long long)offset_stride) / 4ull + j;
        const uint4 c = make_uint4((unsigned)ctr_lo, (unsigned)(ctr_lo >> 32), idx, 0u);
        const uint4 r = philox4x32_10(c, make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
        const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
        const float v[4] = {a.x, a.y, b.x, b.y};
        float* o = out + (long long)k * numel;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const long long li = (long long)idx + ((long long)4 * j + ii) * (long long)S;
            if (li < numel) o[li] = v[ii];
        }
    }
}

static int normal_geometry(long long numel, int device, unsigned* S, unsigned* J) {
    int sms = 0, max_threads = 0;
    MPDB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    MPDB_CHECK_CUDA(cudaDeviceGetAttribute(&max_threads, cudaDevAttrMaxThreadsPerMultiProcessor, device));
    unsigned long long grid = (unsigned long long)((numel + 255) / 256);
    const unsigned long long cap = (unsigned long long)sms * (unsigned long long)(max_threads / 256);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    *S = (unsigned)(256ull * grid);
    *J = (unsigned)((numel - 1) / (4ll * (long long)*S) + 1);
    return 0;
}

}  // namespace mpdb

using namespace mpdb;

extern "C" int64_t mpdb_normal_offset_increment(int64_t numel, int device) {
    if (numel <= 0) return 0;
    unsigned S = 0, J = 0;
    if (normal_geometry(numel, device, &S, &J)) return -1;
    return 4ll * (long long)J;
}

extern "C" int mpdb_normal_fill(float* out, int64_t numel, int32_t n_draws, const int64_t* state_dev, int device, void* stream) {
    MPDB_REQUIRE(out && state_dev && numel > 0 && n_draws > 0, "mpdb_normal_fill: bad argument");
    MPDB_ENTER_DEVICE(device);
    unsigned S = 0, J = 0;
    if (normal_geometry(numel, device, &S, &J)) return 1;
    const unsigned long long total = (unsigned long long)S * J * (unsigned)n_draws;
    unsigned long long blocks = (total + 255) / 256;
    if (blocks > 148ull * 16ull) blocks = 148ull * 16ull;
    normal_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, (long long)numel, n_draws, reinterpret_cast<const long long*>(state_dev),
                                                                           4ll * (long long)J, S, J);
    MPDB_LAUNCH_CHECK();
    return 0;
}
