// Shared device-side pieces of the tensor-core path (unet_tc.cu per-layer kernels, unet_mega.cu whole-forward kernel):
// PTX wrappers for mbarrier / cp.async.bulk / tcgen05 / DSMEM, the fp16-split helpers and the GroupNorm+Mish epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "common.cuh"
#include "internal.h"

namespace mpdb {

constexpr int TC_STAGES = 5;  // 5 x 37,376 B in flight per SM: the wide layers are bound by L2->SM streaming latency
constexpr int TC_THREADS = 512;  // 16 warps: warp w reads TMEM lane quarter (w & 3), column group (w >> 2)
constexpr int TC_A_PLANE_BYTES = (TC_KCH / 8) * TC_RT * 16;  // 8448
constexpr int TC_B_TAP_BYTES = (TC_KCH / 8) * TC_NT * 16;    // 2048
constexpr int TC_STAGE_BYTES = 2 * TC_A_PLANE_BYTES + 2 * 5 * TC_B_TAP_BYTES;  // 37376
// GroupNorm scratch (gn_mish8): two moments x [32 row groups of 4][8 blocks] floats, column sums [2][12][8] doubles,
// {mean, rstd} [12][8][2] floats
constexpr int TC_GN_PART = 32 * 8;
constexpr int TC_GN_SCRATCH_BYTES = (2 * TC_GN_PART + 12 * 8 * 2) * (int)sizeof(float) + 2 * 12 * 8 * (int)sizeof(double);
constexpr int TC_TMEM_COLS = 128;  // main: [0,32) = hi*hi + lo*hi, [32,64) = hi*lo ; residual conv: [64,96), [96,128)

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (visible CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-convergent variants: the WHOLE warp executes the call with identical (warp-uniform) operands and one elected lane
// issues. Issued from a single-lane branch instead, the compiler cannot keep the descriptors in uniform registers and
// wraps every tcgen05.mma / cp.async.bulk in an ELECT + R2UR.BROADCAST loop (~80-100 cycles per instruction).
// Descriptors are given as (low word, high word): only the 14-bit start-address field in the low word changes between
// the MMAs of a K-chunk, so the per-instruction arithmetic is one 32-bit add.
__device__ __forceinline__ void tc_mma_bf16_elect32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                    uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte
// units): 8 rows of a core matrix are contiguous 16-byte rows; SBO = distance between 8-row groups; LBO =
// distance between the two 8-element k-groups of one K=16 MMA. Bits: [0,14) addr>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version = 1 (Blackwell), [61,64) layout type = 0.
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor for kind::f16: c_format F32 (bit 4), a/b format F16 (0 at bits [7,10) and [10,13); BF16 would be 1),
// K-major A and B, N >> 3 at bits [17,23), M >> 4 at bits [24,29).
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
#ifdef MPDB_EXP_BF16FMT
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#else
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#endif
}
// Three 8-column loads (the three partial accumulators of one output tile) with a single wait.
__device__ __forceinline__ void tc_ld8x3(uint32_t t0, uint32_t t1, uint32_t t2, float (&a)[8], float (&b)[8], float (&c)[8]) {
    uint32_t r[24];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t1));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]) : "r"(t2));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); c[i] = __uint_as_float(r[16 + i]); }
}

// Two 8-column loads with a single wait (precision-1 mode: the two K-group halves of one product).
__device__ __forceinline__ void tc_ld8x2(uint32_t t0, uint32_t t1, float (&a)[8], float (&b)[8]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t1));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); }
}

// Operand split x = hi + lo * 2^-11 with hi = fp16(x), lo = fp16((x - hi) * 2^11): 22 significant bits in two fp16 planes.
// The lo plane is stored scaled by 2^11 so that it stays in fp16's normal range whatever |x| (unscaled it would fall into
// the subnormals below |x| ~ 0.1 and lose its low bits); the products that contain one lo factor are accumulated in their
// own TMEM columns and multiplied by 2^-11 (exact) in the epilogue. hi*hi + (hi*lo + lo*hi) * 2^-11 reproduces the fp32
// product to ~2^-22 (the dropped lo*lo term), 2^5 better than a bf16 split at the same cost. |x| is clamped to the fp16
// range (activations and weights of this network are O(1)).
constexpr float TC_LO_SCALE = 2048.f, TC_LO_UNSCALE = 1.f / 2048.f;
__device__ __forceinline__ void split_hl(float x, unsigned short& hi, unsigned short& lo) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn((x - __half2float(h)) * TC_LO_SCALE);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// 8 fp32 values -> 8 "hi" and 8 scaled "lo" halves, packed for 16-byte stores. Same values as split_hl; cvt.rn.f16x2.f32
// converts two at a time.
__device__ __forceinline__ void pack_split8(const float (&v)[8], uint4& ph, uint4& pl) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = fminf(fmaxf(v[2 * k], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * k + 1], -65504.f), 65504.f);
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[k]) : "f"(b), "f"(a));  // {hi16: b, lo16: a}
        float h0, h1;
        asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(h0), "=f"(h1) : "r"(h[k]));
        const float r0 = (a - h0) * TC_LO_SCALE, r1 = (b - h1) * TC_LO_SCALE;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l[k]) : "f"(r1), "f"(r0));
    }
    ph = make_uint4(h[0], h[1], h[2], h[3]);
    pl = make_uint4(l[0], l[1], l[2], l[3]);
}

// precision-1 mode: fp16(x) only (the hi plane of pack_split8)
__device__ __forceinline__ uint4 pack_hi8(const float (&v)[8]) {
    uint32_t h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = fminf(fmaxf(v[2 * k], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * k + 1], -65504.f), 65504.f);
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[k]) : "f"(b), "f"(a));
    }
    return make_uint4(h[0], h[1], h[2], h[3]);
}
// One output tile's accumulator out of TMEM. Columns (relative to t): [0,32) hi*hi, [32,64) hi*lo, [64,96) lo*hi with the
// 22-bit split; in precision-1 mode [0,32) and [64,96) hold the two K-group halves of the single fp16 product.
__device__ __forceinline__ void tc_load_acc(uint32_t t, bool p1, float (&v)[8]) {
    if (p1) {
        float b[8];
        tc_ld8x2(t, t + 2 * TC_NT, v, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += b[j];
    } else {
        float v2[8], v3[8];
        tc_ld8x3(t, t + 2 * TC_NT, t + TC_NT, v, v3, v2);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v3[j] + v2[j], 1.f / 2048.f, v[j]);
    }
}

// Stores 8 consecutive output channels [c8, c8+8) of sample b at output position lo in the fp32 CM layout and (split
// into fp16 hi / scaled lo) in the TC layout of an activation with CO channels and L_out positions. (to, so) = the output's
// (row tile, sample slot within the tile) = (b / SPTo, b % SPTo) with SPTo = TC_RT / (L_out + 4): given by the caller, which
// knows them without a division when the layer keeps the length (the epilogue is instruction-issue bound).
__device__ __forceinline__ void tc_store_row_at(const TcConvArgs& a, const float (&v)[8], int b, int to, int so, int lo, int c8, int L_out) {
    const int Lpo = L_out + 4;
    if (a.out_cm != nullptr) {
        float* op = a.out_cm + ((size_t)b * a.CO + c8) * Lpo + 2 + lo;
#pragma unroll
        for (int j = 0; j < 8; ++j) op[(size_t)j * Lpo] = v[j];
    }
    if (a.out_hi != nullptr) {
        const size_t o = (((size_t)to * (a.CO / 8) + c8 / 8) * TC_RT + (so * Lpo + lo + 2)) * 8;
        if (a.prec == 1) {  // precision 1 reads the hi plane only
            *reinterpret_cast<uint4*>(a.out_hi + o) = pack_hi8(v);
        } else {
            uint4 ph, pl;
            pack_split8(v, ph, pl);  // same values as split_hl, two conversions per instruction
            *reinterpret_cast<uint4*>(a.out_hi + o) = ph;
            *reinterpret_cast<uint4*>(a.out_lo + o) = pl;
        }
    }
}
__device__ __forceinline__ void tc_store_row(const TcConvArgs& a, const float (&v)[8], int b, int lo, int c8, int L_out) {
    const int SPTo = TC_RT / (L_out + 4);
    const int to = b / SPTo;
    tc_store_row_at(a, v, b, to, b - to * SPTo, lo, c8, L_out);
}

// One-pass GroupNorm statistics + Mish on the 8 channels a thread owns (both tensor-core epilogues).
//   level 0: every thread forms sum and sum-of-squares of its two 4-channel blocks; the 4 rows of an aligned row group are
//            added with two xor-shuffles (rows = lanes) and parked: part[0][row/4][blk], part[1][row/4][blk]
//   level 1: thread t < 2*SPT*8 owns one (moment, sample, block) column and adds its L/4 group sums in DOUBLE precision
//            (fixed tree, two accumulators so the loads pipeline) -> cs[moment][sample][blk]
//   level 2: the BPG block sums of a group are combined in double: mean = S/n, var = Q/n - mean^2, and {mean, rstd} is
//            published; everyone reads its two groups' values after the last barrier. With at most 8 samples per tile
//            levels 1 and 2 are one shuffle-connected pass (two barriers in all), otherwise they are separated by a barrier.
// Accumulating the cross-row part in fp64 removes the cancellation of the one-pass formula; what remains is the fp32
// rounding of the 16-term row-group partials (~6e-8 * (1 + mean^2/var)), far below the fp16-split MMA error. Three short
// barriers, no redundant fp64 work. Deterministic and independent of how samples are tiled. BAR1: named barrier of the 512 epilogue
// threads (fused kernel, where a producer warp is not part of the epilogue) instead of __syncthreads.
// The pieces (the persistent per-layer kernel software-pipelines them across work items, unet_tc.cu):
// level 0: per-thread partials of the two 4-channel blocks, then the 4 rows of an aligned row group are added by two
// xor-shuffles (rows = lanes; samples start at multiples of 4 rows, so groups never straddle samples and the tree is the
// same wherever the sample sits in the tile)
__device__ __forceinline__ void gn_level0(const float (&v)[8], bool valid, int r, int cg, float* part) {
    float* part2 = part + TC_GN_PART;
    float sA = (v[0] + v[1]) + (v[2] + v[3]), sB = (v[4] + v[5]) + (v[6] + v[7]);
    float qA = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], v[3] * v[3])));
    float qB = fmaf(v[4], v[4], fmaf(v[5], v[5], fmaf(v[6], v[6], v[7] * v[7])));
    if (!valid) { sA = 0.f; sB = 0.f; qA = 0.f; qB = 0.f; }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        sA += __shfl_xor_sync(0xffffffffu, sA, o); sB += __shfl_xor_sync(0xffffffffu, sB, o);
        qA += __shfl_xor_sync(0xffffffffu, qA, o); qB += __shfl_xor_sync(0xffffffffu, qB, o);
    }
    if ((r & 3) == 0) {
        part[(r >> 2) * 8 + cg * 2 + 0] = sA;
        part[(r >> 2) * 8 + cg * 2 + 1] = sB;
        part2[(r >> 2) * 8 + cg * 2 + 0] = qA;
        part2[(r >> 2) * 8 + cg * 2 + 1] = qB;
    }
}
// levels 1 + 2 with at most 8 samples per tile, in one pass with no shared-memory round trip in between: JS threads per
// (sample, block, moment) column add the sample's L/4 row-group sums in double (thread j takes groups j, j+JS, ...); fixed
// xor-shuffle trees then combine the JS partial sums, bring the two moments together and add the BPG blocks of a GroupNorm
// group (lane bits, low to high: j | moment | block), and one lane per (sample, group) publishes {mean, rstd}. Every thread
// of the 16 warps calls it (whole warps take part in the shuffles).
template <int GS>
__device__ __forceinline__ void gn_stats_small(int tid, int SPT, int Lp, int L, float* part) {
    constexpr int BPG = GS / 4;
    float* part2 = part + TC_GN_PART;
    double* cs = reinterpret_cast<double*>(part + 2 * TC_GN_PART);  // [2][12][8]
    float* stat = reinterpret_cast<float*>(cs + 2 * 12 * 8);         // [12][8][2]
    constexpr int JS = BPG == 8 ? 2 : 4;
    const int j = tid % JS, m = (tid / JS) & 1, blk = (tid / (2 * JS)) & 7, ss = tid / (16 * JS);
    const bool on = ss < SPT;
    // 1 / n up front (an fp32 division is ~20 dependent instructions): it overlaps the shared-memory loads below instead of
    // sitting behind the reductions, on the path between the two barriers that all 16 warps wait on
    const double inv_n = (double)(1.0f / (float)(GS * L));  // L = 8 * 2^k, GS = 2^j: exact, and no fp64 division
    double a0 = 0.0;
    if (on) {
        // (eight independent loads + a pairwise tree per round was tried: slower — the fp32 -> fp64 conversions and double adds of
        // the padding terms cost more than the dependent chain, 178 -> 180 us per forward of the cluster kernel)
        const float* p = (m ? part2 : part) + (size_t)ss * (Lp >> 2) * 8 + blk;
        for (int g = j; g < (L >> 2); g += JS) a0 += (double)p[g * 8];
    }
#pragma unroll
    for (int o = 1; o < JS; o <<= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    const double other = __shfl_xor_sync(0xffffffffu, a0, JS);  // the other moment of the same (sample, block)
    double S = m ? other : a0, Q = m ? a0 : other;
#pragma unroll
    for (int o = 1; o < BPG; o <<= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o * 2 * JS);
        Q += __shfl_xor_sync(0xffffffffu, Q, o * 2 * JS);
    }
    if (on && j == 0 && m == 0 && (blk % BPG) == 0) {
        const int g = blk / BPG;
        const double mean = S * inv_n;
        const double var = fmax(Q * inv_n - mean * mean, 0.0);
        stat[(ss * 8 + g) * 2 + 0] = (float)mean;
        stat[(ss * 8 + g) * 2 + 1] = rsqrtf((float)var + 1e-5f);  // the cancellation-prone part is done; fp32 from here (MUFU.RSQ, 2^-22.9:
                                                                  // an IEEE sqrt + division is ~40 dependent instructions between the two barriers)
    }
}
// more than 8 samples per tile: level 1 = one thread per (moment, sample, block) column; a barrier; level 2 = one thread per
// (sample, group) finishes the statistics in double and publishes {mean, rstd} as floats
__device__ __forceinline__ void gn_stats_big1(int tid, int SPT, int Lp, int L, float* part) {
    float* part2 = part + TC_GN_PART;
    double* cs = reinterpret_cast<double*>(part + 2 * TC_GN_PART);
    if (tid < 2 * SPT * 8) {
        const int m = tid >= SPT * 8 ? 1 : 0;
        const int t2 = tid - m * SPT * 8;
        const int ss = t2 >> 3, blk = t2 & 7;
        const float* p = (m ? part2 : part) + (size_t)ss * (Lp >> 2) * 8 + blk;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
        for (int g = 0; g < (L >> 2); g += 2) {
            a0 += (double)p[(g + 0) * 8];
            a1 += (double)p[(g + 1) * 8];
        }
        cs[(m * 12 + ss) * 8 + blk] = a0 + a1;
    }
}
template <int GS>
__device__ __forceinline__ void gn_stats_big2(int tid, int SPT, int L, float* part) {
    constexpr int BPG = GS / 4, NG = TC_NT / GS;
    double* cs = reinterpret_cast<double*>(part + 2 * TC_GN_PART);
    float* stat = reinterpret_cast<float*>(cs + 2 * 12 * 8);
    if (tid < SPT * NG) {
        const int ss = tid / NG, g = tid - ss * NG;
        double S = 0.0, Q = 0.0;
#pragma unroll
        for (int k = 0; k < BPG; ++k) { S += cs[(0 * 12 + ss) * 8 + g * BPG + k]; Q += cs[(1 * 12 + ss) * 8 + g * BPG + k]; }
        const double inv_n = (double)(1.0f / (float)(GS * L));
        const double m = S * inv_n;
        const double var = fmax(Q * inv_n - m * m, 0.0);
        stat[(ss * 8 + g) * 2 + 0] = (float)m;
        stat[(ss * 8 + g) * 2 + 1] = rsqrtf((float)var + 1e-5f);
    }
}
// normalise + affine + Mish with the published statistics
template <int GS>
__device__ __forceinline__ void gn_apply(float (&v)[8], bool valid, int s, int cg, int SPT, const float* part, const float4& g0,
                                         const float4& g1, const float4& e0, const float4& e1) {
    const float* stat = reinterpret_cast<const float*>(reinterpret_cast<const double*>(part + 2 * TC_GN_PART) + 2 * 12 * 8);
    const int sc = s < SPT ? s : 0;
    const int gA = (cg * 8) / GS, gB = (cg * 8 + 4) / GS;
    const float mA = stat[(sc * 8 + gA) * 2], rA = stat[(sc * 8 + gA) * 2 + 1];
    const float mB = stat[(sc * 8 + gB) * 2], rB = stat[(sc * 8 + gB) * 2 + 1];
    if (!valid) return;  // rows outside the samples (halo, unused lanes): their values are never stored; skipping them halves the
                         // exp / reciprocal (MUFU) traffic of a tile that is half empty (one sample of L = 64 in 128 lanes)
    v[0] = (v[0] - mA) * (rA * g0.x) + e0.x; v[1] = (v[1] - mA) * (rA * g0.y) + e0.y;
    v[2] = (v[2] - mA) * (rA * g0.z) + e0.z; v[3] = (v[3] - mA) * (rA * g0.w) + e0.w;
    v[4] = (v[4] - mB) * (rB * g1.x) + e1.x; v[5] = (v[5] - mB) * (rB * g1.y) + e1.y;
    v[6] = (v[6] - mB) * (rB * g1.z) + e1.z; v[7] = (v[7] - mB) * (rB * g1.w) + e1.w;
    mishf_fast2(v[0], v[1]); mishf_fast2(v[2], v[3]); mishf_fast2(v[4], v[5]); mishf_fast2(v[6], v[7]);
}

template <int GS, bool BAR1>
__device__ __forceinline__ void gn_mish8(float (&v)[8], bool valid, int r, int s, int cg, int tid, int SPT, int Lp, int L,
                                         float* part, const float4& g0, const float4& g1, const float4& e0, const float4& e1,
                                         long long* dbg = nullptr) {
    gn_level0(v, valid, r, cg, part);
    if (BAR1) asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); else __syncthreads();
    if (dbg) dbg[8] = clock64();
    if (SPT <= 8) {
        gn_stats_small<GS>(tid, SPT, Lp, L, part);
        if (dbg) dbg[9] = clock64();
    } else {
        gn_stats_big1(tid, SPT, Lp, L, part);
        if (BAR1) asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); else __syncthreads();
        if (dbg) dbg[9] = clock64();
        gn_stats_big2<GS>(tid, SPT, L, part);
    }
    if (BAR1) asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); else __syncthreads();
    gn_apply<GS>(v, valid, s, cg, SPT, part, g0, g1, e0, e1);
}


__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Remote 16-byte store that reports its own completion: the destination CTA's mbarrier (shared::cluster address, same CTA
// as the data) receives complete_tx(16) when the bytes have landed, so the sender needs no release fence (MEMBAR.ALL.GPU waits
// for the acknowledgement of every outstanding remote store) and no remote arrive. tools/probes/dsmem_probe.cu: an 8-way
// exchange completes ~1000 cycles sooner than with st.shared::cluster + arrive.release.cluster, at any size.
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

}  // namespace mpdb
