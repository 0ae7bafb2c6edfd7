// TemporalUnet kernels, fp32 SIMT path (exact fp32 arithmetic; the tcgen05 path for the wide layers
// lives in unet_tc.cu). One kernel = Conv1d/ConvTranspose1d (+bias) [+GroupNorm+Mish] [+time cond]
// [+residual identity / 1x1 conv]  — reference layers.py:276-355, 258-273.
//
// Tiling: a CTA owns S whole samples x NT output channels (NT a multiple of the GroupNorm group size),
// so GroupNorm statistics never leave the CTA. Each thread owns a 4 (positions) x 4 (channels) register
// tile; input channels stream through shared memory in chunks of KC with a 2-stage cp.async pipeline,
// the k-tap window of a chunk is staged once and reused by all taps.
#include <cuda_bf16.h>

#include "common.cuh"
#include "internal.h"
#include "tc_common.cuh"

namespace mpdb {

constexpr int KC = 16;
constexpr int TP = 4;
constexpr int TC = 4;

template <int MODE>
struct ModeTraits;
template <>
struct ModeTraits<MODE_CONV5> { static constexpr int NTAPS = 5, XV = 8; };
template <>
struct ModeTraits<MODE_CONV1> { static constexpr int NTAPS = 1, XV = 4; };
template <>
struct ModeTraits<MODE_DOWN> { static constexpr int NTAPS = 3, XV = 9; };
template <>
struct ModeTraits<MODE_UP> { static constexpr int NTAPS = 4, XV = 4; };

struct TileCtx {
    int tid, nthreads;
    int s;       // sample within the CTA
    int l0;      // first output position of this thread
    int ct;      // channel tile index within the CTA
    int b0;      // first sample of the CTA
    int co0;     // first output channel of the CTA
    int S, NT, B;
    int G;       // intra-CTA split-K: number of K groups (each group = one full set of output tiles)
    int g;       // this thread's K group
    int T0;      // threads per K group
    int logS;    // log2(S)
};

// Decomposition of a flat staging index into (row, col) with col in [0, ncols), advanced by a fixed stride
// without divisions inside the hot loop.
struct Walk2 {
    int row, col, drow, dcol, ncols;
    __device__ __forceinline__ void init(int idx0, int stride, int ncols_) {
        ncols = ncols_;
        row = idx0 / ncols; col = idx0 - row * ncols;
        drow = stride / ncols; dcol = stride - drow * ncols;
    }
    __device__ __forceinline__ void next() {
        col += dcol; row += drow;
        if (col >= ncols) { col -= ncols; ++row; }
    }
};

// Stages one chunk of KC*G input channels (all S samples, full padded rows) and the matching weights.
// (kept out of line: the kernel is latency-bound at small batch and instruction-cache footprint matters more
// than call overhead — ncu showed `stalled_no_instruction` as the top stall with everything inlined)
template <int NTAPS>
__device__ __noinline__ void stage_chunk(float* xs, float* ws, const ConvSrc& src, const float* __restrict__ w,
                                         int CO, int kbase, const TileCtx& c) {
    const int Lp = src.L + 2 * HALO;
    const int ctot = src.c0 + src.c1;
    const int KCS = KC * c.G;  // channels per stage
    if (src.blc) {
        // raw trajectory [B][L][C]: transpose into xs[kk][s*Lp + l + HALO], halo written as zeros
        const int n = KCS * c.S * Lp;
        for (int idx = c.tid; idx < n; idx += c.nthreads) {
            int kk = idx / (c.S * Lp);
            int rem = idx - kk * (c.S * Lp);
            int s = rem / Lp;
            int j = rem - s * Lp;
            int ch = kbase + kk;
            int b = c.b0 + s;
            int l = j - HALO;
            float v = 0.f;
            if (ch < ctot && b < c.B && l >= 0 && l < src.L) v = src.p0[((long long)b * src.L + l) * src.c0 + ch];
            xs[idx] = v;
        }
    } else {
        // rows = (kk, s) pairs, cols = 16-byte granules of one padded row
        const int Lp4 = Lp >> 2;
        const int nrows = KCS * c.S;
        Walk2 wk;
        wk.init(c.tid, c.nthreads, Lp4);
        // a stage never straddles the two concat sources (c0 % KCS == 0 is enforced on the host)
        const bool second = kbase >= src.c0;
        const float* base = second ? src.p1 : src.p0;
        const int csrc = second ? src.c1 : src.c0;
        const int chb = second ? kbase - src.c0 : kbase;
        for (; wk.row < nrows; wk.next()) {
            const int kk = wk.row >> c.logS;  // S is a power of two
            const int s = wk.row & (c.S - 1);
            const int b = c.b0 + s;
            float* dst = xs + wk.row * Lp + wk.col * 4;
            if (kbase + kk < ctot && b < c.B) {
                cp_async16(dst, base + ((long long)b * csrc + chb + kk) * Lp + wk.col * 4);
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    {
        // rows = (kk, tap) pairs, cols = 16-byte granules of the CTA's NT output channels
        const int NT4 = c.NT >> 2;
        const int nrows = KCS * NTAPS;
        Walk2 wk;
        wk.init(c.tid, c.nthreads, NT4);
        const float* wb = w + (long long)kbase * NTAPS * CO + c.co0;
        const int rows_valid = (ctot - kbase) * NTAPS;  // rows beyond the last real channel are zero-filled
        for (; wk.row < nrows; wk.next()) {
            float* dst = ws + wk.row * c.NT + wk.col * 4;
            if (wk.row < rows_valid) {
                cp_async16(dst, wb + (long long)wk.row * CO + wk.col * 4);
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

// acc[i][j] += sum over the source's channels and taps. i = position in the thread tile, j = channel.
// With G > 1 the K loop is split over G thread groups (each stage holds KC*G channels, group g consumes its
// KC); the partial sums are then added into group 0 in fixed order g = 1..G-1 (deterministic).
template <int MODE>
__device__ __forceinline__ void accumulate(float (&acc)[TP][TC], float* smem, int stage_floats, int xs_floats,
                                           const ConvSrc& src, const float* __restrict__ w, int CO, const TileCtx& c) {
    constexpr int NTAPS = ModeTraits<MODE>::NTAPS;
    constexpr int XV = ModeTraits<MODE>::XV;
    const int Lp = src.L + 2 * HALO;
    const int ctot = src.c0 + src.c1;
    const int KCS = KC * c.G;
    const int nchunks = (ctot + KCS - 1) / KCS;
    const int xs_stride = c.S * Lp;

    // first input column (in padded coordinates) read by this thread
    int xbase;
    if (MODE == MODE_CONV5) xbase = c.s * Lp + c.l0;              // l + tap + HALO - 2
    else if (MODE == MODE_CONV1) xbase = c.s * Lp + c.l0 + HALO;  // l + HALO
    else if (MODE == MODE_DOWN) xbase = c.s * Lp + 2 * c.l0 + 1;  // 2l + tap + HALO - 1
    else xbase = c.s * Lp + (c.l0 >> 1) + 1;                      // m0 - 1 + HALO

    stage_chunk<NTAPS>(smem, smem + xs_floats, src, w, CO, 0, c);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        float* cur = smem + (ch & 1) * stage_floats;
        if (ch + 1 < nchunks) {
            float* nxt = smem + ((ch + 1) & 1) * stage_floats;
            stage_chunk<NTAPS>(nxt, nxt + xs_floats, src, w, CO, (ch + 1) * KCS, c);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* xs = cur + c.g * KC * xs_stride;
        const float* ws = cur + xs_floats + c.g * KC * NTAPS * c.NT;
        if (ch * KCS + c.g * KC < ctot) {
#pragma unroll 4
        for (int kk = 0; kk < KC; ++kk) {
            float xv[XV];
            const float* xp = xs + kk * xs_stride + xbase;
            if (MODE == MODE_CONV5) {
                float4 a = *reinterpret_cast<const float4*>(xp);
                float4 b = *reinterpret_cast<const float4*>(xp + 4);
                xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
                xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < XV; ++i) xv[i] = xp[i];
            }
            const float* wp = ws + kk * NTAPS * c.NT + c.ct * TC;
            if (MODE == MODE_UP) {
                float4 w0 = *reinterpret_cast<const float4*>(wp);
                float4 w1 = *reinterpret_cast<const float4*>(wp + c.NT);
                float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * c.NT);
                float4 w3 = *reinterpret_cast<const float4*>(wp + 3 * c.NT);
                const float a0[4] = {w0.x, w0.y, w0.z, w0.w};
                const float a1[4] = {w1.x, w1.y, w1.z, w1.w};
                const float a2[4] = {w2.x, w2.y, w2.z, w2.w};
                const float a3[4] = {w3.x, w3.y, w3.z, w3.w};
                // ConvTranspose1d(k=4, s=2, p=1): out[2m] = W1 x[m] + W3 x[m-1]; out[2m+1] = W0 x[m+1] + W2 x[m]
                // xv = x[m0-1 .. m0+2]
#pragma unroll
                for (int j = 0; j < TC; ++j) {
                    acc[0][j] = fmaf(a1[j], xv[1], fmaf(a3[j], xv[0], acc[0][j]));
                    acc[1][j] = fmaf(a0[j], xv[2], fmaf(a2[j], xv[1], acc[1][j]));
                    acc[2][j] = fmaf(a1[j], xv[2], fmaf(a3[j], xv[1], acc[2][j]));
                    acc[3][j] = fmaf(a0[j], xv[3], fmaf(a2[j], xv[2], acc[3][j]));
                }
            } else {
#pragma unroll
                for (int tap = 0; tap < NTAPS; ++tap) {
                    float4 wv4 = *reinterpret_cast<const float4*>(wp + tap * c.NT);
                    const float wv[4] = {wv4.x, wv4.y, wv4.z, wv4.w};
#pragma unroll
                    for (int i = 0; i < TP; ++i) {
                        const float x = (MODE == MODE_DOWN) ? xv[2 * i + tap] : xv[i + tap];
#pragma unroll
                        for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(x, wv[j], acc[i][j]);
                    }
                }
            }
        }
        }
        __syncthreads();
    }
    if (c.G > 1) {
        // cross-group reduction through shared memory (the stage buffers are free after the last barrier)
        float* part = smem;  // [(G-1)][T0][16]
        const int t0 = c.tid - c.g * c.T0;
        if (c.g > 0) {
            float4* dst = reinterpret_cast<float4*>(part + ((size_t)(c.g - 1) * c.T0 + t0) * (TP * TC));
#pragma unroll
            for (int i = 0; i < TP; ++i) dst[i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        if (c.g == 0) {
            for (int gg = 1; gg < c.G; ++gg) {
                const float4* srcp = reinterpret_cast<const float4*>(part + ((size_t)(gg - 1) * c.T0 + t0) * (TP * TC));
#pragma unroll
                for (int i = 0; i < TP; ++i) {
                    float4 v = srcp[i];
                    acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
                }
            }
        }
        __syncthreads();
    }
}

template <int MODE>
__global__ void __launch_bounds__(512) conv_kernel(ConvArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NTAPS = ModeTraits<MODE>::NTAPS;
    pdl_wait();  // our inputs come from the previous kernel

    TileCtx c;
    c.tid = threadIdx.x;
    c.nthreads = blockDim.x;
    c.S = a.S;
    c.NT = a.NT;
    c.B = a.B;
    c.b0 = blockIdx.x * a.S;
    c.co0 = blockIdx.y * a.NT;
    const int n_pt = (a.S * a.L_out) / TP;
    const int n_ct = a.NT / TC;
    c.T0 = n_pt * n_ct;
    c.logS = 31 - __clz(a.S);
    c.G = a.G;
    c.g = c.tid / c.T0;
    const int warp = c.tid >> 5, lane = c.tid & 31;
    const int warp0 = warp - c.g * (c.T0 >> 5);  // warp index within the K group
    const int pt_blocks = n_pt >> 3;
    const int ptb = warp0 % pt_blocks, ctb = warp0 / pt_blocks;
    const bool lead = (c.g == 0);  // group 0 owns the epilogue (statistics, residual, stores)
    const int pt = ptb * 8 + (lane & 7);
    c.ct = ctb * 4 + (lane >> 3);
    c.s = (pt * TP) / a.L_out;
    c.l0 = (pt * TP) - c.s * a.L_out;

    // shared memory carve-up: 2 stages of {xs, ws}, then the reduction scratch
    const int Lp_in = a.in.L + 2 * HALO;
    int xs_floats = KC * a.G * a.S * Lp_in;
    int ws_floats = KC * a.G * NTAPS * a.NT;
    if (a.res_w != nullptr) {
        int xr = KC * a.G * a.S * (a.res.L + 2 * HALO);
        xs_floats = xs_floats > xr ? xs_floats : xr;
    }
    const int stage_floats = xs_floats + ws_floats;
    float* red = smem + 2 * stage_floats;          // [n_ct][n_pt]

    float acc[TP][TC];
#pragma unroll
    for (int i = 0; i < TP; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;

    accumulate<MODE>(acc, smem, stage_floats, xs_floats, a.in, a.w, a.CO, c);
    pdl_launch_dependents();  // main loop over: the next kernel may run its prologue under our epilogue

    const int b = c.b0 + c.s;
    const int co = c.co0 + c.ct * TC;
    {
        float4 bv = *reinterpret_cast<const float4*>(a.bias + co);
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < TP; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] += bb[j];
    }

    if (a.gamma != nullptr) {
        // GroupNorm over (gs channels x L_out positions) of one sample: two-pass (mean, then centred second
        // moment), every thread adds the per-thread partials of its statistic in the same fixed order.
        const int gpt = a.NT / a.gs;            // groups per CTA tile
        const int ptl = a.L_out / TP;           // position tiles per sample
        const int ctg = a.gs / TC;              // channel tiles per group
        const int ct_first = ((c.ct * TC) / a.gs) * ctg;
        const int pt_first = c.s * ptl;
        const float inv_n = 1.f / (float)(a.gs * a.L_out);
        (void)gpt;
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < TP; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) s1 += acc[i][j];
        if (lead) red[c.ct * n_pt + pt] = s1;
        __syncthreads();
        float tot = 0.f;
        for (int cc = 0; cc < ctg; ++cc)
            for (int pp = 0; pp < ptl; ++pp) tot += red[(ct_first + cc) * n_pt + pt_first + pp];
        const float mean = tot * inv_n;
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < TP; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) {
                float d = acc[i][j] - mean;
                s2 = fmaf(d, d, s2);
            }
        __syncthreads();  // everyone has consumed the first-pass partials
        if (lead) red[c.ct * n_pt + pt] = s2;
        __syncthreads();
        tot = 0.f;
        for (int cc = 0; cc < ctg; ++cc)
            for (int pp = 0; pp < ptl; ++pp) tot += red[(ct_first + cc) * n_pt + pt_first + pp];
        const float rstd = 1.0f / sqrtf(tot * inv_n + 1e-5f);
        float4 gv = *reinterpret_cast<const float4*>(a.gamma + co);
        float4 bev = *reinterpret_cast<const float4*>(a.beta + co);
        const float ga[4] = {gv.x * rstd, gv.y * rstd, gv.z * rstd, gv.w * rstd};
        const float be[4] = {bev.x, bev.y, bev.z, bev.w};
#pragma unroll
        for (int i = 0; i < TP; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = mishf((acc[i][j] - mean) * ga[j] + be[j]);
    }

    if (a.cond != nullptr && b < a.B) {
        int tt = a.t_dev ? (int)a.t_dev[b] : a.t_uniform;
        float4 cv = *reinterpret_cast<const float4*>(a.cond + (long long)tt * a.CO + co);
        const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int i = 0; i < TP; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] += cc[j];
    }

    if (a.res.p0 != nullptr) {
        if (a.res_w != nullptr) {
            float racc[TP][TC];
#pragma unroll
            for (int i = 0; i < TP; ++i)
#pragma unroll
                for (int j = 0; j < TC; ++j) racc[i][j] = 0.f;
            __syncthreads();
            accumulate<MODE_CONV1>(racc, smem, stage_floats, xs_floats, a.res, a.res_w, a.CO, c);
            float4 bv = *reinterpret_cast<const float4*>(a.res_bias + co);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < TP; ++i)
#pragma unroll
                for (int j = 0; j < TC; ++j) acc[i][j] += racc[i][j] + bb[j];
        } else if (b < a.B && lead) {
            const int Lp = a.L_out + 2 * HALO;
#pragma unroll
            for (int j = 0; j < TC; ++j) {
                const float* rp = a.res.p0 + ((long long)b * a.CO + co + j) * Lp + HALO + c.l0;
                float2 r0 = *reinterpret_cast<const float2*>(rp);
                float2 r1 = *reinterpret_cast<const float2*>(rp + 2);
                acc[0][j] += r0.x; acc[1][j] += r0.y; acc[2][j] += r1.x; acc[3][j] += r1.y;
            }
        }
    }

    if (b < a.B && lead) {
        const int Lp = a.L_out + 2 * HALO;
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            float* op = a.out + ((long long)b * a.CO + co + j) * Lp + HALO + c.l0;
            *reinterpret_cast<float2*>(op) = make_float2(acc[0][j], acc[1][j]);
            *reinterpret_cast<float2*>(op + 2) = make_float2(acc[2][j], acc[3][j]);
        }
        if (a.out_hi != nullptr) {
            // second copy in the tensor-core layout (fp16 hi / scaled-lo planes, unet_tc.cu): 4 channels = 8 bytes per row
            const int SPT = TC_RT / Lp;
            const int tile = b / SPT, sb = b - tile * SPT;
#pragma unroll
            for (int i = 0; i < TP; ++i) {
                const long long o = (((long long)tile * (a.CO / 8) + co / 8) * TC_RT + (sb * Lp + c.l0 + i + HALO)) * 8 + (co & 7);
                unsigned short h[4], l[4];
#pragma unroll
                for (int j = 0; j < TC; ++j) split_hl(acc[i][j], h[j], l[j]);  // the tensor-core path's operand split
                *reinterpret_cast<uint2*>(a.out_hi + o) = make_uint2(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16));
                *reinterpret_cast<uint2*>(a.out_lo + o) = make_uint2(l[0] | ((unsigned)l[1] << 16), l[2] | ((unsigned)l[3] << 16));
            }
        }
    }
}

static int ntaps_of(int mode) { return mode == MODE_CONV5 ? 5 : mode == MODE_CONV1 ? 1 : mode == MODE_DOWN ? 3 : 4; }

static size_t conv_smem_bytes(int mode, const ConvArgs& a) {
    const int ntaps = ntaps_of(mode);
    const int P = a.S * a.L_out;
    int xs_floats = KC * a.G * a.S * (a.in.L + 2 * HALO);
    if (a.res_w != nullptr) {
        int xr = KC * a.G * a.S * (a.res.L + 2 * HALO);
        xs_floats = xs_floats > xr ? xs_floats : xr;
    }
    const int ws_floats = KC * a.G * ntaps * a.NT;
    const int n_pt = P / TP, n_ct = a.NT / TC;
    const int n_stats = a.S * (a.gs > 0 ? a.NT / a.gs : 1);
    size_t fl = (size_t)(2 * (xs_floats + ws_floats) + n_pt * n_ct + n_stats + 4);
    size_t red = (size_t)(a.G - 1) * n_pt * n_ct * TP * TC;  // cross-group reduction reuses the stage buffers
    if (red > (size_t)2 * (xs_floats + ws_floats)) fl += red - (size_t)2 * (xs_floats + ws_floats);
    return fl * sizeof(float);
}

// Picks (S, NT, G) for a layer. G (intra-CTA split-K) depends ONLY on the layer's input width, never on the
// batch, so that the summation order of every output element — and therefore the result, bit for bit — is
// independent of the batch size and of how the batch is sharded. (S, NT) only decide which CTA computes what.
void choose_tile(int mode, ConvArgs* a) {
    const int cin = a->in.c0 + a->in.c1;
    const int gs = a->gamma ? a->gs : 4;
    a->G = cin >= 64 ? 4 : cin >= 32 ? 2 : 1;
    while (a->G > 1 && a->in.c1 > 0 && a->in.c0 % (KC * a->G)) a->G /= 2;  // a stage must not straddle concat sources
    while (a->G > 1 && a->res_w && a->res.c1 > 0 && a->res.c0 % (KC * a->G)) a->G /= 2;
    long long best_cost = -1;
    int bS = 0, bNT = 0;
    const int nts[3] = {64, 32, 16};
    for (int k = 0; k < 3; ++k) {
        int NT = nts[k];
        if (NT > a->CO || a->CO % NT || NT % gs) continue;
        for (int S = 1; S <= 32; S *= 2) {
            int P = S * a->L_out;
            if (P % 32) continue;
            int T0 = P * NT / 16;
            if (T0 < 32 || T0 * a->G > 512) continue;
            ConvArgs t = *a;
            t.S = S;
            t.NT = NT;
            if (conv_smem_bytes(mode, t) > 200 * 1024) continue;
            long long ctas = (long long)((a->B + S - 1) / S) * (a->CO / NT);
            long long waves = (ctas + 147) / 148;
            int threads = T0 * a->G;
            long long cost = waves * (long long)(P * NT) * 16 + (threads == 512 ? 0 : threads == 256 ? 2 : 6) +
                             (NT == 32 ? 0 : NT == 64 ? 1 : 3);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; bS = S; bNT = NT; }
        }
    }
    a->S = bS;
    a->NT = bNT;
}

int launch_conv(int mode, ConvArgs a, cudaStream_t stream) {
    MPDB_REQUIRE(a.S > 0 && a.NT > 0, "conv tile: no feasible tiling for this layer shape");
    const int P = a.S * a.L_out;
    MPDB_REQUIRE(P % 32 == 0 && a.NT % 16 == 0 && a.CO % a.NT == 0, "conv tile: unsupported S/NT");
    MPDB_REQUIRE(a.gamma == nullptr || (a.NT % a.gs == 0 && a.gs % TC == 0), "conv tile: GroupNorm group not tile-aligned");
    MPDB_REQUIRE(a.in.L % 4 == 0 && a.L_out % 4 == 0, "conv: lengths must be multiples of 4");
    MPDB_REQUIRE(a.S > 0 && a.NT > 0, "conv tile: no feasible tiling for this layer shape");
    MPDB_REQUIRE(a.G == 1 || a.G == 2 || a.G == 4, "conv tile: G must be 1, 2 or 4");
    const int threads = a.G * P * a.NT / 16;
    MPDB_REQUIRE(threads >= 32 && threads <= 512, "conv tile: thread count out of range");
    const size_t smem = conv_smem_bytes(mode, a);
    MPDB_REQUIRE(smem <= 227 * 1024, "conv tile: shared memory budget exceeded");
    dim3 grid((a.B + a.S - 1) / a.S, a.CO / a.NT);
#define MPDB_CONV_CASE(M)                                                                                      \
    case M: {                                                                                                  \
        static unsigned long long configured = 0ull;                                                                        \
        if (mpdb::first_use_on_device(configured)) {                                                                                     \
            MPDB_CHECK_CUDA(cudaFuncSetAttribute(conv_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                                 227 * 1024));                                                 \
        }                                                                                                      \
        MPDB_CHECK_CUDA(launch_kernel(conv_kernel<M>, grid, dim3(threads), smem, stream, a));                  \
        break;                                                                                                 \
    }
    switch (mode) {
        MPDB_CONV_CASE(MODE_CONV5)
        MPDB_CONV_CASE(MODE_CONV1)
        MPDB_CONV_CASE(MODE_DOWN)
        MPDB_CONV_CASE(MODE_UP)
        default: MPDB_REQUIRE(false, "conv: bad mode");
    }
#undef MPDB_CONV_CASE
    MPDB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// final_conv.1 (1x1, C -> D) fused with the DDPM posterior mean (diffusion_model_base.py:126-138,149-150),
// optionally with the noise add + hard conditioning of ddpm_sample_fn / p_sample_loop
// (sample_functions.py:50-62, 5-8). One thread per (b, l).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) final_kernel(FinalArgs a) {
    // one thread per (b, l): consecutive threads read consecutive positions of every channel row of h (coalesced; the
    // one-thread-per-element mapping re-read each h value D times through 2-3 sectors per warp and took 57 us at B = 512,
    // H = 128), keep 8 output columns at a time in registers and apply the elementwise DDPM update. Per output the dot product
    // is the same fma chain over c = 0..C-1 plus the bias as before: bit-identical results.
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                 // [D][C]
    float* bsm = smem + a.D * a.C;     // [D]
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < a.D * a.C; i += blockDim.x) wsm[i] = a.w[i];  // constants: before the dependency wait
    for (int i = threadIdx.x; i < a.D; i += blockDim.x) bsm[i] = a.bias[i];
    pdl_wait();
    __syncthreads();
    const long long bl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_rows = (long long)a.B * a.L;
    bool viol = false;
    if (bl < n_rows) {
        const int l = (int)(bl % a.L);
        const int b = (int)(bl / a.L);
        const int Lp = a.L + 2 * HALO;
        const float* hp = a.h + (long long)b * a.C * Lp + HALO + l;
        const int tt = a.mode != 0 ? (a.t_dev ? (int)a.t_dev[b] : a.t_uniform) : 0;
        const float sd = a.mode == 2 ? a.stdv[tt] : 0.f;
        for (int d0 = 0; d0 < a.D; d0 += 8) {
            float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int c = 0; c < a.C; ++c) {
                const float hv = hp[(long long)c * Lp];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (d0 + j < a.D) e[j] = fmaf(wsm[(d0 + j) * a.C + c], hv, e[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d = d0 + j;
                if (d >= a.D) break;
                const long long idx = bl * a.D + d;
                float r = e[j] + bsm[d];
                if (a.mode != 0) {
                    const float xv = a.x[idx];
                    r = final_update_value(a, r, xv, tt);
                    if (a.mode == 2 || a.mode == 3) {
                        if (a.mode == 2) {
                            const float nz = (tt == 0) ? 0.f : a.noise[idx];
                            r = __fadd_rn(r, __fmul_rn(__fmul_rn(sd, nz), a.noise_std));
                        }
                        for (int k = 0; k < a.n_hc; ++k)  // later entries win, as in the reference's dict iteration
                            if (a.hc_rows[k] == l) r = a.hc_vals[((long long)k * a.B + b) * a.D + d];
                    } else {
                        viol |= (r > 1.0001f) || (r < -1.0001f);
                    }
                }
                a.out[idx] = r;
                if (a.out2) a.out2[(long long)b * a.out2_bstride + (long long)l * a.D + d] = r;
            }
        }
    }
    if (a.flag_out != nullptr) {
        if (__syncthreads_or(viol ? 1 : 0) && threadIdx.x == 0) atomicOr(a.flag_out, 1);
    }
}

int launch_final(const FinalArgs& a, cudaStream_t stream) {
    MPDB_REQUIRE(a.D <= MPDB_MAX_STATE_DIM, "state_dim too large");
    const long long n = (long long)a.B * a.L;
    const int threads = 128;
    const int blocks = (int)((n + threads - 1) / threads);
    const size_t smem = sizeof(float) * (size_t)(a.D * a.C + a.D);
    MPDB_CHECK_CUDA(launch_kernel(final_kernel, dim3(blocks), dim3(threads), smem, stream, a));
    MPDB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// weight repacking and time-conditioning tables (run once per load_state_dict)
// ---------------------------------------------------------------------------------------------------
__global__ void repack_conv_kernel(const float* __restrict__ src, float* __restrict__ dst, int CO, int CI, int K,
                                   int transposed) {
    // dst[ci][k][co] <- Conv1d weight [co][ci][k]  or  ConvTranspose1d weight [ci][co][k]
    long long n = (long long)CO * CI * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int co = (int)(i % CO);
        int k = (int)((i / CO) % K);
        int ci = (int)(i / ((long long)CO * K));
        long long s = transposed ? ((long long)ci * CO + co) * K + k : ((long long)co * CI + ci) * K + k;
        dst[i] = src[s];
    }
}

int launch_repack_conv(const float* src, float* dst, int CO, int CI, int K, int transposed, cudaStream_t stream) {
    long long n = (long long)CO * CI * K;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    repack_conv_kernel<<<blocks, 256, 0, stream>>>(src, dst, CO, CI, K, transposed);
    MPDB_LAUNCH_CHECK();
    return 0;
}

// TimeEncoder (layers.py:229-255): Sinusoidal(32) -> Linear(32,128) -> Mish -> Linear(128,32), then the
// Mish that opens every cond_mlp (layers.py:334-338). t is the same for the whole batch on the sampling
// path, so the result is a [T][32] table. One block per t.
__global__ void __launch_bounds__(128) time_table_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w3, const float* __restrict__ b3,
                                                         float* __restrict__ temb_mish) {
    __shared__ float emb[32];
    __shared__ float hid[128];
    const int t = blockIdx.x;
    const int j = threadIdx.x;
    if (j < 32) {
        const int half = 16;
        const float e = -(float)(9.210340371976184 / (half - 1));  // -(log(10000) / (half_dim - 1)) cast to fp32
        int k = j < half ? j : j - half;
        float freq = expf((float)k * e);
        float arg = (float)t * freq;
        emb[j] = j < half ? sinf(arg) : cosf(arg);
    }
    __syncthreads();
    {
        float acc = 0.f;
        for (int k = 0; k < 32; ++k) acc = fmaf(w1[j * 32 + k], emb[k], acc);
        hid[j] = mishf_ref(acc + b1[j]);
    }
    __syncthreads();
    if (j < 32) {
        float acc = 0.f;
        for (int k = 0; k < 128; ++k) acc = fmaf(w3[j * 128 + k], hid[k], acc);
        temb_mish[t * 32 + j] = mishf_ref(acc + b3[j]);
    }
}

int launch_time_tables(const float* w1, const float* b1, const float* w3, const float* b3, float* temb_mish, int T,
                       cudaStream_t stream) {
    time_table_kernel<<<T, 128, 0, stream>>>(w1, b1, w3, b3, temb_mish);
    MPDB_LAUNCH_CHECK();
    return 0;
}

__global__ void cond_table_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                  const float* __restrict__ temb_mish, float* __restrict__ table, int CO) {
    const int t = blockIdx.x;
    for (int co = threadIdx.x; co < CO; co += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < 32; ++k) acc = fmaf(w[co * 32 + k], temb_mish[t * 32 + k], acc);
        table[(long long)t * CO + co] = acc + b[co];
    }
}

int launch_cond_table(const float* w, const float* b, const float* temb_mish, float* table, int T, int CO,
                      cudaStream_t stream) {
    cond_table_kernel<<<T, 128, 0, stream>>>(w, b, temb_mish, table, CO);
    MPDB_LAUNCH_CHECK();
    return 0;
}

__global__ void cm_to_bcl_kernel(const float* __restrict__ cm, float* __restrict__ out, long long n, int L) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long row = i / L;
        int l = (int)(i - row * L);
        out[i] = cm[row * (L + 2 * HALO) + HALO + l];
    }
}

int launch_cm_to_bcl(const float* cm, float* out, int B, int C, int L, cudaStream_t stream) {
    long long n = (long long)B * C * L;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    cm_to_bcl_kernel<<<blocks, 256, 0, stream>>>(cm, out, n, L);
    MPDB_LAUNCH_CHECK();
    return 0;
}

__global__ void add_noise_kernel(float* __restrict__ x, const long long* __restrict__ t_dev,
                                 const float* __restrict__ stdv, const float* __restrict__ noise, float noise_std,
                                 long long n, int HD) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int b = (int)(i / HD);
        int tt = (int)t_dev[b];
        float nz = tt == 0 ? 0.f : noise[i];
        x[i] = __fadd_rn(x[i], __fmul_rn(__fmul_rn(stdv[tt], nz), noise_std));
    }
}

int launch_add_noise(float* x, const long long* t_dev, const float* stdv, const float* noise, float noise_std, int B,
                     int HD, cudaStream_t stream) {
    long long n = (long long)B * HD;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    add_noise_kernel<<<blocks, 256, 0, stream>>>(x, t_dev, stdv, noise, noise_std, n, HD);
    MPDB_LAUNCH_CHECK();
    return 0;
}

// x[:, row_k, :] = val_k (in place) and optional copy of the whole tensor to a chain slot
__global__ void copy_hc_kernel(float* __restrict__ x, float* __restrict__ out2, long long out2_bstride, int n_hc,
                               const int* __restrict__ hc_rows_dev, const float* __restrict__ hc_vals, int B, int L,
                               int D, int r0, int r1, int r2, int r3, int r4, int r5, int r6, int r7) {
    const int rows[MPDB_MAX_HARD_CONDS] = {r0, r1, r2, r3, r4, r5, r6, r7};
    long long n = (long long)B * L * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int d = (int)(i % D);
        int l = (int)((i / D) % L);
        int b = (int)(i / ((long long)D * L));
        float v = x[i];
        for (int k = 0; k < n_hc; ++k)
            if (rows[k] == l) v = hc_vals[((long long)k * B + b) * D + d];
        x[i] = v;
        if (out2) out2[(long long)b * out2_bstride + (long long)l * D + d] = v;
    }
}

int launch_copy_hc(float* x, float* out2, long long out2_bstride, int n_hc, const int* hc_rows, const float* hc_vals,
                   int B, int L, int D, cudaStream_t stream) {
    long long n = (long long)B * L * D;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    int r[MPDB_MAX_HARD_CONDS];
    for (int k = 0; k < MPDB_MAX_HARD_CONDS; ++k) r[k] = k < n_hc ? hc_rows[k] : -1;
    copy_hc_kernel<<<blocks, 256, 0, stream>>>(x, out2, out2_bstride, n_hc, nullptr, hc_vals, B, L, D, r[0], r[1],
                                               r[2], r[3], r[4], r[5], r[6], r[7]);
    MPDB_LAUNCH_CHECK();
    return 0;
}

// LimitsNormalizer.normalize (normalization.py:150-155) of rows [n_rows][d_in], zero-extended to d_out columns (the hard
// conditions are `cat(position, zeros)` before they are normalised, trajectories.py:214-237): out = 2 * ((v - min) / range) - 1
// with the reference's operation order, each operation rounded separately (torch runs them as separate kernels).
__global__ void limits_normalize_kernel(const float* __restrict__ x, long long n_rows, int d_in, const float* __restrict__ mins,
                                        const float* __restrict__ range, float* __restrict__ out, int d_out) {
    const long long n = n_rows * d_out;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % d_out);
        const long long r = i / d_out;
        const float v = d < d_in ? x[r * d_in + d] : 0.f;
        const float u = __fdiv_rn(__fsub_rn(v, mins[d]), range[d]);
        out[i] = __fsub_rn(__fmul_rn(2.f, u), 1.f);
    }
}

int launch_limits_normalize(const float* x, long long n_rows, int d_in, const float* mins, const float* range, float* out,
                            int d_out, cudaStream_t stream) {
    const long long n = n_rows * d_out;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    limits_normalize_kernel<<<blocks, 256, 0, stream>>>(x, n_rows, d_in, mins, range, out, d_out);
    MPDB_LAUNCH_CHECK();
    return 0;
}

}  // namespace mpdb
