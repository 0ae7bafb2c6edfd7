// TemporalUnet k=5 convolutions on the 5th-generation tensor cores (tcgen05 / TMEM), fed by bulk-async
// (TMA engine) copies. One kernel = Conv1d(k5) + bias -> GroupNorm -> Mish [+ time cond] [+ residual identity |
// + fused 1x1 residual conv in a second TMEM accumulator]  — reference layers.py:276-355.
//
// Implicit GEMM, "positions on M":
//   D[r, n] = sum_tap sum_ci A[r + tap, ci] * W_tap[n, ci]
//   r  = padded row of a tile of SPT whole samples (each sample = 2 zero rows + L rows + 2 zero rows),
//        128 rows per CTA = the 128 TMEM lanes; rows that fall on a halo are computed and ignored;
//   n  = 32 output channels per CTA (a whole number of GroupNorm groups);
//   the k=5 window is NOT materialised: the activation tile is staged ONCE per K-chunk in the canonical
//   no-swizzle K-major layout [k-group][row][8 x fp16] and each tap is the same tile with the matrix
//   descriptor's start address advanced by one 16-byte row.
//
// Precision: fp32 parity with eps amplified by up to 4602x rules out a single 16-bit operand, so operands are split
// x = hi + lo * 2^-11 (two fp16 planes, the lo plane stored scaled by 2^11 to stay in fp16's normal range: 22 significant
// bits, tc_common.cuh split_hl) and each K step computes hi*hi, hi*lo and lo*hi with fp32 TMEM accumulation into separate
// columns, combined as hi*hi + (hi*lo + lo*hi) * 2^-11 in the epilogue (~2^-22 relative per product: the fp32 FMA path's
// own accuracy). The weight tile stores [hi rows | lo rows] as one N=64 operand, so A_hi x [W_hi;W_lo] yields hi*hi and
// hi*lo with a single shared-memory read of A_hi; A_lo x W_hi (N=32) is issued by the second issuer warp.
//
// Global "TC layout" of an activation [B, C, L] (one tensor per plane):
//   plane[tile][C/8][132][8]  fp16,  tile = b / SPT, row = (b % SPT) * (L + 4) + l + 2 ; halo / spare rows are
//   zero forever. A K-chunk of a CTA's tile is ONE contiguous block -> one cp.async.bulk per plane.
#include "tc_common.cuh"

namespace mpdb {


// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
// MODE: TCM_CONV5 = Conv1d k5 (+GN+Mish+cond+residual);  TCM_DOWN = Conv1d k3 stride 2 evaluated at every input position
// (3 taps) with only the even rows written (MMA work is negligible next to the fixed costs);  TCM_UP = ConvTranspose1d
// k4 stride 2 as two 2-tap accumulators (even / odd outputs).
// Warp roles (tools/probes/mma_probe.cu: one thread issues a tcgen05.mma every ~80-100 cycles whatever its shape, two
// threads in different warps reach the shared-memory operand floor; issued from a single-lane branch every MMA / bulk copy
// is wrapped in an ELECT + R2UR loop, so the issuing warps run their loops warp-convergently and elect one lane):
//   warps 0-15  epilogue (warp w reads TMEM lane quarter w & 3, column group w >> 2)
//   warp  16    producer (cp.async.bulk ring)
//   warp  17    MMA issuer 0: A_hi x [W_hi | W_lo] (N = 64) -> columns [0,64)   (second accumulator: +128)
//   warp  18    MMA issuer 1: A_lo x W_hi        (N = 32) -> columns [64,96)  (second accumulator: +128)
constexpr int TCL_THREADS = TC_THREADS + 96;
constexpr int TCL_ACC_COLS = 256;   // one accumulator stage: main [0,96) + second accumulator [128,224)
constexpr int TCL_TMEM_COLS = 512;  // two stages: the MMAs of work item k+1 run under the epilogue of item k

// PERSISTENT: a CTA walks the work items (row tile, 32-channel chunk) it = blockIdx.x, blockIdx.x + gridDim.x, ... of the layer
// (grid = min(#items, #SMs)). The operand ring runs on across items (the producer prefetches the next item's chunks while
// the current one is in its epilogue), the accumulators are double-buffered in TMEM (acc_full / acc_empty barriers per
// stage; the epilogue releases a stage as soon as its values are in registers), so per item the CTA pays
// max(MMA + operand streaming, epilogue) instead of setup + first-copy latency + MMA + epilogue + teardown: at 512
// trajectories per GPU a layer is 500-1000 items on 148 SMs and one-CTA-per-item launches spent ~10 us per item on ~2 us of work.
// Operand ring: 2-8 stages of [activations | weights], the geometry chosen per layer by launch_conv5_tc (TC_PS_RING_BYTES cut into
// stages of the layer's largest activation box + weight group). A stage holds a GROUP of K-chunks of one source
// tensor — 2 with the 22-bit split (both planes), 4 in precision 1 (hi plane only; fewer when the tensor is narrower) —
// brought by TWO asynchronous copies: one cp.async.bulk.tensor (TMA tensor map, 5-D box {8, 132 rows, 4 * chunks k-groups, 1
// tile, planes}: lands as [plane][k-group][row][8], the tcgen05 operand layout) and one cp.async.bulk for the group's weights
// (contiguous in the packed layout). A bulk copy costs ~800 cycles of the copy engine whatever its size
// (tools/probes/bulk_probe.cu), so the wide layers were bound by the NUMBER of copies: 3 per K-chunk before, 2 per group now.
// The epilogue is software-pipelined over the items (see the epilogue branch), and consecutive layer launches are chained by
// programmatic dependent launch (TcConvArgs::pdl): the next layer's CTA starts on an SM when this layer's CTA there has exited,
// prefetches weights and blocks in griddepcontrol.wait before it touches activations or writes anything.
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n\t}"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
        : "memory");
}

template <int MODE, int GS>
__global__ void __launch_bounds__(TCL_THREADS, 1) conv5_tc_kernel(const __grid_constant__ TcConvArgs a) {
    constexpr int NTAPS = MODE == TCM_CONV5 ? 5 : MODE == TCM_DOWN ? 3 : 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // carve-up: stages | barriers | tmem slot | epilogue scratch
    // ring geometry chosen by the launcher for this layer: NS stages of a.ps_stage_bytes = [activation box | weight group]
    // (narrow layers: many small stages = deep prefetch over the ~2 us L2 -> shared latency; wide layers: few large ones)
    const int NS = a.ps_stages;
    const uint32_t STAGE_BYTES = (uint32_t)a.ps_stage_bytes, ACT_BYTES = (uint32_t)a.ps_act_bytes;
    unsigned char* stages = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + TC_PS_RING_BYTES);  // full[8], empty[8], acc_full[2], acc_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_PS_MAX_STAGES + 4);
    float* part = reinterpret_cast<float*>(tmem_slot + 4);  // GroupNorm scratch, one copy per item parity (see the epilogue loop)
    // per-item epilogue parameters of the item's 32 output channels {bias, gamma, beta, residual bias, time-conditioning row at
    // uniform t}, filled by issuer warp 1 one item ahead: in the item loop the epilogue warps would otherwise pay a
    // global-memory round trip per item with nothing to hide it behind. Four slots: the software-pipelined epilogue reads the
    // rows of item k (gamma, beta, ...) after it has released the accumulator stage of item k + 1; slot k & 3 is rewritten
    // for item k + 4, after every epilogue warp has released the stage of item k + 2, which it does after finishing item k
    float* ptab = part + 2 * (TC_GN_SCRATCH_BYTES / (int)sizeof(float));  // [4][5][32]: slot k & 3 for the CTA's k-th item

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool dbg = a.dbg != nullptr && blockIdx.x == 0;
    if (dbg && tid == 64) a.dbg[0] = clock64();
    const int Lp = a.L + 4;
    const int SPT = TC_RT / Lp;
    const int NC = a.CO / TC_NT;
    const int n_items = ((a.B + SPT - 1) / SPT) * NC;  // item = tile * NC + ntile
    // precision 1 (a.prec == 1, engine.cu step_prec): one fp16 product per MMA step; the two issuers split it by K-group
    // into separate accumulators ([0,32) and [64,96)), the lo planes are neither loaded nor written
    const bool p1 = a.prec == 1;
    // K-chunk groups of an item, in order: main conv over source 0, source 1 (concatenated input), then the fused 1x1 residual
    // conv over its sources 2, 3. Source i: Cs[i] channels in groups of gsz[i] chunks (the box of its tensor map).
    const int Cs[4] = {a.c0, a.c1, a.res_w ? a.rc0 : 0, a.res_w ? a.rc1 : 0};
    int ng[4], n_groups = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ch = Cs[i] / TC_KCH;
        ng[i] = ch > 0 ? (ch + a.tm_nch[i] - 1) / a.tm_nch[i] : 0;
        n_groups += ng[i];
    }
    const int n_main_ch = (a.c0 + a.c1) / TC_KCH, n_res_ch = a.res_w ? (a.rc0 + a.rc1) / TC_KCH : 0;

    const uint32_t stages_u32 = smem_u32(stages);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TC_PS_MAX_STAGES);
    const uint32_t acc_full0 = smem_u32(bars + 2 * TC_PS_MAX_STAGES), acc_empty0 = acc_full0 + 16;

    // group gi of an item -> (source, first chunk within the source, chunks in the group, first chunk within main / residual conv)
    auto group_geom = [&](int gi, int& src, int& c_src, int& nch, int& c_conv) {
        src = 0;
        int g = gi;
        while (g >= ng[src]) { g -= ng[src]; ++src; }
        c_src = g * a.tm_nch[src];
        const int ch = Cs[src] / TC_KCH;
        nch = ch - c_src < a.tm_nch[src] ? ch - c_src : a.tm_nch[src];
        c_conv = c_src + ((src == 1) ? a.c0 / TC_KCH : (src == 3) ? a.rc0 / TC_KCH : 0);
    };

    // one K-chunk group of an item -> ring stage i % NS (producer warp only: the whole warp calls it, one elected lane issues)
    auto produce = [&](int i, int item, int gi, bool weights, bool acts) {
        const int tile = item / NC, ntile = item - tile * NC;
        const int s = i % NS;
        int src, c_src, nch, c_conv;
        group_geom(gi, src, c_src, nch, c_conv);
        const bool is_res = src >= 2;
        const int ntaps = is_res ? 1 : NTAPS;
        const uint32_t wmul = p1 ? 1u : 2u;  // precision 1 streams the hi halves only
        const uint32_t bbytes = (uint32_t)nch * wmul * ntaps * TC_B_TAP_BYTES;
        const uint32_t abytes = (uint32_t)a.tm_nch[src] * wmul * TC_A_PLANE_BYTES;  // the whole box (rows past the tensor are zero-filled)
        const uint32_t st = stages_u32 + (uint32_t)s * STAGE_BYTES;
        if (weights) {
            const size_t welems = is_res ? ((size_t)ntile * n_res_ch + c_conv) * (2 * 1 * TC_B_TAP_BYTES / 2)
                                         : ((size_t)ntile * n_main_ch + c_conv) * (2 * NTAPS * TC_B_TAP_BYTES / 2);
            const unsigned short* wsrc = p1 ? (is_res ? a.res_w_hi : a.w_hi) + welems / 2 : (is_res ? a.res_w : a.w) + welems;
            mbar_expect_tx_elect(full0 + 8 * s, abytes + bbytes);  // covers the activation copy too
            bulk_g2s_elect(st + ACT_BYTES, wsrc, bbytes, full0 + 8 * s);
        }
        if (acts) tma_load_5d(st, &a.tm[src], 0, 0, c_src * (TC_KCH / 8), tile, 0, full0 + 8 * s);
    };
    // The producer warp owns the operand ring: it initialises the ring's barriers itself and, outside programmatic dependent
    // launch, issues the FIRST item's first groups before the block-wide setup barrier, so they travel while TMEM is being
    // allocated (one produce() costs the warp ~0.3 us: two of them fit under the ~0.9 us of setup; issuing the whole ring
    // here held the barrier for 2.4 us and was slower).
    int n_early = 0;
    if (warp == TC_THREADS / 32) {
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) {
                mbar_init(full0 + 8 * s, 1);
                mbar_init(empty0 + 8 * s, 2);  // both issuers release a stage
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the copy engine completes transactions on these barriers
        }
        __syncwarp();
        if (!a.pdl) {
            const int e = n_groups < 2 ? n_groups : 2;
            for (; n_early < e && n_early < NS; ++n_early) produce(n_early, (int)blockIdx.x, n_early, true, true);
        }
    }
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(acc_full0 + 8 * s, 3);                  // both issuers' MMAs of the item have retired + issuer 1's parameter rows are in place
            mbar_init(acc_empty0 + 8 * s, TC_THREADS / 32);   // every epilogue warp has read its part of the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is a warp-wide operation; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TCL_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && tid == 64) a.dbg[1] = clock64();  // setup done (barriers, TMEM)

    if (warp == TC_THREADS / 32) {
        // ===== producer warp: group i of the CTA's item sequence -> stage i % STAGES. Weights are constants: the first groups'
        // weight copies are issued before the dependency wait, so under programmatic dependent launch they overlap the
        // previous kernel's epilogue. Activations follow it. =====
        int i = 0, pre = 0;
        if (a.pdl) {
            for (int item = blockIdx.x; item < n_items && pre < NS; item += gridDim.x)
                for (int gi = 0; gi < n_groups && pre < NS; ++gi, ++pre) produce(pre, item, gi, true, false);
            pdl_wait();
            pre = 0;
            for (int item = blockIdx.x; item < n_items && pre < NS; item += gridDim.x)
                for (int gi = 0; gi < n_groups && pre < NS; ++gi, ++pre) produce(pre, item, gi, false, true);
        } else {
            // plain stream / graph order (the default): nothing to overlap with, so every stage gets its weights AND its
            // activations at once, in ring order. One produce() costs the warp ~0.3 us; filling 8 stages weights-first meant the
            // activations of the FIRST group left ~2.5 us after the CTA started and every launch saw its first accumulator at ~4 us
            for (int item = blockIdx.x; item < n_items && pre < NS; item += gridDim.x)
                for (int gi = 0; gi < n_groups && pre < NS; ++gi, ++pre)
                    if (pre >= n_early) produce(pre, item, gi, true, true);  // the first one or two left before the setup barrier
        }
        // the ring is bounded by its own empty barriers only: operands run as far ahead of the MMAs as the stages allow
        for (int item = blockIdx.x; item < n_items; item += gridDim.x)
            for (int gi = 0; gi < n_groups; ++gi, ++i) {
                if (i < NS) continue;  // issued above
                mbar_wait(empty0 + 8 * (i % NS), ((uint32_t)(i / NS) & 1u) ^ 1u);
                __syncwarp();
                produce(i, item, gi, true, true);
            }
    } else if (warp > TC_THREADS / 32) {
        // ===== two MMA-issue warps =====
        const int which = __shfl_sync(0xffffffffu, warp, 0) - (TC_THREADS / 32 + 1);
        const uint32_t idesc = (which == 0 && !p1) ? tc_idesc(128, 2 * TC_NT) : tc_idesc(128, TC_NT);
        const uint32_t colw = __shfl_sync(0xffffffffu, tmem_base, 0) + (which == 0 ? 0u : 2u * TC_NT);
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);                      // SBO = 128 B, descriptor version 1
        constexpr uint32_t a_lo_fixed = ((uint32_t)(TC_RT * 16) >> 4) << 16;        // activation tile: LBO = 132 rows x 16 B
        // weight tile: rows [0,32) = W_hi, [32,64) = W_lo per k-group (LBO = 64 rows x 16 B); precision 1: W_hi only (LBO = 32 rows)
        const uint32_t b_lo_fixed = (((p1 ? 1u : 2u) * TC_NT * 16u) >> 4) << 16;
        constexpr uint32_t kstep_a = (2 * TC_RT * 16) >> 4;
        const uint32_t kstep_b = ((p1 ? 1u : 2u) * (2 * TC_NT * 16)) >> 4, tap_b = ((p1 ? 1u : 2u) * TC_B_TAP_BYTES) >> 4;
        // issuer 1 also keeps the epilogue's parameter table filled: the rows of the NEXT item are fetched into registers right
        // after this item's MMAs were issued (the loads fly while the warp waits on barriers) and written to the table slot
        // once the slot's previous reader has released it (the same acc_empty wait the MMAs need anyway)
        float prm[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        auto fetch_params = [&](int item) {
            const int c = (item % NC) * TC_NT + lane;
            prm[0] = a.bias ? a.bias[c] : 0.f;
            prm[1] = (MODE == TCM_CONV5 && a.gamma) ? a.gamma[c] : 0.f;
            prm[2] = (MODE == TCM_CONV5 && a.beta) ? a.beta[c] : 0.f;
            prm[3] = a.res_w ? a.res_bias[c] : 0.f;
            prm[4] = (a.cond && !a.t_dev) ? a.cond[(size_t)a.t_uniform * a.CO + c] : 0.f;
        };
        if (which == 1 && (int)blockIdx.x < n_items) fetch_params(blockIdx.x);
        int i = 0, k = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
            const int stage = k & 1;
            const uint32_t col0 = colw + (uint32_t)stage * TCL_ACC_COLS;
            // the epilogue of the item that used this stage two items ago has emptied it (the first use of a stage passes)
            mbar_wait(acc_empty0 + 8 * stage, (((uint32_t)k >> 1) & 1u) ^ 1u);
            tc_fence_after();
            if (which == 1) {
                float* pt = ptab + (k & 3) * 160;
#pragma unroll
                for (int j = 0; j < 5; ++j) pt[32 * j + lane] = prm[j];
                __syncwarp();
                if (lane == 0) mbar_arrive_local(acc_full0 + 8 * stage);  // release: the rows are visible to whoever sees the phase complete
            }
            uint32_t acc0 = 0u, acc1 = 0u;
            for (int gi = 0; gi < n_groups; ++gi, ++i) {
                const int s = i % NS;
                const uint32_t st = stages_u32 + (uint32_t)s * STAGE_BYTES;
                int src, c_src, nch, c_conv;
                group_geom(gi, src, c_src, nch, c_conv);
                const bool is_res = src >= 2;
                // the box lands as [plane][k-group][row][8]: the lo plane starts after the box's chunks of the hi plane
                const uint32_t plane_off = (which == 1 && !p1) ? (uint32_t)a.tm_nch[src] * TC_A_PLANE_BYTES : 0u;
                mbar_wait(full0 + 8 * s, (uint32_t)(i / NS) & 1u);
                tc_fence_after();
                for (int u = 0; u < nch; ++u) {
                    const uint32_t a_lo = (((st + plane_off + (uint32_t)u * TC_A_PLANE_BYTES) >> 4) & 0x3FFFu) | a_lo_fixed;
                    const uint32_t wst = st + ACT_BYTES + (uint32_t)u * (uint32_t)(is_res ? 1 : NTAPS) * (p1 ? 1u : 2u) * TC_B_TAP_BYTES;
                    const uint32_t b_lo = ((wst >> 4) & 0x3FFFu) | b_lo_fixed;
                    if (is_res) {  // fused 1x1 residual conv: centre row (+2), second accumulator
#pragma unroll
                        for (int kk = 0; kk < TC_KCH / 16; ++kk) {
                            if (p1 && kk != which) continue;
                            tc_mma_bf16_elect32(col0 + 128, a_lo + kk * kstep_a + 2, desc_hi, b_lo + kk * kstep_b, desc_hi, idesc, acc1);
                            acc1 = 1u;
                        }
                    } else {
#pragma unroll
                        for (int tap = 0; tap < NTAPS; ++tap) {
                            // row shift of the tap (in 16-byte rows; +2 is the centre) and the accumulator it feeds
                            //   CONV5: taps -2..2 -> shifts 0..4           DOWN (k3, pad 1): taps -1..1 -> shifts 1..3
                            //   UP  : packed taps [W1, W3 | W0, W2]: even = W1 x[m] + W3 x[m-1], odd = W0 x[m+1] + W2 x[m]
                            const int shift = MODE == TCM_CONV5 ? tap : MODE == TCM_DOWN ? tap + 1 : (tap == 0 ? 2 : tap == 1 ? 1 : tap == 2 ? 3 : 2);
                            const bool second_acc = MODE == TCM_UP && tap >= 2;
#pragma unroll
                            for (int kk = 0; kk < TC_KCH / 16; ++kk) {
                                if (p1 && kk != which) continue;
                                if (second_acc) { tc_mma_bf16_elect32(col0 + 128, a_lo + kk * kstep_a + shift, desc_hi, b_lo + tap * tap_b + kk * kstep_b, desc_hi, idesc, acc1); acc1 = 1u; }
                                else { tc_mma_bf16_elect32(col0, a_lo + kk * kstep_a + shift, desc_hi, b_lo + tap * tap_b + kk * kstep_b, desc_hi, idesc, acc0); acc0 = 1u; }
                            }
                        }
                    }
                }
                tc_commit_elect(empty0 + 8 * s);  // frees the stage when both issuers' MMAs that read it have retired
            }
            tc_commit_elect(acc_full0 + 8 * stage);  // accumulators of this item complete
            if (which == 1 && item + (int)gridDim.x < n_items) fetch_params(item + (int)gridDim.x);
        }
        // main loop over: let the next kernel start its prologue (barriers, TMEM, weight prefetch) under our last epilogues
        if (which == 0 && lane == 0) pdl_launch_dependents();
        if (dbg && which == 0 && lane == 0) a.dbg[3] = clock64();  // all MMAs issued
    } else {
        // ===== epilogue: 16 warps. TMEM lane quarter q = warp & 3 (hardware rule: a warp reads lanes 32*(warp%4)..+31),
        // column group cg = warp >> 2 -> each thread owns 8 consecutive channels of one row. Per-item geometry and
        // every parameter / residual value this thread will need are fetched BEFORE waiting for the accumulators, so
        // their global-memory latency overlaps the MMA main loop =====
        pdl_wait();  // nothing produced by the previous kernel is read, and nothing is written, before this
        const int q = warp & 3, cg = warp >> 2;
        const int r = q * 32 + lane;  // padded row of the tile = TMEM lane
        const int s = r / Lp, l = r - s * Lp;
        const bool full = a.raw_out == nullptr;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // (row tile, channel chunk) of the CTA's items without a division per item: item += gridDim.x moves them by (gq, gr)
        const int gq = (int)gridDim.x / NC, gr = (int)gridDim.x - gq * NC;
        int tile = (int)blockIdx.x / NC, ntile = (int)blockIdx.x - tile * NC;
        const bool row_ok = (s < SPT) && (l < a.L);
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + cg * 8;
        if (MODE == TCM_CONV5 && full) {
            // ---- Conv1d k5 + GroupNorm + Mish [+ cond] [+ residual], SOFTWARE-PIPELINED over the CTA's items ----
            // GroupNorm needs two block-wide hand-offs per item (row-group partial sums -> statistics -> everyone). Run item by
            // item that is three barriers with little work between them and 16 warps that mostly wait (ncu: 30 % of the
            // stall samples on the barriers, issue slots 35 % busy). Here iteration k does
            //     X:  statistics of item k (from the partial sums written one iteration earlier)
            //         + accumulators of item k+1 out of TMEM, bias, partial sums of item k+1          -> ONE barrier
            //     Y:  normalise + Mish + cond + residual + stores of item k
            // so a thread always has two independent instruction streams between barriers and an item costs one barrier
            // (two when a tile holds more than 8 samples). Scratch (partial sums, statistics) alternates with the item parity:
            // item k+2 writes partial sums after the barrier of iteration k+1, i.e. after the statistics of item k were read
            // (iteration k, X) — and statistics after the barrier of iteration k+1, i.e. after every thread finished Y of item k.
            constexpr int SCR = TC_GN_SCRATCH_BYTES / (int)sizeof(float);
            const bool has_resw = a.res_w != nullptr, has_rid = a.res_w == nullptr && a.res_cm != nullptr;
            auto load_item = [&](int kk, int tile_, int ntile_, float (&v)[8], float (&z)[8]) {
                const int st = kk & 1;
                const int b_ = tile_ * SPT + s;
                const bool valid_ = row_ok && (b_ < a.B);
#pragma unroll
                for (int j = 0; j < 8; ++j) z[j] = 0.f;
                if (has_rid && valid_) {  // identity residual: requested before the accumulator wait, consumed one barrier later
                    const float* rp = a.res_cm + ((size_t)b_ * a.CO + ntile_ * TC_NT + cg * 8) * Lp + 2 + l;
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] = rp[(size_t)j * Lp];
                }
                mbar_wait(acc_full0 + 8 * st, ((uint32_t)kk >> 1) & 1u);
                __syncwarp();
                tc_fence_after();
                if (dbg && tid == 64 && kk == 0) a.dbg[4] = clock64();  // accumulators ready
                const uint32_t taddr = taddr0 + (uint32_t)st * TCL_ACC_COLS;
                tc_load_acc(taddr, p1, v);
                if (has_resw) tc_load_acc(taddr + 128, p1, z);  // the block's 1x1 residual conv
                const float* pt = ptab + (kk & 3) * 160 + cg * 8;
                const float4 pb0 = *reinterpret_cast<const float4*>(pt), pb1 = *reinterpret_cast<const float4*>(pt + 4);
                // the accumulators are in registers: hand the stage back (the MMAs of the item after next may start)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_local(acc_empty0 + 8 * st);
                if (dbg && tid == 64 && kk == 0) a.dbg[5] = clock64();  // TMEM read
                v[0] += pb0.x; v[1] += pb0.y; v[2] += pb0.z; v[3] += pb0.w;
                v[4] += pb1.x; v[5] += pb1.y; v[6] += pb1.z; v[7] += pb1.w;
                if (has_resw) {
                    const float4 pr0 = *reinterpret_cast<const float4*>(pt + 96), pr1 = *reinterpret_cast<const float4*>(pt + 100);
                    z[0] += pr0.x; z[1] += pr0.y; z[2] += pr0.z; z[3] += pr0.w;
                    z[4] += pr1.x; z[5] += pr1.y; z[6] += pr1.z; z[7] += pr1.w;
                }
                gn_level0(v, valid_, r, cg, part + st * SCR);
            };
            float v0[8], z0[8];  // item k: conv accumulator + bias; residual values (1x1 conv + its bias | identity | zero)
            if ((int)blockIdx.x < n_items) load_item(0, tile, ntile, v0, z0);
            epi_sync();
            int k = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
                float* scr = part + (k & 1) * SCR;
                const int b = tile * SPT + s;
                const bool valid = row_ok && (b < a.B);
                const int c8 = ntile * TC_NT + cg * 8;  // first of this thread's 8 output channels
                int tile_n = tile + gq, ntile_n = ntile + gr;
                if (ntile_n >= NC) { ntile_n -= NC; ++tile_n; }
                const bool has_next = item + (int)gridDim.x < n_items;
                float v1[8], z1[8];
                // ---- X ----
                if (SPT <= 8) {
                    gn_stats_small<GS>(tid, SPT, Lp, a.L, scr);
                    if (has_next) load_item(k + 1, tile_n, ntile_n, v1, z1);
                    epi_sync();
                } else {
                    gn_stats_big1(tid, SPT, Lp, a.L, scr);
                    if (has_next) load_item(k + 1, tile_n, ntile_n, v1, z1);
                    epi_sync();
                    gn_stats_big2<GS>(tid, SPT, a.L, scr);
                    epi_sync();
                }
                if (dbg && tid == 64 && k == 0) a.dbg[9] = clock64();  // statistics of item 0 published
                // ---- Y ----
                const float* pt = ptab + (k & 3) * 160 + cg * 8;
                const float4 pg0 = *reinterpret_cast<const float4*>(pt + 32), pg1 = *reinterpret_cast<const float4*>(pt + 36);
                const float4 pe0 = *reinterpret_cast<const float4*>(pt + 64), pe1 = *reinterpret_cast<const float4*>(pt + 68);
                float4 pc0 = z4, pc1 = z4;
                if (a.t_dev == nullptr) { pc0 = *reinterpret_cast<const float4*>(pt + 128); pc1 = *reinterpret_cast<const float4*>(pt + 132); }
                else if (a.cond != nullptr && valid) {  // per-sample t (per-call entry points): the row depends on the sample
                    const float* cp = a.cond + (size_t)a.t_dev[b] * a.CO + c8;
                    pc0 = *reinterpret_cast<const float4*>(cp); pc1 = *reinterpret_cast<const float4*>(cp + 4);
                }
                gn_apply<GS>(v0, valid, s, cg, SPT, scr, pg0, pg1, pe0, pe1);
                if (dbg && tid == 64 && k == 0) a.dbg[6] = clock64();  // GroupNorm + Mish done
                // time conditioning (zero when absent), then the residual: fused 1x1 conv accumulator + bias or identity values
                v0[0] += pc0.x; v0[1] += pc0.y; v0[2] += pc0.z; v0[3] += pc0.w;
                v0[4] += pc1.x; v0[5] += pc1.y; v0[6] += pc1.z; v0[7] += pc1.w;
#pragma unroll
                for (int j = 0; j < 8; ++j) v0[j] += z0[j];
                if (valid) tc_store_row_at(a, v0, b, tile, s, l, c8, a.L);  // same length: same (row tile, slot) as the input
                if (has_next) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v0[j] = v1[j]; z0[j] = z1[j]; }
                }
                tile = tile_n; ntile = ntile_n;
            }
        } else {
            int k = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
                const int stage = k & 1;
                const int b = tile * SPT + s;
                const bool valid = row_ok && (b < a.B);
                const int c8 = ntile * TC_NT + cg * 8;  // first of this thread's 8 output channels
                mbar_wait(acc_full0 + 8 * stage, ((uint32_t)k >> 1) & 1u);
                __syncwarp();
                tc_fence_after();
                const uint32_t taddr = taddr0 + (uint32_t)stage * TCL_ACC_COLS;
                if (dbg && tid == 64 && k == 0) a.dbg[4] = clock64();  // accumulators ready
                float v[8], w[8];  // main accumulator; second accumulator (odd outputs of UP)
                tc_load_acc(taddr, p1, v);
                if (MODE == TCM_UP) tc_load_acc(taddr + 128, p1, w);
                const float* pt = ptab + (k & 3) * 160 + cg * 8;  // parameter rows of the item's channels (issuer-filled table)
                const float4 pb0 = *reinterpret_cast<const float4*>(pt), pb1 = *reinterpret_cast<const float4*>(pt + 4);
                // the stage (accumulators and parameter rows) is in registers: hand it back (the MMAs of the item after next may start)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_local(acc_empty0 + 8 * stage);
                if (dbg && tid == 64 && k == 0) a.dbg[5] = clock64();  // TMEM read
                if (!full) {
                    float* dst = a.raw_out + (((size_t)tile * NC + ntile) * 128 + r) * 32 + cg * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = v[j];
                } else if (MODE == TCM_DOWN) {
                    // stride-2 conv: the MMA evaluated every input position; keep the even ones (out[m] = conv at l = 2m)
                    v[0] += pb0.x; v[1] += pb0.y; v[2] += pb0.z; v[3] += pb0.w;
                    v[4] += pb1.x; v[5] += pb1.y; v[6] += pb1.z; v[7] += pb1.w;
                    if (valid && (l & 1) == 0) tc_store_row(a, v, b, l >> 1, c8, a.L >> 1);
                } else if (MODE == TCM_UP) {
                    // transposed conv: accumulator 0 = even outputs (2l), accumulator 1 = odd outputs (2l + 1)
                    v[0] += pb0.x; v[1] += pb0.y; v[2] += pb0.z; v[3] += pb0.w;
                    v[4] += pb1.x; v[5] += pb1.y; v[6] += pb1.z; v[7] += pb1.w;
                    w[0] += pb0.x; w[1] += pb0.y; w[2] += pb0.z; w[3] += pb0.w;
                    w[4] += pb1.x; w[5] += pb1.y; w[6] += pb1.z; w[7] += pb1.w;
                    if (valid) {
                        tc_store_row(a, v, b, 2 * l, c8, 2 * a.L);
                        tc_store_row(a, w, b, 2 * l + 1, c8, 2 * a.L);
                    }
                }
                int tile_n = tile + gq, ntile_n = ntile + gr;
                if (ntile_n >= NC) { ntile_n -= NC; ++tile_n; }
                tile = tile_n; ntile = ntile_n;
            }
        }
        tc_fence_before();
    }

    // teardown: all TMEM reads done before the allocating warp frees the columns
    __syncthreads();
    if (dbg && tid == 64) a.dbg[7] = clock64();  // stores issued, CTA done
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCL_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Cluster-fused ResidualTemporalBlock (layers.py:323-355) for C_out <= 128: both k=5 convolutions in ONE launch.
// The CO/32 CTAs that share a row tile form a thread-block cluster. Each runs conv0 for its 32 channels as in
// conv5_tc_kernel, and its epilogue writes the fp16-split h1 values straight into the conv1 A-operand buffer ("A2",
// the full C_out x 132-row tile, both planes) of EVERY CTA of the cluster through distributed shared memory; a
// cluster-scope mbarrier per CTA counts the writers. conv1 then reads its activations from local shared memory (no
// global round trip, no second launch) while its weights stream through the same ring. Saves one kernel boundary
// (launch edge + setup + first-copy wait) per block.
// Threads: 16 epilogue warps (one of them hosts the MMA issuer lane) + 1 producer warp.
// ---------------------------------------------------------------------------------------------------
constexpr int RTB_THREADS = TC_THREADS + 32;
constexpr int RTB_TMEM_COLS = 512;  // hi*hi | hi*lo: conv0 [0,64), conv1 [64,128), residual conv [128,192); lo*hi: the same + 256 ([.., +32))


// accumulator of one tile of the block kernel: [t, t+32) hi*hi, [t+32, t+64) hi*lo, [t+256, t+288) lo*hi; precision 1: the
// single product sits in [t, t+32)
__device__ __forceinline__ void rtb_load_acc(uint32_t t, bool p1, float (&v)[8]) {
    if (p1) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
    } else {
        float v2[8], v3[8];
        tc_ld8x3(t, t + 256, t + TC_NT, v, v3, v2);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v3[j] + v2[j], TC_LO_UNSCALE, v[j]);
    }
}

template <int GS, int NSTAGE>
__global__ void __launch_bounds__(RTB_THREADS, 1) rtb_tc_kernel(TcRtbArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TcConvArgs& a0 = args.c0;
    const TcConvArgs& a1 = args.c1;
    const int CO = a0.CO;
    const int a2_plane = (CO / 8) * TC_RT * 16;  // bytes per plane of the conv1 operand
    unsigned char* a2 = smem_raw;                // [2 planes][CO/8][132][16 B]
    unsigned char* stages = a2 + 2 * a2_plane;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + NSTAGE * TC_STAGE_BYTES);  // full[S], empty[S], done1, done2, a2_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 3);
    float* part = reinterpret_cast<float*>(tmem_slot + 4);  // GroupNorm scratch: [2][128][8] floats + [2][12][8] doubles

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, ntile = blockIdx.y;
    const int CS = CO / TC_NT;  // cluster size = CTAs per row tile
    const int n0 = ntile * TC_NT;
    const int Lp = a0.L + 4;
    const int SPT = TC_RT / Lp;
    const int n1 = (a0.c0 + a0.c1) / TC_KCH;                 // conv0 chunks (activations + weights)
    const int n2 = CO / TC_KCH;                              // conv1 chunks (weights only)
    const int n3 = a1.res_w ? (a1.rc0 + a1.rc1) / TC_KCH : 0;  // residual 1x1 conv chunks (activations + weights)
    const int n_steps = n1 + n2 + n3;
    const bool p1 = a0.prec == 1;  // one fp16 product per MMA step (see conv5_tc_kernel); the lo planes are not touched
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NSTAGE);
    const uint32_t done1 = smem_u32(bars + 2 * NSTAGE), done2 = done1 + 8, a2_full = done1 + 16;

    // ---- setup ----
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(done1, 1);
        mbar_init(done2, 1);
        mbar_init(a2_full, (uint32_t)(CS * TC_THREADS));  // every epilogue thread of every CTA of the cluster arrives once
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(RTB_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 2 * a2_plane / 16; i += RTB_THREADS) reinterpret_cast<uint4*>(a2)[i] = make_uint4(0u, 0u, 0u, 0u);  // halos
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // peers must have zeroed their A2 and initialised their barriers before anyone writes / arrives remotely
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == TC_THREADS / 32) {
        // ===== producer warp (one lane): conv0 chunks, conv1 weight chunks, residual-conv chunks through one ring =====
        if (lane == 0) {
            for (int i = 0; i < n_steps; ++i) {
                const int s = i % NSTAGE;
                if (i >= NSTAGE) mbar_wait(empty0 + 8 * s, ((uint32_t)(i / NSTAGE) & 1u) ^ 1u);
                const uint32_t st = smem_u32(stages + (size_t)s * TC_STAGE_BYTES);
                const int phase = i < n1 ? 0 : i < n1 + n2 ? 1 : 2;
                const int c = phase == 0 ? i : phase == 1 ? i - n1 : i - n1 - n2;
                const int ntaps = phase == 2 ? 1 : 5;
                const uint32_t bbytes = 2u * ntaps * TC_B_TAP_BYTES;
                const TcConvArgs& A = phase == 0 ? a0 : a1;
                const unsigned short* wsrc = phase == 0   ? a0.w + ((size_t)ntile * n1 + c) * (2 * 5 * TC_B_TAP_BYTES / 2)
                                             : phase == 1 ? a1.w + ((size_t)ntile * n2 + c) * (2 * 5 * TC_B_TAP_BYTES / 2)
                                                          : a1.res_w + ((size_t)ntile * n3 + c) * (2 * 1 * TC_B_TAP_BYTES / 2);
                mbar_expect_tx(full0 + 8 * s, bbytes + (phase == 1 ? 0u : (p1 ? 1u : 2u) * TC_A_PLANE_BYTES));
                bulk_g2s(st + 2 * TC_A_PLANE_BYTES, wsrc, bbytes, full0 + 8 * s);
                if (phase != 1) {
                    const int C0 = phase == 0 ? A.c0 : A.rc0, C1 = phase == 0 ? A.c1 : A.rc1;
                    const bool second = c * TC_KCH >= C0;
                    const unsigned short* ahi = phase == 0 ? (second ? A.in1_hi : A.in0_hi) : (second ? A.r1_hi : A.r0_hi);
                    const unsigned short* alo = phase == 0 ? (second ? A.in1_lo : A.in0_lo) : (second ? A.r1_lo : A.r0_lo);
                    const int Csrc = second ? C1 : C0;
                    const int kg0 = (c * TC_KCH - (second ? C0 : 0)) / 8;
                    const size_t aoff = ((size_t)tile * (Csrc / 8) + kg0) * TC_RT * 8;
                    bulk_g2s(st, ahi + aoff, TC_A_PLANE_BYTES, full0 + 8 * s);
                    if (!p1) bulk_g2s(st + TC_A_PLANE_BYTES, alo + aoff, TC_A_PLANE_BYTES, full0 + 8 * s);
                }
            }
        }
        return;  // the producer warp takes no part in the epilogues (their barriers count TC_THREADS threads)
    }

    constexpr uint32_t idesc32 = tc_idesc(128, TC_NT), idesc64 = tc_idesc(128, 2 * TC_NT);
    // MMAs of one ring step. a_hi/a_lo: shared addresses of the activation planes for this K-chunk.
    auto issue_step = [&](int i, uint32_t a_hi, uint32_t a_lo, int ntaps, bool centre, uint32_t dcol, bool& first) {
        const int s = i % NSTAGE;
        mbar_wait(full0 + 8 * s, (uint32_t)(i / NSTAGE) & 1u);
        tc_fence_after();
        const uint32_t st = smem_u32(stages + (size_t)s * TC_STAGE_BYTES);
        const uint64_t dA_hi = tc_desc(a_hi ? a_hi : st, TC_RT * 16, 128);
        const uint64_t dA_lo = tc_desc(a_lo ? a_lo : st + TC_A_PLANE_BYTES, TC_RT * 16, 128);
        const uint64_t dB = tc_desc(st + 2 * TC_A_PLANE_BYTES, 2 * TC_NT * 16, 128);
        bool first_lo = first;  // the lo*hi accumulator starts with the same K-chunk as the main one
        for (int tap = 0; tap < ntaps; ++tap) {
            const int shift = centre ? 2 : tap;
#pragma unroll
            for (int kk = 0; kk < TC_KCH / 16; ++kk) {
                const uint64_t aofs = (uint64_t)((kk * 2 * (TC_RT * 16) + shift * 16) >> 4);
                const uint64_t bofs = (uint64_t)((tap * 2 * TC_B_TAP_BYTES + kk * 2 * (2 * TC_NT * 16)) >> 4);
                tc_mma_bf16(dcol, dA_hi + aofs, dB + bofs, p1 ? idesc32 : idesc64, first ? 0u : 1u);
                first = false;
                if (!p1) tc_mma_bf16(dcol + 256, dA_lo + aofs, dB + bofs, idesc32, first_lo ? 0u : 1u);  // lo*hi: own columns (scaled sum)
                first_lo = false;
            }
        }
        tc_commit(empty0 + 8 * s);
    };

    // ===== phase 1: conv0 =====
    if (tid == 32) {
        bool first = true;
        for (int i = 0; i < n1; ++i) issue_step(i, 0u, 0u, 5, false, tmem_base, first);
        tc_commit(done1);
    }
    const int q = warp & 3, cg = warp >> 2;
    const int r = q * 32 + lane;
    const int s = r / Lp, l = r - s * Lp;
    const int b = tile * SPT + s;
    const bool valid = (s < SPT) && (l < a0.L) && (b < a0.B);
    const int c8 = n0 + cg * 8;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pb0 = *reinterpret_cast<const float4*>(a0.bias + c8), pb1 = *reinterpret_cast<const float4*>(a0.bias + c8 + 4);
    float4 pg0 = *reinterpret_cast<const float4*>(a0.gamma + c8), pg1 = *reinterpret_cast<const float4*>(a0.gamma + c8 + 4);
    float4 pe0 = *reinterpret_cast<const float4*>(a0.beta + c8), pe1 = *reinterpret_cast<const float4*>(a0.beta + c8 + 4);
    float4 pc0 = z4, pc1 = z4;
    if (a0.cond != nullptr && valid) {
        const int tt = a0.t_dev ? (int)a0.t_dev[b] : a0.t_uniform;
        const float* cp = a0.cond + (size_t)tt * CO + c8;
        pc0 = *reinterpret_cast<const float4*>(cp); pc1 = *reinterpret_cast<const float4*>(cp + 4);
    }
    mbar_wait(done1, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cg * 8;
    float v[8];
    rtb_load_acc(taddr, p1, v);  // hi*hi, lo*hi, hi*lo
    v[0] += pb0.x; v[1] += pb0.y; v[2] += pb0.z; v[3] += pb0.w;
    v[4] += pb1.x; v[5] += pb1.y; v[6] += pb1.z; v[7] += pb1.w;
    gn_mish8<GS, true>(v, valid, r, s, cg, tid, SPT, Lp, a0.L, part, pg0, pg1, pe0, pe1);
    v[0] += pc0.x; v[1] += pc0.y; v[2] += pc0.z; v[3] += pc0.w;
    v[4] += pc1.x; v[5] += pc1.y; v[6] += pc1.z; v[7] += pc1.w;
    {
        // h1 -> the conv1 operand buffer of every CTA of the cluster (k-group = c8 / 8, row = s*Lp + l + 2)
        uint4 ph = make_uint4(0u, 0u, 0u, 0u), pl = ph;
        if (valid) {
            unsigned short h[8], lo8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_hl(v[e], h[e], lo8[e]);
            ph.x = h[0] | ((uint32_t)h[1] << 16); ph.y = h[2] | ((uint32_t)h[3] << 16);
            ph.z = h[4] | ((uint32_t)h[5] << 16); ph.w = h[6] | ((uint32_t)h[7] << 16);
            pl.x = lo8[0] | ((uint32_t)lo8[1] << 16); pl.y = lo8[2] | ((uint32_t)lo8[3] << 16);
            pl.z = lo8[4] | ((uint32_t)lo8[5] << 16); pl.w = lo8[6] | ((uint32_t)lo8[7] << 16);
        }
        const uint32_t off = (uint32_t)(((c8 / 8) * TC_RT + (s * Lp + l + 2)) * 16);
        const uint32_t local_hi = smem_u32(a2) + off, local_lo = local_hi + a2_plane;
        for (int cta = 0; cta < CS; ++cta) {
            if (valid) {
                st_cluster_v4(map_to_cta(local_hi, cta), ph);
                if (!p1) st_cluster_v4(map_to_cta(local_lo, cta), pl);
            }
        }
        // generic-proxy stores must be visible to the tensor core (async proxy) of the consumer CTAs
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int cta = 0; cta < CS; ++cta) mbar_arrive_cluster(map_to_cta(a2_full, cta));
    }

    // ===== phase 2: conv1 (activations from A2) + residual 1x1 conv =====
    // parameters of the second epilogue are fetched while the MMAs run
    pb0 = *reinterpret_cast<const float4*>(a1.bias + c8); pb1 = *reinterpret_cast<const float4*>(a1.bias + c8 + 4);
    pg0 = *reinterpret_cast<const float4*>(a1.gamma + c8); pg1 = *reinterpret_cast<const float4*>(a1.gamma + c8 + 4);
    pe0 = *reinterpret_cast<const float4*>(a1.beta + c8); pe1 = *reinterpret_cast<const float4*>(a1.beta + c8 + 4);
    float4 pr0 = z4, pr1 = z4;
    float rid[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (a1.res_w != nullptr) {
        pr0 = *reinterpret_cast<const float4*>(a1.res_bias + c8); pr1 = *reinterpret_cast<const float4*>(a1.res_bias + c8 + 4);
    } else if (a1.res_cm != nullptr && valid) {
        const float* rp = a1.res_cm + ((size_t)b * CO + c8) * Lp + 2 + l;
#pragma unroll
        for (int j = 0; j < 8; ++j) rid[j] = rp[(size_t)j * Lp];
    }
    if (tid == 32) {
        mbar_wait_cluster(a2_full, 0);  // all h1 rows of all channel tiles have landed in our A2
        asm volatile("fence.proxy.async;" ::: "memory");
        tc_fence_after();
        bool first = true;
        const uint32_t a2_hi = smem_u32(a2), a2_lo = a2_hi + a2_plane;
        for (int c = 0; c < n2; ++c)
            issue_step(n1 + c, a2_hi + c * (TC_KCH / 8) * TC_RT * 16, a2_lo + c * (TC_KCH / 8) * TC_RT * 16, 5, false,
                       tmem_base + 2 * TC_NT, first);
        bool first_r = true;
        for (int c = 0; c < n3; ++c) issue_step(n1 + n2 + c, 0u, 0u, 1, true, tmem_base + 4 * TC_NT, first_r);
        tc_commit(done2);
    }
    mbar_wait(done2, 0);
    __syncwarp();
    tc_fence_after();
    rtb_load_acc(taddr + 2 * TC_NT, p1, v);
    v[0] += pb0.x; v[1] += pb0.y; v[2] += pb0.z; v[3] += pb0.w;
    v[4] += pb1.x; v[5] += pb1.y; v[6] += pb1.z; v[7] += pb1.w;
    gn_mish8<GS, true>(v, valid, r, s, cg, tid, SPT, Lp, a0.L, part, pg0, pg1, pe0, pe1);
    if (a1.res_w != nullptr) {
        float rv[8];
        rtb_load_acc(taddr + 4 * TC_NT, p1, rv);
        v[0] += rv[0] + pr0.x; v[1] += rv[1] + pr0.y; v[2] += rv[2] + pr0.z; v[3] += rv[3] + pr0.w;
        v[4] += rv[4] + pr1.x; v[5] += rv[5] + pr1.y; v[6] += rv[6] + pr1.z; v[7] += rv[7] + pr1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += rid[j];
    }
    if (valid) tc_store_row(a1, v, b, l, c8, a0.L);

    tc_fence_before();
    epi_sync();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(RTB_TMEM_COLS) : "memory");
    }
}

int launch_rtb_tc(const TcRtbArgs& a, cudaStream_t stream) {
    const TcConvArgs& a0 = a.c0;
    const TcConvArgs& a1 = a.c1;
    MPDB_REQUIRE(a0.CO == a1.CO && a0.L == a1.L && a0.B == a1.B && a0.gs == a1.gs, "rtb: the two convs disagree");
    MPDB_REQUIRE(a0.CO % TC_NT == 0 && a0.CO <= 128, "rtb: C_out must be 32, 64, 96 or 128");
    MPDB_REQUIRE(a0.c0 % TC_KCH == 0 && a0.c1 % TC_KCH == 0 && a0.c0 > 0, "rtb: input widths must be multiples of 32");
    MPDB_REQUIRE(!a1.res_w || (a1.rc0 % TC_KCH == 0 && a1.rc1 % TC_KCH == 0 && a1.rc0 > 0), "rtb: residual widths");
    MPDB_REQUIRE(a0.L + 4 <= TC_RT && a0.L % 4 == 0, "rtb: L too large for one 128-row tile");
    MPDB_REQUIRE(a0.gs == 4 || a0.gs == 8 || a0.gs == 16 || a0.gs == 32, "rtb: bad GroupNorm group size");
    const int SPT = TC_RT / (a0.L + 4);
    const int CS = a0.CO / TC_NT;
    const int nstage = a0.CO <= 64 ? 5 : 4;
    const size_t smem = (size_t)2 * (a0.CO / 8) * TC_RT * 16 + (size_t)nstage * TC_STAGE_BYTES + (2 * nstage + 3) * 8 + 16 +
                        (2 * 128 * 8 + 12 * 8 * 2) * sizeof(float) + 2 * 12 * 8 * sizeof(double);
    MPDB_REQUIRE(smem <= 227 * 1024, "rtb: shared memory budget exceeded");
    dim3 grid((a0.B + SPT - 1) / SPT, CS);
#define MPDB_RTB_LAUNCH(G, S)                                                                                          \
    {                                                                                                                  \
        static unsigned long long configured = 0ull;                                                                                \
        if (mpdb::first_use_on_device(configured)) {                                                                                             \
            MPDB_CHECK_CUDA(cudaFuncSetAttribute(rtb_tc_kernel<G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
        }                                                                                                              \
        MPDB_CHECK_CUDA(launch_kernel_cluster(rtb_tc_kernel<G, S>, grid, dim3(RTB_THREADS), smem, stream, (unsigned)CS, a)); \
    }
    if (nstage == 5) {
        if (a0.gs == 4) MPDB_RTB_LAUNCH(4, 5) else if (a0.gs == 8) MPDB_RTB_LAUNCH(8, 5) else if (a0.gs == 16) MPDB_RTB_LAUNCH(16, 5) else MPDB_RTB_LAUNCH(32, 5)
    } else {
        if (a0.gs == 4) MPDB_RTB_LAUNCH(4, 4) else if (a0.gs == 8) MPDB_RTB_LAUNCH(8, 4) else if (a0.gs == 16) MPDB_RTB_LAUNCH(16, 4) else MPDB_RTB_LAUNCH(32, 4)
    }
#undef MPDB_RTB_LAUNCH
    MPDB_LAUNCH_CHECK();
    return 0;
}

int launch_conv5_tc(const TcConvArgs& a_in, cudaStream_t stream) {
    const TcConvArgs& a0_ = a_in;
#define a a0_
    MPDB_REQUIRE(a.CO % TC_NT == 0, "tc conv: C_out must be a multiple of 32");
    MPDB_REQUIRE(a.c0 % TC_KCH == 0 && a.c1 % TC_KCH == 0 && a.c0 > 0, "tc conv: input widths must be multiples of 32");
    MPDB_REQUIRE(!a.res_w || (a.rc0 % TC_KCH == 0 && a.rc1 % TC_KCH == 0 && a.rc0 > 0), "tc conv: residual widths");
    MPDB_REQUIRE(a.L + 4 <= TC_RT && a.L % 4 == 0, "tc conv: L too large for one 128-row tile");
    MPDB_REQUIRE(a.mode == TCM_CONV5 || a.mode == TCM_DOWN || a.mode == TCM_UP, "tc conv: bad mode");
    if (a.mode == TCM_CONV5)
        MPDB_REQUIRE(a.gamma && a.beta && (a.gs == 4 || a.gs == 8 || a.gs == 16 || a.gs == 32),
                     "tc conv: GroupNorm group size must be 4, 8, 16 or 32");
    else
        MPDB_REQUIRE(!a.res_w && !a.res_cm && !a.cond && !a.raw_out, "tc down/up: no residual / conditioning");
    const int SPT = TC_RT / (a.L + 4);
    MPDB_REQUIRE(SPT <= 12, "tc conv: too many samples per tile");
#undef a
    const size_t smem = (size_t)TC_PS_RING_BYTES + (2 * TC_PS_MAX_STAGES + 4) * 8 + 16 + 2 * TC_GN_SCRATCH_BYTES + 4 * 160 * sizeof(float) + 64;
    static_assert((size_t)TC_PS_RING_BYTES + (2 * TC_PS_MAX_STAGES + 4) * 8 + 16 + 2 * TC_GN_SCRATCH_BYTES + 4 * 160 * sizeof(float) + 64 <= 227 * 1024, "conv5_tc_kernel: shared memory");
    TcConvArgs a = a_in;
    a.pdl = (g_use_pdl || g_pdl_layers) ? 1 : 0;
    {   // ring geometry: a stage = the largest activation box + the largest weight group of this layer
        const int wmul = a.prec == 1 ? 1 : 2, ntaps = a.mode == TCM_DOWN ? 3 : a.mode == TCM_UP ? 4 : 5;
        int max_nch = 1;
        for (int i = 0; i < 4; ++i) {
            const int Ci = i == 0 ? a.c0 : i == 1 ? a.c1 : i == 2 ? (a.res_w ? a.rc0 : 0) : (a.res_w ? a.rc1 : 0);
            if (Ci > 0 && a.tm_nch[i] > max_nch) max_nch = a.tm_nch[i];
        }
        a.ps_act_bytes = max_nch * wmul * TC_A_PLANE_BYTES;
        a.ps_stage_bytes = a.ps_act_bytes + max_nch * wmul * ntaps * TC_B_TAP_BYTES;
        a.ps_stage_bytes = (a.ps_stage_bytes + 127) / 128 * 128;
        a.ps_stages = TC_PS_RING_BYTES / a.ps_stage_bytes;
        if (a.ps_stages > TC_PS_MAX_STAGES) a.ps_stages = TC_PS_MAX_STAGES;
        MPDB_REQUIRE(a.ps_stages >= 2, "tc conv: operand ring too small for this layer");
    }
    for (int i = 0; i < 4; ++i) {
        const int Ci = i == 0 ? a.c0 : i == 1 ? a.c1 : i == 2 ? (a.res_w ? a.rc0 : 0) : (a.res_w ? a.rc1 : 0);
        MPDB_REQUIRE(Ci == 0 || (a.tm_nch[i] >= 1 && a.tm_nch[i] <= (a.prec == 1 ? 4 : 2)), "tc conv: missing / oversized activation tensor map");
    }
    // persistent: one CTA per SM (TMEM holds two accumulator stages) walking the (row tile, channel chunk) items
    static int sm_count[64] = {0};
    int dev = 0;
    MPDB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && sm_count[dev] == 0) MPDB_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    const int sms = (dev >= 0 && dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
    const long long n_items = (long long)((a.B + SPT - 1) / SPT) * (a.CO / TC_NT);
    dim3 grid((unsigned)(n_items < sms ? n_items : sms));
#define MPDB_TC_LAUNCH(M, G)                                                                                       \
    {                                                                                                              \
        static unsigned long long configured = 0ull;                                                                            \
        if (mpdb::first_use_on_device(configured)) {                                                                                         \
            MPDB_CHECK_CUDA(cudaFuncSetAttribute(conv5_tc_kernel<M, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 227 * 1024));                                                     \
        }                                                                                                          \
        MPDB_CHECK_CUDA(launch_kernel_pdl(conv5_tc_kernel<M, G>, grid, dim3(TCL_THREADS), smem, stream, a.pdl != 0, a)); \
    }
    if (a.mode == TCM_DOWN) MPDB_TC_LAUNCH(TCM_DOWN, 4)
    else if (a.mode == TCM_UP) MPDB_TC_LAUNCH(TCM_UP, 4)
    else if (a.gs == 4) MPDB_TC_LAUNCH(TCM_CONV5, 4)
    else if (a.gs == 8) MPDB_TC_LAUNCH(TCM_CONV5, 8)
    else if (a.gs == 16) MPDB_TC_LAUNCH(TCM_CONV5, 16)
    else MPDB_TC_LAUNCH(TCM_CONV5, 32)
#undef MPDB_TC_LAUNCH
    MPDB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// packing helpers
// ---------------------------------------------------------------------------------------------------
// weights: src fp32 [ci][ntaps][CO] (the SIMT path's packed layout) ->
//   dst fp16 [CO/32][CI/32][ntaps][kg 4][64 rows: 32 hi | 32 lo][8]
//   CI is the padded input width (multiple of 32); channels >= CI_src are zero. `perm` maps destination tap -> source tap
//   (ConvTranspose: [1, 3, 0, 2], see the kernel), packed 4 bits per tap.
// dst_hi (optional): the hi halves alone, [CO/32][CI/32][ntaps][kg 4][32 rows][8] — what a precision-1 step streams
__global__ void pack_tc_weights_kernel(const float* __restrict__ src, unsigned short* __restrict__ dst, unsigned short* __restrict__ dst_hi,
                                       int CI, int CI_src, int CO, int ntaps, unsigned perm) {
    const long long n = (long long)CI * CO * ntaps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i % 8);
        const int nn = (int)((i / 8) % TC_NT);
        const int kg = (int)((i / (8 * TC_NT)) % (TC_KCH / 8));
        const int tap = (int)((i / (8 * TC_NT * (TC_KCH / 8))) % ntaps);
        const long long rest = i / ((long long)8 * TC_NT * (TC_KCH / 8) * ntaps);
        const int chunk = (int)(rest % (CI / TC_KCH));
        const int ntile = (int)(rest / (CI / TC_KCH));
        const int ci = chunk * TC_KCH + kg * 8 + e;
        const int co = ntile * TC_NT + nn;
        const int stap = (int)((perm >> (4 * tap)) & 0xF);
        unsigned short hi = 0, lo = 0;
        if (ci < CI_src) split_hl(src[((long long)ci * ntaps + stap) * CO + co], hi, lo);
        const long long blk = ((long long)ntile * (CI / TC_KCH) + chunk) * (2LL * ntaps * (TC_KCH / 8) * TC_NT * 8);
        const long long row0 = (((long long)tap * (TC_KCH / 8) + kg) * (2 * TC_NT)) * 8;
        dst[blk + row0 + (long long)nn * 8 + e] = hi;
        dst[blk + row0 + (long long)(TC_NT + nn) * 8 + e] = lo;
        if (dst_hi != nullptr) dst_hi[blk / 2 + row0 / 2 + (long long)nn * 8 + e] = hi;
    }
}

int launch_pack_tc_weights(const float* src, unsigned short* dst, unsigned short* dst_hi, int CI, int CI_src, int CO, int ntaps,
                           unsigned perm, cudaStream_t stream) {
    long long n = (long long)CI * CO * ntaps;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    pack_tc_weights_kernel<<<blocks, 256, 0, stream>>>(src, dst, dst_hi, CI, CI_src, CO, ntaps, perm);
    MPDB_LAUNCH_CHECK();
    return 0;
}

// trajectory x [B][L][D] (BLC fp32) -> TC layout planes with the channels padded to C (>= D, multiple of 8); the padding
// channels are never written (zero forever). One thread per (b, l): D fp16-split values -> 16-byte stores per k-group.
__global__ void blc_to_tc_kernel(const float* __restrict__ x, unsigned short* __restrict__ hi, unsigned short* __restrict__ lo,
                                 int B, int L, int D, int C) {
    pdl_wait();
    const int Lp = L + 4, SPT = TC_RT / Lp;
    const long long n = (long long)B * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % L), b = (int)(i / L);
        const int tile = b / SPT, s = b - tile * SPT;
        const float* xp = x + i * D;
        for (int kg = 0; kg * 8 < D; ++kg) {
            unsigned short h[8], w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                h[e] = 0; w[e] = 0;
                if (kg * 8 + e < D) split_hl(xp[kg * 8 + e], h[e], w[e]);
            }
            const long long o = (((long long)tile * (C / 8) + kg) * TC_RT + (s * Lp + l + 2)) * 8;
            uint4 ph, pl;
            ph.x = h[0] | ((uint32_t)h[1] << 16); ph.y = h[2] | ((uint32_t)h[3] << 16);
            ph.z = h[4] | ((uint32_t)h[5] << 16); ph.w = h[6] | ((uint32_t)h[7] << 16);
            pl.x = w[0] | ((uint32_t)w[1] << 16); pl.y = w[2] | ((uint32_t)w[3] << 16);
            pl.z = w[4] | ((uint32_t)w[5] << 16); pl.w = w[6] | ((uint32_t)w[7] << 16);
            *reinterpret_cast<uint4*>(hi + o) = ph;
            *reinterpret_cast<uint4*>(lo + o) = pl;
        }
    }
}

int launch_blc_to_tc(const float* x, unsigned short* hi, unsigned short* lo, int B, int L, int D, int C, cudaStream_t stream) {
    long long n = (long long)B * L;
    int blocks = (int)((n + 127) / 128);
    MPDB_CHECK_CUDA(launch_kernel(blc_to_tc_kernel, dim3(blocks), dim3(128), 0, stream, x, hi, lo, B, L, D, C));
    MPDB_LAUNCH_CHECK();
    return 0;
}

// activation fp32 CM [B][C][L+4] -> TC layout planes (used by tests and for inputs produced outside the TC path)
__global__ void cm_to_tc_kernel(const float* __restrict__ cm, unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int B,
                                int C, int L) {
    const int Lp = L + 4, SPT = TC_RT / Lp;
    const long long n = (long long)B * C * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % L);
        const int c = (int)((i / L) % C);
        const int b = (int)(i / ((long long)L * C));
        unsigned short h, lw;
        split_hl(cm[((long long)b * C + c) * Lp + 2 + l], h, lw);
        const int tile = b / SPT, s = b - tile * SPT;
        const long long o = (((long long)tile * (C / 8) + c / 8) * TC_RT + (s * Lp + l + 2)) * 8 + (c % 8);
        hi[o] = h;
        lo[o] = lw;
    }
}

int launch_cm_to_tc(const float* cm, unsigned short* hi, unsigned short* lo, int B, int C, int L, cudaStream_t stream) {
    long long n = (long long)B * C * L;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    cm_to_tc_kernel<<<blocks, 256, 0, stream>>>(cm, hi, lo, B, C, L);
    MPDB_LAUNCH_CHECK();
    return 0;
}

}  // namespace mpdb
