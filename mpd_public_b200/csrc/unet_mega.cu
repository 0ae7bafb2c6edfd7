// Whole TemporalUnet forward (temporal_unet.py:118-171) as ONE persistent launch of thread-block clusters.
//
// The reverse loop at 100 trajectories per GPU is a chain of ~40 dependent convolutions per forward; run layer by layer
// (unet_tc.cu) every link pays a kernel boundary: launch edge, barrier/TMEM setup, a cold first copy from L2, drain.
// GroupNorm is per sample, so a group of G trajectories can run the whole network with no grid-wide dependency. Here a
// cluster of 8 CTAs owns G (8 at H=64) trajectories from the input projection to final_conv.0 — and, inside the timed loop,
// through final_conv.1 and the DDPM update (fuse_final):
//
//   * activations never leave the cluster: the current tensor lives in every CTA's shared memory ("A buffer") in the
//     tcgen05 no-swizzle K-major operand layout [plane hi|lo][C/8][RT rows][8 x fp16] (the lo plane at one fixed offset,
//     so layers of one length share their zero halo rows and the buffer is cleared only where the length changes); a
//     layer's epilogue writes its 32 output channels straight into the A buffer of every CTA that consumes them through
//     distributed shared memory (st.async), so the next layer's MMAs read local shared memory;
//   * per layer the 8 CTAs tile (row tiles of <=128 padded rows) x (32-channel output chunks): 8x1 at L=64, 3x2 at
//     L=32, 2x4 at L=16, 1x8 at L=8; the (sample -> row tile) map changes at the stride-2 layers and is applied by the
//     writer;
//   * weights stream from L2 through a 3-stage cp.async.bulk ring driven by a dedicated producer warp that runs ahead
//     across layer boundaries (weights do not depend on activations), so a layer's first chunk is already resident
//     when its inputs land; the 1x1 residual conv of a block rides along with conv0's chunks (sixth tap of a stage);
//     skip connections go through global memory (written once, read once, L2-resident);
//   * TMEM, mbarriers and the layer program (a __grid_constant__ table) are set up once per forward;
//   * layer-to-layer synchronisation is two mbarriers per CTA. a_free (cluster scope): every CTA's MMAs of this layer have
//     retired -> its A buffer may be overwritten; one arrival per CTA per layer, sent by an issuer warp (off the epilogue's
//     critical path), active or not, so idle CTAs stay in lock step. a_full (local): the layer's outputs have landed in
//     THIS CTA's A buffer = its own 16 epilogue warps have arrived + the peers' bytes have been counted: remote slices
//     are written with st.async, which performs complete_tx on the destination's barrier when the data is there; the
//     expected byte count per (layer, CTA) is a host-computed constant (MegaLayer::tx_in). No release fence and no
//     remote arrive on the hand-off (MEMBAR.ALL.GPU waited for the acknowledgement of every remote store: ~0.5 us per
//     layer, tools/probes/dsmem_probe.cu). The consumer's issuer warp executes the generic->async proxy fence after its
//     acquire;
//   * the second accumulators (odd rows of the transposed convolutions, the blocks' 1x1 residual convs) stay in TMEM until
//     they are needed — holding them in registers through GroupNorm spilled (96 registers at 608 threads).
//
// Tried and dropped (measured, profiles/README.md): per-source "slice landed" barriers so that a K-chunk's MMAs start as soon
// as its 32 channels have arrived — the incoming DSMEM stores and the tensor core's operand reads share the destination's
// shared-memory port, so the overlap bought nothing (273 -> 309 us per forward); pushing slices with cp.async.bulk
// shared::cta -> shared::cluster (14 B/clk per SM in an 8-way all-to-all, the same network limit).
//
// Arithmetic follows the per-layer tensor-core path (same fp16-split products per K-chunk, fp32 residual values kept in
// registers); the three partial products are summed from separate accumulators, so the two paths agree to the fp16-split
// rounding level (~5e-7 relative with the 22-bit split, tested), and each is deterministic, bit-identical over repeated
// runs and independent of the batch composition (tested).
#include "tc_common.cuh"

namespace mpdb {

constexpr int MG_THREADS = TC_THREADS + 96;  // 16 epilogue warps + 1 producer warp + 2 MMA-issue warps
// One thread issues a tcgen05.mma every ~80-97 cycles whatever its shape (tools/probes/mma_probe.cu); two threads in
// different warps reach the shared-memory operand floor (~44 cycles per MMA at N = 64 / 32). Issuer 0 drives the
// A_hi x [W_hi | W_lo] products, issuer 1 the A_lo x W_hi products, into separate accumulators (deterministic sums).
constexpr int MG_STAGES = 3;
constexpr int MG_TMEM_COLS = 256;  // main: [0,32) hi*hi, [32,64) hi*lo, [64,96) lo*hi; residual conv / odd outputs: the same at +128
constexpr int MG_SCRATCH_BYTES = TC_GN_SCRATCH_BYTES;
// A ring stage = [acts hi | acts lo | 5 taps of weights | the block's 1x1 residual weights of the same 32 input channels]: the
// residual conv rides along with conv0's chunks (same activations, a sixth "tap" into the second accumulator) instead of
// running as chunks of its own, which were bound by the ring's depth / L2 latency (2 MMA pairs per ~340-cycle chunk) and
// loaded the skip activations a second time.
constexpr int MG_STAGE_BYTES = TC_STAGE_BYTES + 2 * TC_B_TAP_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-chunks of weights per ring stage (see the producer loop)
__device__ __forceinline__ int mega_chunks_per_stage(const MegaLayer& Ld, int prec) {
    return (Ld.n_skip == 0 && Ld.n_res_a + Ld.n_res_skip == 0) ? (prec == 1 ? 4 : 2) : 1;
}
__global__ void __launch_bounds__(MG_THREADS, 1) unet_mega_kernel(const __grid_constant__ MegaProgram P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* abuf = smem_raw;                                   // current activation, operand layout
    unsigned char* stages = abuf + P.a_bytes;                         // ring: [acts hi | acts lo | weights] per stage
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + MG_STAGES * MG_STAGE_BYTES);  // full[S], empty[S], acc_done, a_full, a_free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MG_STAGES + 3);
    volatile int* mma_progress = reinterpret_cast<volatile int*>(tmem_slot + 1);
    float* part = reinterpret_cast<float*>(tmem_slot + 4);            // GroupNorm scratch

    long long t_entry;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster = blockIdx.x / MEGA_CLUSTER;
    const uint32_t abuf_u32 = smem_u32(abuf), stages_u32 = smem_u32(stages);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + MG_STAGES);
    const uint32_t acc_done = smem_u32(bars + 2 * MG_STAGES), a_full = acc_done + 8, a_free = acc_done + 16;

    // ---- setup (once per forward) ----
    if (tid == 0) {
        for (int s = 0; s < MG_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 2); }  // both issuers release a stage
        mbar_init(acc_done, 2);
        mbar_init(a_full, TC_THREADS / 32 + 1);  // this CTA's epilogue warps + issuer 0's expect_tx (the peers' st.async bytes), every layer
        mbar_init(a_free, MEGA_CLUSTER);                      // one arrival per CTA, every layer
        *mma_progress = -1;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(MG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < P.a_bytes / 16; i += MG_THREADS) reinterpret_cast<uint4*>(abuf)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();  // peers have initialised their barriers and cleared their A buffers before any remote access
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    // debug timeline: per-cluster {kernel entry, setup done, exit} on the GPU-wide nanosecond timer (layer slots 44..47 of the buffer)
    long long* cdbg = (P.dbg != nullptr && tid == 0 && rank == 0 && cluster < 64) ? P.dbg + (size_t)44 * MEGA_CLUSTER * MEGA_DBG + cluster * 4 : nullptr;
    if (cdbg) { cdbg[0] = t_entry; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cdbg[1])); }

    if (warp == TC_THREADS / 32) {
        // ===== producer warp: every K-chunk of every layer this CTA takes part in, in program order. The whole warp runs
        // the loop (warp-uniform values stay in uniform registers); one elected lane issues the copies. =====
        {
            int i = 0;
            for (int l = 0; l < P.n_layers; ++l) {
                const MegaLayer& Ld = P.layers[l];
                if (Ld.type == MG_INPUT || rank >= Ld.MT * Ld.NC) continue;
                const int mt = rank / Ld.NC, nc = rank - mt * Ld.NC;
                const int ntaps = Ld.type == MG_CONV5 ? 5 : Ld.type == MG_DOWN ? 3 : 4;
                const int n_main = Ld.n_a + Ld.n_skip;
                const bool has_res = Ld.n_res_a + Ld.n_res_skip > 0;  // same chunking as the main conv (engine.cu)
                const uint32_t act_bytes = (uint32_t)(TC_KCH / 8) * Ld.RT * 16;  // one plane of one K-chunk
                bool skip_checked = false;
                // A bulk copy costs ~800 cycles whatever its size (tools/probes/bulk_probe.cu), so the wide layers were bound by the
                // NUMBER of copies, not their bytes. Where a stage needs no room for activations (all K-chunks come from the A
                // buffer, no residual conv riding along) it takes several K-chunks of weights in ONE copy: 2 with the 22-bit
                // split, 4 when only the hi halves are streamed (40 KB either way).
                const int cps = mega_chunks_per_stage(Ld, P.prec);
                for (int c = 0; c < n_main; c += cps, ++i) {
                    const int s = i % MG_STAGES;
                    if (i >= MG_STAGES) mbar_wait(empty0 + 8 * s, ((uint32_t)(i / MG_STAGES) & 1u) ^ 1u);
                    const int cc = c;
                    const int na = Ld.n_a;
                    const bool from_skip = cc >= na;
                    const int nch = n_main - c < cps ? n_main - c : cps;
                    const uint32_t wmul = P.prec == 1 ? 1u : 2u;  // precision 1 streams the hi halves of the weights only
                    const uint32_t wbytes = (uint32_t)nch * ntaps * wmul * TC_B_TAP_BYTES;
                    const uint32_t rbytes = has_res ? wmul * TC_B_TAP_BYTES : 0u;
                    const size_t welems = ((size_t)nc * n_main + cc) * ((size_t)ntaps * 2 * TC_B_TAP_BYTES / 2);
                    const unsigned short* wsrc = P.prec == 1 ? Ld.w_hi + welems / 2 : Ld.w + welems;
                    const uint32_t st = stages_u32 + (uint32_t)s * MG_STAGE_BYTES;
                    if (from_skip && !skip_checked) {
                        // the skip tensor was written (by CTAs of this cluster) many layers ago; make the dependency explicit
                        const long long t0 = clock64();
                        while (*mma_progress < Ld.skip_ready) { if (clock64() - t0 > 4000000000LL) __trap(); }
                        asm volatile("fence.acq_rel.cluster;" ::: "memory");
                        skip_checked = true;
                    }
                    __syncwarp();
                    mbar_expect_tx_elect(full0 + 8 * s, wbytes + rbytes + (from_skip ? (P.prec == 1 ? act_bytes : 2u * act_bytes) : 0u));
                    bulk_g2s_elect(st + (cps > 1 ? 0u : 2u * TC_A_PLANE_BYTES), wsrc, wbytes, full0 + 8 * s);
                    if (has_res)  // the residual conv's weights of this chunk, behind the (at most 5) taps
                        bulk_g2s_elect(st + 2 * TC_A_PLANE_BYTES + 5 * wmul * TC_B_TAP_BYTES,
                                       P.prec == 1 ? Ld.res_w_hi + ((size_t)nc * n_main + cc) * (TC_B_TAP_BYTES / 2)
                                                   : Ld.res_w + ((size_t)nc * n_main + cc) * (2 * TC_B_TAP_BYTES / 2),
                                       rbytes, full0 + 8 * s);
                    if (from_skip) {
                        const size_t aoff = (((size_t)cluster * Ld.MT + mt) * (Ld.skip_C / 8) + (size_t)(cc - na) * (TC_KCH / 8)) * Ld.RT * 8;
                        bulk_g2s_elect(st, Ld.skip_hi + aoff, act_bytes, full0 + 8 * s);
                        if (P.prec != 1) bulk_g2s_elect(st + TC_A_PLANE_BYTES, Ld.skip_lo + aoff, act_bytes, full0 + 8 * s);
                    }
                }
            }
        }
        __syncwarp();
        cluster_sync_all();  // no CTA leaves while peers may still address its shared memory
        return;
    }


    if (warp > TC_THREADS / 32) {
        // ===== two MMA-issue warps. Each runs its loop warp-convergently (descriptors in uniform registers) and one elected
        // lane issues. They write no activations, so their fences never wait on remote stores. =====
        const int which = __shfl_sync(0xffffffffu, warp, 0) - (TC_THREADS / 32 + 1);  // 0: A_hi x [W_hi | W_lo] (N = 64), 1: A_lo x W_hi (N = 32)
        {
            // precision 3 (22-bit operands): issuer 0 drives A_hi x [W_hi | W_lo] (N = 64), issuer 1 A_lo x W_hi (N = 32).
            // precision 1 (fp16 operands, steps where the schedule damps eps errors, engine.cu): only A_hi x W_hi is issued; the
            // two issuers split it by K-group (issuer k takes the k-th 16 channels of every 32-channel chunk) into separate
            // accumulators, so the MMA phase reads 5 KB of operands per (tap, K16) instead of 11 KB and the sum stays ordered.
            const bool p1 = P.prec == 1;
            const uint32_t idesc = (which == 0 && !p1) ? tc_idesc(128, 2 * TC_NT) : tc_idesc(128, TC_NT);
            const uint32_t col0 = __shfl_sync(0xffffffffu, tmem_base, 0) + (which == 0 ? 0u : 2u * TC_NT);  // this issuer's main accumulator
            // descriptor words: low = start address >> 4 | LBO >> 4 << 16 ; high = SBO >> 4 | version 1 << 14
            constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
            const uint32_t b_lo_fixed = (((p1 ? 1u : 2u) * TC_NT * 16u) >> 4) << 16;  // weight tile: LBO = 64 rows x 16 B (W_hi | W_lo), 32 rows when only W_hi is streamed
            int ring_i = 0;
            uint32_t my_acc_ph = 0;
            // a_full phase l = "the outputs of layer l have landed in this CTA's A buffer": the peers' bytes are counted by
            // complete_tx (st.async), announced here before any peer can send them (they send after a_free of layer l, which
            // needs this warp's arrival below)
            if (which == 0 && lane == 0 && P.n_layers > 1) mbar_expect_tx(a_full, (uint32_t)P.layers[0].tx_in[rank] >> (p1 ? 1 : 0));
            __syncwarp();
            if (which == 0 && lane < MEGA_CLUSTER) mbar_arrive_cluster(map_to_cta(a_free, (uint32_t)lane));  // layer 0 reads no A buffer
            for (int l = 1; l < P.n_layers; ++l) {
                const MegaLayer& Ld = P.layers[l];
                // layer parameters first, so that nothing but the issue itself follows the wait
                const bool active = rank < Ld.MT * Ld.NC;
                const int n_main = Ld.n_a + Ld.n_skip;
                const bool has_res = Ld.n_res_a + Ld.n_res_skip > 0;
                const int n_a = Ld.n_a, type = Ld.type, zero_bytes = Ld.zero_bytes;
                const uint32_t lbo = (uint32_t)Ld.RT * 16;
                const uint32_t lo_plane = (which == 1 && !p1) ? (uint32_t)Ld.a_plane : 0u;
                // the first chunk's weights do not depend on the previous layer: wait for them first
                if (active) mbar_wait(full0 + 8 * (ring_i % MG_STAGES), (uint32_t)(ring_i / MG_STAGES) & 1u);
                mbar_wait_cluster(a_full, (uint32_t)(l - 1) & 1u);  // operands of this layer have landed (cluster-wide)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // peers' generic-proxy stores -> tensor core reads
                tc_fence_after();
                long long* mdbg = (P.dbg != nullptr && cluster == P.dbg_cluster && lane == 0 && which == 0) ? P.dbg + ((size_t)l * MEGA_CLUSTER + rank) * MEGA_DBG : nullptr;
                if (mdbg) mdbg[8] = clock64();
                if (which == 0 && lane == 0) {
                    *mma_progress = l;
                    if (l + 1 < P.n_layers) mbar_expect_tx(a_full, (uint32_t)Ld.tx_in[rank] >> (p1 ? 1 : 0));  // phase l (phase l - 1 is complete)
                }
                if (active) {
                    uint32_t acc0 = 0u, acc1 = 0u;  // accumulate flags of the main and the second (residual / odd) accumulator
                    const int cps = mega_chunks_per_stage(Ld, P.prec);
                    const int ntaps_l = type == MG_CONV5 ? 5 : type == MG_DOWN ? 3 : 4;
                    for (int c = 0; c < n_main; c += cps, ++ring_i) {
                        const int sidx = ring_i % MG_STAGES;
                        const uint32_t st = stages_u32 + (uint32_t)sidx * MG_STAGE_BYTES;
                        const uint32_t kstep_a = (2 * lbo) >> 4, kstep_b = ((p1 ? 1u : 2u) * (2 * TC_NT * 16)) >> 4, tap_b = ((p1 ? 1u : 2u) * TC_B_TAP_BYTES) >> 4;
                        if (c > 0) {
                            mbar_wait(full0 + 8 * sidx, (uint32_t)(ring_i / MG_STAGES) & 1u);
                            tc_fence_after();
                        }
                        if (mdbg && c == 0) mdbg[9] = clock64();
                        const int nch = n_main - c < cps ? n_main - c : cps;
                        for (int u = 0; u < nch; ++u) {
                        const int cc = c + u;
                        const bool from_a = cc < n_a;
                        const uint32_t aaddr = from_a ? abuf_u32 + (uint32_t)cc * (TC_KCH / 8) * lbo + lo_plane
                                                      : st + ((which == 1 && !p1) ? (uint32_t)TC_A_PLANE_BYTES : 0u);
                        const uint32_t a_lo = ((aaddr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16);
                        // weights: behind the activation planes of a single-chunk stage, from the stage's start in a multi-chunk one
                        const uint32_t wst = cps > 1 ? st + (uint32_t)u * (uint32_t)ntaps_l * (p1 ? 1u : 2u) * TC_B_TAP_BYTES : st + 2 * TC_A_PLANE_BYTES;
                        const uint32_t b_lo = ((wst >> 4) & 0x3FFFu) | b_lo_fixed;  // rows [0,32) = W_hi, [32,64) = W_lo
                        if (type == MG_CONV5) {  // taps -2..2 -> row shifts 0..4
#pragma unroll
                            for (int tap = 0; tap < 5; ++tap)
#pragma unroll
                                for (int kk = 0; kk < TC_KCH / 16; ++kk) {
                                    if (p1 && kk != which) continue;
                                    tc_mma_bf16_elect32(col0, a_lo + kk * kstep_a + tap, desc_hi, b_lo + tap * tap_b + kk * kstep_b, desc_hi, idesc, acc0);
                                    acc0 = 1u;
                                }
                            if (has_res) {  // the block's 1x1 residual conv on the same activations: centre row (+2), weights behind the taps
#pragma unroll
                                for (int kk = 0; kk < TC_KCH / 16; ++kk) {
                                    if (p1 && kk != which) continue;
                                    tc_mma_bf16_elect32(col0 + 128, a_lo + kk * kstep_a + 2, desc_hi, b_lo + 5 * tap_b + kk * kstep_b, desc_hi, idesc, acc1);
                                    acc1 = 1u;
                                }
                            }
                        } else if (type == MG_DOWN) {  // k3, pad 1: taps -1..1 -> row shifts 1..3
#pragma unroll
                            for (int tap = 0; tap < 3; ++tap)
#pragma unroll
                                for (int kk = 0; kk < TC_KCH / 16; ++kk) {
                                    if (p1 && kk != which) continue;
                                    tc_mma_bf16_elect32(col0, a_lo + kk * kstep_a + tap + 1, desc_hi, b_lo + tap * tap_b + kk * kstep_b, desc_hi, idesc, acc0);
                                    acc0 = 1u;
                                }
                        } else {  // ConvTranspose k4 s2, packed taps [W1, W3 | W0, W2]: even = W1 x[m] + W3 x[m-1], odd = W0 x[m+1] + W2 x[m]
#pragma unroll
                            for (int tap = 0; tap < 4; ++tap) {
                                const int shift = tap == 0 ? 2 : tap == 1 ? 1 : tap == 2 ? 3 : 2;
#pragma unroll
                                for (int kk = 0; kk < TC_KCH / 16; ++kk) {
                                    if (p1 && kk != which) continue;
                                    if (tap < 2) { tc_mma_bf16_elect32(col0, a_lo + kk * kstep_a + shift, desc_hi, b_lo + tap * tap_b + kk * kstep_b, desc_hi, idesc, acc0); acc0 = 1u; }
                                    else { tc_mma_bf16_elect32(col0 + 128, a_lo + kk * kstep_a + shift, desc_hi, b_lo + tap * tap_b + kk * kstep_b, desc_hi, idesc, acc1); acc1 = 1u; }
                                }
                            }
                        }
                        }  // K-chunks of this stage
                        tc_commit_elect(empty0 + 8 * sidx);  // the stage is free when both issuers' MMAs that read it have retired
                    }
                    tc_commit_elect(acc_done);
                    if (mdbg) mdbg[10] = clock64();
                }
                if (which == 0) {
                    // a_free: this CTA's MMAs of layer l have retired (and, when the layout changes, the epilogue warps have
                    // cleared the A buffer) -> peers may overwrite it. Sent from here, off the epilogue's critical path.
                    if (active) { mbar_wait(acc_done, my_acc_ph); my_acc_ph ^= 1u; }
                    if (zero_bytes > 0) asm volatile("bar.sync 2, %0;" ::"n"(TC_THREADS + 32) : "memory");
                    if (lane < MEGA_CLUSTER) mbar_arrive_cluster(map_to_cta(a_free, (uint32_t)lane));
                }
            }
        }
        __syncwarp();
        cluster_sync_all();
        return;
    }

    // ===== 16 epilogue warps =====
    const int q = warp & 3, cg = warp >> 2;
    const int r = q * 32 + lane;  // padded row of the tile = TMEM lane
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cg * 8;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t acc_ph = 0;
    const bool p1e = P.prec == 1;
    float keep[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // fp32 output of the last residual block owned by this thread

    for (int l = 0; l < P.n_layers; ++l) {
        const MegaLayer& Ld = P.layers[l];
        // outputs of the previous layer have landed everywhere (also keeps idle CTAs in lock step)
        if (l > 0) mbar_wait(a_full, (uint32_t)(l - 1) & 1u);  // these warps read nothing the peers wrote: CTA-scope wait
        if (l == P.n_layers - 1 && tid == 0) pdl_launch_dependents();  // last layer: a programmatic dependent may start where SMs are free
        long long* dbg = (P.dbg != nullptr && cluster == P.dbg_cluster && tid == 0) ? P.dbg + ((size_t)l * MEGA_CLUSTER + rank) * MEGA_DBG : nullptr;
        if (dbg) dbg[0] = clock64();  // inputs landed
        const bool active = rank < Ld.MT * Ld.NC;
        // small-operand divisions by host-computed reciprocals (a generic 32-bit division is ~35 dependent instructions)
        const int mt = (rank * Ld.inv_NC) >> 16, nc = rank - mt * Ld.NC;
        const int Lp = Ld.Lp;
        const int s = (r * Ld.inv_Lp) >> 16, ll = r - s * Lp;
        const int sg = mt * Ld.SPT + s;  // sample within the cluster
        const int b = cluster * P.G + sg;
        // in_tile: this thread owns a real (sample slot, row) of the cluster; slots past the batch end (ragged last cluster)
        // run as all-zero trajectories so that the delivered byte counts are the same constants in every cluster
        const bool in_tile = active && (s < Ld.SPT) && (ll < Ld.L) && (sg < P.G);
        const bool valid = in_tile && (b < P.B);
        const int c8 = nc * TC_NT + cg * 8;
        // delivery geometry of this thread (consumer row tile, first destination, byte offset of its row slot): index math
        // done here, in the shadow of the MMAs, not between the a_free hand-off and the stores
        uint32_t d_off = 0u, d_meta = 0u;
        if (Ld.oNC > 0) {
            const int mt2 = (sg * Ld.inv_oSPT) >> 16, s2 = sg - mt2 * Ld.oSPT;
            const int j0 = (rank + 1) - (((rank + 1) * Ld.inv_oNC) >> 16) * Ld.oNC;  // staggered destination order: the writers of a row tile address different peers
            d_off = (uint32_t)(((c8 / 8) * Ld.oRT + (s2 * Ld.oLp + 2)) * 16);
            d_meta = (uint32_t)(mt2 * Ld.oNC) | ((uint32_t)j0 << 8);
        }
        // skip tensor written by the previous layer: its global stores are made visible at cluster scope here, in the
        // shadow of this layer's MMAs (MEMBAR.ALL.GPU waits for their acknowledgement: ~1 us on the hand-off otherwise);
        // the arrive at the end of this layer publishes it (MegaLayer::skip_ready counts on that)
        if (l > 0 && P.layers[l - 1].skip_out_hi != nullptr) asm volatile("fence.acq_rel.cluster;" ::: "memory");
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float4 pb0 = z4, pb1 = z4;  // bias (also needed when the odd rows of an up-sampling layer are delivered)

        if (Ld.type == MG_INPUT) {
            // trajectory x [B][H][D] fp32 -> channels [cg*8, cg*8+8) of row ll (channels >= D stay zero)
            if (valid) {
                const float* xp = P.x + ((size_t)b * P.H + ll) * P.D;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (cg * 8 + e < P.D) v[e] = xp[cg * 8 + e];
            }
        } else {
            float4 pg0 = z4, pg1 = z4, pe0 = z4, pe1 = z4, pc0 = z4, pc1 = z4, pr0 = z4, pr1 = z4;
            if (active) {
                // parameters are fetched while the MMAs run
                pb0 = *reinterpret_cast<const float4*>(Ld.bias + c8); pb1 = *reinterpret_cast<const float4*>(Ld.bias + c8 + 4);
                if (Ld.type == MG_CONV5) {
                    pg0 = *reinterpret_cast<const float4*>(Ld.gamma + c8); pg1 = *reinterpret_cast<const float4*>(Ld.gamma + c8 + 4);
                    pe0 = *reinterpret_cast<const float4*>(Ld.beta + c8); pe1 = *reinterpret_cast<const float4*>(Ld.beta + c8 + 4);
                    if (Ld.cond != nullptr) {
                        const float* cp = Ld.cond + (size_t)P.t * Ld.CO + c8;
                        pc0 = *reinterpret_cast<const float4*>(cp); pc1 = *reinterpret_cast<const float4*>(cp + 4);
                    }
                    if (Ld.res_mode == 2) {
                        pr0 = *reinterpret_cast<const float4*>(Ld.res_bias + c8); pr1 = *reinterpret_cast<const float4*>(Ld.res_bias + c8 + 4);
                    }
                }
                mbar_wait(acc_done, acc_ph);
                acc_ph ^= 1u;
                if (dbg) dbg[1] = clock64();  // accumulators complete
                __syncwarp();
                tc_fence_after();
                tc_load_acc(taddr, p1e, v);  // hi*hi + (lo*hi + hi*lo) * 2^-11, or the two K-group halves of hi*hi
                // the second accumulator (odd outputs of an up-sampling layer, the block's 1x1 residual conv) stays in TMEM
                // until it is needed: it is not overwritten before the next layer's MMAs, and holding it in registers
                // through GroupNorm costs spills
                if (dbg) dbg[4] = clock64();  // accumulators in registers
            }
            // this CTA no longer reads its A buffer: clear it if the next layer uses another layout (issuer warp 0 then tells the cluster)
            if (Ld.zero_bytes > 0) {
                // generic-proxy stores, like the delivered slices: made visible to the tensor core by the consuming issuer
                // warp's proxy fence after its a_full acquire (no fence here, on the epilogue's critical path)
                const int nz = Ld.zero_bytes / 16;
                uint4* lo_plane = reinterpret_cast<uint4*>(abuf + Ld.a_plane);
                for (int i = tid; i < nz; i += TC_THREADS) {
                    reinterpret_cast<uint4*>(abuf)[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (!p1e) lo_plane[i] = make_uint4(0u, 0u, 0u, 0u);
                }
                asm volatile("bar.arrive 2, %0;" ::"n"(TC_THREADS + 32) : "memory");  // issuer warp 0 sends a_free once all 16 warps are here
            }
            if (dbg) dbg[5] = clock64();

            if (active) {
                v[0] += pb0.x; v[1] += pb0.y; v[2] += pb0.z; v[3] += pb0.w;
                v[4] += pb1.x; v[5] += pb1.y; v[6] += pb1.z; v[7] += pb1.w;
                if (Ld.type == MG_CONV5) {
                    switch (Ld.gs) {
                        case 4: gn_mish8<4, true>(v, in_tile, r, s, cg, tid, Ld.SPT, Lp, Ld.L, part, pg0, pg1, pe0, pe1, dbg ? dbg + 4 : nullptr); break;
                        case 8: gn_mish8<8, true>(v, in_tile, r, s, cg, tid, Ld.SPT, Lp, Ld.L, part, pg0, pg1, pe0, pe1, dbg ? dbg + 4 : nullptr); break;
                        case 16: gn_mish8<16, true>(v, in_tile, r, s, cg, tid, Ld.SPT, Lp, Ld.L, part, pg0, pg1, pe0, pe1, dbg ? dbg + 4 : nullptr); break;
                        default: gn_mish8<32, true>(v, in_tile, r, s, cg, tid, Ld.SPT, Lp, Ld.L, part, pg0, pg1, pe0, pe1, dbg ? dbg + 4 : nullptr); break;
                    }
                    v[0] += pc0.x; v[1] += pc0.y; v[2] += pc0.z; v[3] += pc0.w;
                    v[4] += pc1.x; v[5] += pc1.y; v[6] += pc1.z; v[7] += pc1.w;
                    if (Ld.res_mode == 2) {
                        float rv[8];
                        tc_load_acc(taddr + 128, p1e, rv);
                        v[0] += rv[0] + pr0.x; v[1] += rv[1] + pr0.y; v[2] += rv[2] + pr0.z; v[3] += rv[3] + pr0.w;
                        v[4] += rv[4] + pr1.x; v[5] += rv[5] + pr1.y; v[6] += rv[6] + pr1.z; v[7] += rv[7] + pr1.w;
                    } else if (Ld.res_mode == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += keep[j];
                    }
                    if (Ld.res_mode != 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) keep[j] = v[j];
                    }
                }
            }
        }

        // ===== deliver: every CTA that consumes these channels gets them in its A buffer (operand layout) =====
        if (dbg) dbg[2] = clock64();  // epilogue arithmetic done
        mbar_wait(a_free, (uint32_t)l & 1u);  // orders our stores after the peers' operand reads: no acquire (L1 invalidate) needed
        if (dbg) dbg[6] = clock64();  // every CTA's MMAs of this layer have retired
        if (Ld.oNC > 0) {
            // layer constants into registers: inside the loops below every use would be an indexed constant-bank load
            const int oNC = Ld.oNC, ltype = Ld.type;
            const uint32_t o_plane = (uint32_t)Ld.o_plane;
            const int n_out = ltype == MG_UP ? 2 : 1;
            const bool emit = in_tile && (ltype != MG_DOWN || (ll & 1) == 0);
            const int j0 = (int)(d_meta >> 8);
            const uint32_t cta0 = d_meta & 0xffu;
            unsigned short* const skip_hi = Ld.skip_out_hi;
            unsigned short* const skip_lo = Ld.skip_out_lo;
            for (int k = 0; k < n_out; ++k) {
                if (k == 1) {  // up-sampling: the odd output row comes from the second accumulator (warp-wide TMEM load)
                    tc_load_acc(taddr + 128, p1e, v);
                    v[0] += pb0.x; v[1] += pb0.y; v[2] += pb0.z; v[3] += pb0.w;
                    v[4] += pb1.x; v[5] += pb1.y; v[6] += pb1.z; v[7] += pb1.w;
                }
                if (emit) {
                    const int lo = ltype == MG_DOWN ? (ll >> 1) : ltype == MG_UP ? 2 * ll + k : ll;
                    uint4 ph, pl = make_uint4(0u, 0u, 0u, 0u);
                    if (p1e) ph = pack_hi8(v); else pack_split8(v, ph, pl);
                    const uint32_t off = d_off + (uint32_t)lo * 16u;
                    const uint32_t local_hi = abuf_u32 + off, local_lo = local_hi + o_plane;
                    int j = j0;
                    for (int jj = 0; jj < oNC; ++jj) {
                        const uint32_t cta = cta0 + (uint32_t)j;
                        if ((int)cta == rank) {  // own A buffer: plain shared-memory stores
                            *reinterpret_cast<uint4*>(abuf + off) = ph;
                            if (!p1e) *reinterpret_cast<uint4*>(abuf + off + o_plane) = pl;
                        } else {
                            const uint32_t bar = map_to_cta(a_full, cta);
                            st_async_v4(map_to_cta(local_hi, cta), ph, bar);
                            if (!p1e) st_async_v4(map_to_cta(local_lo, cta), pl, bar);
                        }
                        if (++j == oNC) j = 0;
                    }
                    if (skip_hi != nullptr) {  // skip connection: same-level layout in global memory
                        const size_t o = ((((size_t)cluster * Ld.MT + mt) * (Ld.CO / 8) + c8 / 8) * Ld.RT + (r + 2)) * 8;
                        *reinterpret_cast<uint4*>(skip_hi + o) = ph;
                        if (!p1e) *reinterpret_cast<uint4*>(skip_lo + o) = pl;
                    }
                }
            }
        }
        if (valid) {
            if (Ld.out_cm != nullptr && !P.fuse_final) {  // final_conv.0: fp32 channel-major output for final_kernel
                float* op = Ld.out_cm + ((size_t)b * Ld.CO + c8) * Lp + 2 + ll;
#pragma unroll
                for (int j = 0; j < 8; ++j) op[(size_t)j * Lp] = v[j];
            }
        }
        if (Ld.out_cm != nullptr && P.fuse_final) {
            // ===== final_conv.1 (1x1, C -> D) + DDPM posterior mean [+ noise, hard conditions, chain slot] (final_kernel's
            // arithmetic, diffusion_model_base.py:126-150, sample_functions.py:50-62) on the values still in registers.
            // The A buffer is dead after this layer's MMAs: [0, D*C) holds the projection weights, then one partial dot
            // product per (row, 8-channel group, d). =====
            const FinalArgs& F = P.fin;
            const int D = P.D, C = Ld.CO;
            float* wsm = reinterpret_cast<float*>(abuf);           // [D][C]
            float* psm = wsm + ((D * C + 3) & ~3);                 // [128 rows][4 groups][D]
            if (active) {
                for (int i = tid; i < D * C; i += TC_THREADS) wsm[i] = F.w[i];
                epi_sync();
                for (int d = 0; d < D; ++d) {
                    const float* wp = wsm + d * C + c8;
                    float e = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) e = fmaf(wp[j], v[j], e);
                    psm[(r * 4 + cg) * D + d] = e;
                }
                epi_sync();
                const int tt = P.t;
                const float sd = F.stdv[tt];
                bool viol = false;
                const int n_out = Ld.SPT * Ld.L * D;
                for (int idx = tid; idx < n_out; idx += TC_THREADS) {
                    const int d = idx % D;
                    const int sl = idx / D;
                    const int lq = sl % Ld.L, ss = sl / Ld.L;
                    const int sgq = mt * Ld.SPT + ss;
                    const int bq = cluster * P.G + sgq;
                    if (sgq >= P.G || bq >= P.B) continue;
                    const float* pp = psm + ((ss * Lp + lq) * 4) * D + d;
                    float e = ((pp[0] + pp[D]) + (pp[2 * D] + pp[3 * D])) + F.bias[d];
                    const long long gi = ((long long)bq * Ld.L + lq) * D + d;
                    float res = e;
                    if (F.mode != 0) {
                        const float xv = F.x[gi];
                        res = final_update_value(F, e, xv, tt);
                        if (F.mode == 2 || F.mode == 3) {
                            if (F.mode == 2) {
                                const float nz = (tt == 0) ? 0.f : F.noise[gi];
                                res = __fadd_rn(res, __fmul_rn(__fmul_rn(sd, nz), F.noise_std));
                            }
                            for (int k = 0; k < F.n_hc; ++k)  // later entries win, as in the reference's dict iteration
                                if (F.hc_rows[k] == lq) res = F.hc_vals[((long long)k * P.B + bq) * D + d];
                        } else {
                            viol |= (res > 1.0001f) || (res < -1.0001f);
                        }
                    }
                    F.out[gi] = res;
                    if (F.out2) F.out2[(long long)bq * F.out2_bstride + (long long)lq * D + d] = res;
                }
                if (F.flag_out != nullptr && __any_sync(0xffffffffu, viol) && lane == 0) atomicOr(F.flag_out, 1);
            }
        }
        if (dbg) dbg[7] = clock64();  // stores issued
        // generic-proxy stores -> async proxy of the consumers. The skip tensor (global, read by bulk copies several layers
        // later) is fenced here; the A-buffer slices are fenced by the consuming issuer warp after its acquire (the stores
        // are complete in the destination's shared memory once the cluster-scope release below is observed), which takes a
        // store round trip off every layer hand-off.
        if (Ld.skip_out_hi != nullptr) asm volatile("fence.proxy.async;" ::: "memory");
        if (dbg) dbg[3] = clock64();  // outputs delivered
        tc_fence_before();  // all TMEM reads of this layer precede the hand-off
        __syncwarp();
        // local arrive only: the slices sent to the peers complete on THEIR barriers by themselves (st.async)
        if (lane == 0 && l + 1 < P.n_layers) mbar_arrive_local(a_full);
    }

    // teardown
    if (cdbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cdbg[2]));
    tc_fence_before();
    epi_sync();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(MG_TMEM_COLS) : "memory");
    }
    cluster_sync_all();
}

size_t mega_smem_bytes(int a_bytes) {
    return (size_t)a_bytes + (size_t)MG_STAGES * MG_STAGE_BYTES + (2 * MG_STAGES + 3) * 8 + 16 + MG_SCRATCH_BYTES;
}

// How many clusters of the kernel can be resident at once (they must all fit in one wave for the kernel to pay off:
// a second wave doubles the latency chain, and the per-layer kernels win, profiles/README.md).
int mega_max_active_clusters(int a_bytes) {
    static int cached_bytes = -1, cached = 0;
    if (a_bytes == cached_bytes) return cached;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(MEGA_CLUSTER * 32);
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = mega_smem_bytes(a_bytes);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = MEGA_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaFuncSetAttribute(unet_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n, unet_mega_kernel, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    cached_bytes = a_bytes;
    cached = n;
    return n;
}

int launch_unet_mega(const MegaProgram& P, cudaStream_t stream) {
    MPDB_REQUIRE(P.n_layers > 0 && P.n_layers <= MEGA_MAX_LAYERS, "mega: bad layer count");
    MPDB_REQUIRE(P.a_bytes % 128 == 0, "mega: A buffer size must be a multiple of 128");
    const size_t smem = mega_smem_bytes(P.a_bytes);
    MPDB_REQUIRE(smem <= 227 * 1024, "mega: shared memory budget exceeded");
    static unsigned long long configured = 0ull;
    if (mpdb::first_use_on_device(configured)) {
        MPDB_CHECK_CUDA(cudaFuncSetAttribute(unet_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int n_clusters = (P.B + P.G - 1) / P.G;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_clusters * MEGA_CLUSTER);
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = MEGA_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (g_use_pdl || g_pdl_loop) ? 2 : 1;
    MPDB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, unet_mega_kernel, P));
    MPDB_LAUNCH_CHECK();
    return 0;
}

}  // namespace mpdb
