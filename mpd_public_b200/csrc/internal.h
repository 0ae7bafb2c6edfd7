// Internal (non-ABI) declarations shared between the translation units of libmpdb200.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; cuTensorMapEncodeTiled is resolved at run time through the runtime API)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpdb200.h"

namespace mpdb {

// Activation layout "CM": act[b][c][Lp], Lp = L + 4, element (b,c,l) at ((b*C + c)*Lp + l + 2); the two
// halo columns on each side stay zero for the lifetime of the buffer (they are the conv zero padding).
// Layout "BLC": the reference's [B][L][C] (trajectory tensors x, eps).
constexpr int HALO = 2;

struct ConvSrc {
    const float* p0;  // first source
    int c0;
    const float* p1;  // second source (channel concat), may be null
    int c1;
    int blc;  // p0 is BLC (raw trajectory); p1 must be null
    int L;    // positions per sample of the source
};

enum ConvMode { MODE_CONV5 = 0, MODE_CONV1 = 1, MODE_DOWN = 2, MODE_UP = 3, MODE_INPUT = 7 };

struct ConvArgs {
    ConvSrc in;
    const float* w;     // packed [ci][taps][CO]
    const float* bias;  // [CO]
    const float* gamma; // GroupNorm affine, null -> no GroupNorm/Mish
    const float* beta;
    const float* cond;  // [T][CO] time-conditioning table added after Mish, null -> none
    const long long* t_dev;
    int t_uniform;
    ConvSrc res;        // residual source (res.p0 == null -> none)
    const float* res_w; // packed [ci][1][CO]; null -> identity residual
    const float* res_bias;
    float* out;         // CM layout [B][CO][L_out + 4]
    int CO;
    int L_out;
    int B;
    int gs;  // channels per GroupNorm group
    int S;   // samples per CTA
    int NT;  // output channels per CTA
    int G;   // intra-CTA split-K groups (1, 2 or 4)
    // optional second copy of the output in the tensor-core ("TC") layout, fp16 hi / scaled-lo planes (unet_tc.cu)
    unsigned short* out_hi;
    unsigned short* out_lo;
};

int launch_conv(int mode, ConvArgs a, cudaStream_t stream);
void choose_tile(int mode, ConvArgs* a);

// ---- tensor-core path (unet_tc.cu) ----
constexpr int TC_RT = 132;   // rows of one tile block in the TC layout (128 MMA rows + 4 rows of tap reach)
constexpr int TC_NT = 32;    // output channels per CTA
constexpr int TC_KCH = 32;   // input channels per pipeline stage
enum TcMode { TCM_CONV5 = 0, TCM_DOWN = 1, TCM_UP = 2 };

struct TcConvArgs {
    // main conv input: up to two concatenated sources in TC layout (fp16 hi / scaled-lo planes)
    const unsigned short *in0_hi, *in0_lo, *in1_hi, *in1_lo;
    int c0, c1;
    const unsigned short* w;      // packed [CO/32][CI/32][taps][4][hi 32 | lo 32][8]
    const unsigned short *w_hi, *res_w_hi;  // the hi halves alone ([..][4][32][8]): what precision-1 steps stream
    const float* bias;
    const float* gamma;
    const float* beta;
    const float* cond;
    const long long* t_dev;
    int t_uniform;
    // residual: identity (res_cm: fp32 CM tensor with CO channels) or fused 1x1 conv (res_w != null) of r0|r1
    const float* res_cm;
    const unsigned short *r0_hi, *r0_lo, *r1_hi, *r1_lo;
    int rc0, rc1;
    const unsigned short* res_w;  // packed [CO/32][RC/32][hi|lo][1][4][32][8]
    const float* res_bias;
    float* out_cm;                // fp32 CM output (may be null)
    unsigned short *out_hi, *out_lo;  // TC-layout output (may be null)
    float* raw_out;               // debug: raw main accumulator [tile][ntile][128][32]
    long long* dbg;               // debug: 8 clock64 stamps of CTA (0,0) (null = off)
    int CO, L, B, gs;
    int mode;                     // TcMode; L is the INPUT length (DOWN writes L/2 positions, UP writes 2L)
    int prec;                     // 1: one fp16 product per MMA step (hi planes only); otherwise the 22-bit three-product split
    // TMA tensor maps of the four activation sources (in0, in1, r0, r1) for conv5_tc_kernel: 5-D views
    // {8 elements, 132 rows, C/8 k-groups, tiles, 2 planes} of the TC layout whose box brings tm_nch[i] K-chunks of a tile —
    // both planes with the 22-bit split, the hi plane alone in precision 1 — in ONE cp.async.bulk.tensor (UTMALDG)
    CUtensorMap tm[4];
    int tm_nch[4];
    int ps_stages, ps_stage_bytes, ps_act_bytes;  // operand-ring geometry (set by launch_conv5_tc)
    int pdl;                      // launched with programmatic dependent launch (set by launch_conv5_tc): weights before the dependency wait
};
constexpr int TC_PS_MAX_STAGES = 8;
constexpr int TC_PS_RING_BYTES = 216064;  // operand ring of the persistent kernel, cut into stages of (activation box + weight group) bytes by launch_conv5_tc
int make_act_tensor_map(CUtensorMap* out, const unsigned short* hi_plane, long long plane_elems, int C, long long tiles, int nch, int planes);
int launch_conv5_tc(const TcConvArgs& a, cudaStream_t stream);

// Whole ResidualTemporalBlock in one launch (cluster of CO/32 CTAs per row tile, h1 exchanged through DSMEM):
//   h1 = Mish(GN(conv5(in) + b0)) + cond ;  out = Mish(GN(conv5(h1) + b1)) + residual
struct TcRtbArgs {
    TcConvArgs c0;  // first conv: inputs, w, bias, gamma, beta, cond (+t); outputs unused
    TcConvArgs c1;  // second conv: w, bias, gamma, beta, residual (res_cm | r0/r1 + res_w + res_bias), outputs
};
int launch_rtb_tc(const TcRtbArgs& a, cudaStream_t stream);
int launch_pack_tc_weights(const float* src, unsigned short* dst, unsigned short* dst_hi, int CI, int CI_src, int CO, int ntaps,
                           unsigned perm, cudaStream_t stream);
int launch_limits_normalize(const float* x, long long n_rows, int d_in, const float* mins, const float* range, float* out,
                            int d_out, cudaStream_t stream);
int launch_blc_to_tc(const float* x, unsigned short* hi, unsigned short* lo, int B, int L, int D, int C, cudaStream_t stream);
int launch_cm_to_tc(const float* cm, unsigned short* hi, unsigned short* lo, int B, int C, int L, cudaStream_t stream);

// ---- whole-forward persistent cluster kernel (unet_mega.cu) ----
constexpr int MEGA_MAX_LAYERS = 48;
constexpr int MEGA_CLUSTER = 8;  // CTAs per cluster (portable maximum)
constexpr int MEGA_DBG = 16;     // timeline stamps per (layer, CTA rank)
enum MegaType { MG_INPUT = 0, MG_CONV5 = 1, MG_DOWN = 2, MG_UP = 3 };

// One layer of the program. Input geometry: the A buffer holds C_a = 32 * n_a channels of G samples at length L as
// MT row tiles of SPT samples (RT = SPT * (L + 4) rows per k-group); CTA rank -> (row tile rank / NC, channel chunk
// rank % NC). Output geometry (o*) = the input geometry of the next layer.
struct MegaLayer {
    int type;
    int L, Lp, SPT, MT, NC, RT;
    int n_a, n_skip;          // main K-chunks (32 channels each) from the A buffer / from the global skip tensor
    int n_res_a, n_res_skip;  // K-chunks of the block's 1x1 residual conv, issued after the main chunks of conv0
    int a_plane;              // bytes between the hi and lo planes of the A buffer (one fixed offset for the whole program)
    int skip_C, skip_ready;   // channels of the skip tensor; index of the first layer at which it is complete
    int CO, gs;
    int res_mode;             // conv1 of a block: 1 = identity (fp32 values kept in registers), 2 = fused 1x1 conv; else 0
    int oSPT, oLp, oNC, oRT, o_plane;
    int zero_bytes;           // > 0: the sample length changes after this layer; every CTA clears this many bytes of each plane
    int inv_Lp, inv_NC, inv_oSPT, inv_oNC;  // ceil(65536 / d): x / d == (x * inv) >> 16 for the small operands of the kernel's index math
    int tx_in[MEGA_CLUSTER];  // bytes this layer's epilogues deliver into each CTA's A buffer FROM OTHER CTAs (st.async
                              // complete_tx count the CTA's a_full barrier expects for the layer)
    const unsigned short* w;      // packed [NC][n_a + n_skip][taps][4][hi 32 | lo 32][8]
    const unsigned short* res_w;  // packed [NC][n_res_a + n_res_skip][1][4][hi 32 | lo 32][8]
    const unsigned short *w_hi, *res_w_hi;            // hi halves alone (precision-1 steps)
    const unsigned short *skip_hi, *skip_lo;          // [cluster][MT][skip_C/8][RT][8]
    unsigned short *skip_out_hi, *skip_out_lo;        // same layout, CO channels (output also kept as a skip connection)
    const float *bias, *gamma, *beta, *cond, *res_bias;
    float* out_cm;            // final_conv.0: fp32 [B][CO][L+4] for the fused projection kernel
};



struct FinalArgs {
    const float* h;     // CM [B][C][L+4]
    int C;
    const float* w;     // [D][C]
    const float* bias;  // [D]
    const float* x;     // BLC [B][L][D]
    const long long* t_dev;
    int t_uniform;
    // schedule tables (device, [T])
    const float* sr;
    const float* srm1;
    const float* c1;
    const float* c2;
    const float* stdv;
    int predict_epsilon;
    int clip_denoised;
    int mode;           // 0: write eps; 1: write mean; 2: write mean + std*noise*noise_std, hard conds applied;
                        // 3: DDIM step (ddim_sample, diffusion_model_base.py:184-259; eta = 0), hard conds applied;
                        // 4: DDIM step whose result is guided next (out-of-range flag written, no hard conds)
    float ddim_san, ddim_c;  // sqrt(alpha_next), sqrt(1 - alpha_next - sigma^2) of this DDIM step (host fp32, torch op order)
    int ddim_last;           // time_next < 0: the step returns x_start
    const float* noise; // BLC (mode 2)
    float noise_std;
    int n_hc;
    int hc_rows[MPDB_MAX_HARD_CONDS];
    const float* hc_vals;  // [n_hc][B][D]
    float* out;            // BLC
    float* out2;           // optional second copy (chain slot), batch stride below
    long long out2_bstride;
    int* flag_out;         // set to 1 if any |out| > 1 + 1e-4 (mode 1), may be null
    int B, L, D;
};
int launch_final(const FinalArgs& a, cudaStream_t stream);

#ifdef __CUDACC__
// The elementwise update that follows the UNet output e at one element (xv = x_t there, tt = its timestep), shared by
// final_kernel and the cluster kernel's fused epilogue. Same fp32 operation order as the reference, no FMA contraction:
//   modes 1/2  p_mean_variance (diffusion_model_base.py:143-155): x0 = sr*x - srm1*eps ; clamp ; c1*x0 + c2*x
//   modes 3/4  ddim_sample (:228-248, eta = 0): x_start = sr*x - srm1*eps (NOT clamped) ; pred_noise = eps ;
//              x = x_start * sqrt(alpha_next) + c * pred_noise, or x_start itself on the last step
__device__ __forceinline__ float final_update_value(const FinalArgs& F, float e, float xv, int tt) {
    const float sr = F.sr[tt], srm1 = F.srm1[tt];
    if (F.mode >= 3) {
        float x0, pn;
        if (F.predict_epsilon) { x0 = __fsub_rn(__fmul_rn(sr, xv), __fmul_rn(srm1, e)); pn = e; }
        else { x0 = e; pn = __fdiv_rn(__fsub_rn(__fmul_rn(sr, xv), e), srm1); }
        if (F.ddim_last) return x0;
        return __fadd_rn(__fmul_rn(x0, F.ddim_san), __fmul_rn(F.ddim_c, pn));
    }
    float x0 = F.predict_epsilon ? __fsub_rn(__fmul_rn(sr, xv), __fmul_rn(srm1, e)) : e;
    if (F.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    return __fadd_rn(__fmul_rn(F.c1[tt], x0), __fmul_rn(F.c2[tt], xv));
}
#endif

struct MegaProgram {
    MegaLayer layers[MEGA_MAX_LAYERS];
    int n_layers;
    int G;        // samples per cluster
    int B, H, D;
    int t;        // uniform timestep (row of the time-conditioning tables)
    int a_bytes;  // size of the A buffer
    int prec;     // 3: 22-bit fp16-split operands (three products per MMA step); 1: fp16 operands (one product), engine.cu
    const float* x;  // trajectory [B][H][D] fp32
    // final_conv.1 (1x1, C -> D) + DDPM posterior mean [+ noise, hard conditions, chain slot] in the last layer's epilogue
    // (what final_kernel does as a separate launch): the timed loop passes its FinalArgs here
    int fuse_final;
    FinalArgs fin;
    long long* dbg;  // optional timeline: [n_layers][8 ranks][MEGA_DBG] clock64 stamps of cluster dbg_cluster
    int dbg_cluster;
};
size_t mega_smem_bytes(int a_bytes);
int mega_max_active_clusters(int a_bytes);
int launch_unet_mega(const MegaProgram& P, cudaStream_t stream);

int launch_repack_conv(const float* src, float* dst, int CO, int CI, int K, int transposed, cudaStream_t stream);
int launch_time_tables(const float* w1, const float* b1, const float* w3, const float* b3, float* temb_mish, int T,
                       cudaStream_t stream);
int launch_cond_table(const float* w, const float* b, const float* temb_mish, float* table, int T, int CO,
                      cudaStream_t stream);
int launch_cm_to_bcl(const float* cm, float* out, int B, int C, int L, cudaStream_t stream);
int launch_add_noise(float* x, const long long* t_dev, const float* stdv, const float* noise, float noise_std, int B,
                     int HD, cudaStream_t stream);
int launch_copy_hc(float* x, float* out2, long long out2_bstride, int n_hc, const int* hc_rows, const float* hc_vals,
                   int B, int L, int D, cudaStream_t stream);

// guide (guide.cu)
struct GuideStepArgs {
    const float* x_in;   // BLC normalised
    float* x_out;        // x_in + scale * guide(x_in) with hard conds (step mode) or the gradient itself (grad mode)
    int grad_only;
    const int* flag_in;  // batch-global out-of-range flag for x_in (LimitsNormalizer.unnormalize branch)
    int* flag_out;       // flag for x_out, may be null
    const float* model_var;  // [B] or null (scale_grad_by_std)
    float var_uniform;       // used when model_var == null and use_var_uniform
    int use_var_uniform;
    // fused tail of ddpm_sample_fn (last guide iteration of a step)
    const float* noise;  // null -> no noise
    float noise_sd;      // model_std
    float noise_mult;    // noise_std (extra schedule)
    int n_hc;
    int hc_rows[MPDB_MAX_HARD_CONDS];
    const float* hc_vals;
    float* out2;         // chain slot or null
    long long out2_bstride;
    int B, H;
    // several guide evaluations in ONE launch (guide_gradient_steps, sample_functions.py:65-83): the trajectory stays in
    // shared memory between evaluations; the batch-global clip flag of evaluation k+1 is an atomicOr into iter_flags[k+1]
    // followed by a grid barrier on iter_counters[k] (all CTAs must be co-resident: guide_launch_step checks). n_iters <= 1:
    // one evaluation, flag_in / flag_out as before. Noise, chain slot and x_out are written by the last evaluation.
    // position-only state (GuideManagerTrajectories, guides.py:15-146): x_in / x_out hold q columns (normalised positions)
    // and the unnormalised velocity trajectory [B][H][q] lives here; it is read as the velocity half of the state and
    // updated in place (velocity -= sum_c w_c * clip(d cost_c / d velocity)). Position and velocity gradients of a cost are
    // clipped separately. grad_only must be set.
    float* vel_io;
    int n_iters;
    int* iter_flags;            // [n_iters], [0] written by the producer of x_in; zeroed by the caller
    unsigned int* iter_counters;  // [n_iters - 1], zeroed by the caller
    long long* dbg;      // optional clock64 stamps of thread 0 of CTA 0 (MPDB_GUIDE_TIMELINE=1 in mpdb_profile_guide)
    unsigned int* dep_count;  // counts (trajectory, evaluation) pairs whose clamp was decided by OTHER trajectories (set by the launcher)
    int32_t* dec;        // parity instrumentation (mpdb_guide_record_decisions): [n_iters][B][n_costs][n_interp][n_spheres], else null
    int pdl;             // launch with the programmatic-serialization attribute (the loop sets it next to cluster-kernel forwards)
};
int guide_launch_step(mpdb_guide* g, const GuideStepArgs& a, cudaStream_t stream);
int guide_max_coresident(mpdb_guide* g, int H);  // CTAs of the guide kernel that can be resident at once (grid-barrier bound)
int guide_launch_flag(const float* x, long long n, int* flag, cudaStream_t stream);
int guide_device(mpdb_guide* g);
int guide_state_dim(mpdb_guide* g);
bool guide_recording(mpdb_guide* g);  // decisions are being recorded: loops run without CUDA graphs

}  // namespace mpdb
