// Host side of libmpdb200: the engine (TemporalUnet weights in device-native layout, schedule tables,
// activation workspace, launch plan) and the fused reverse-diffusion loop.
//   GaussianDiffusionModel.p_sample_loop      diffusion_model_base.py:158-182
//   ddpm_sample_fn                            sample_functions.py:18-62
//   TemporalUnet.forward                      temporal_unet.py:118-171
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "internal.h"

namespace mpdb {

static thread_local std::string g_error;
std::atomic<long long> g_launch_count{0};
// Programmatic dependent launch is wired through every loop kernel but OFF by default: measured on B200 it shortens
// stream-ordered launches (0.45 -> 0.37 ms per UNet forward) yet is ~2 % slower than plain kernel nodes under CUDA-graph
// replay, which is the production path (profiles/README.md). MPDB_PDL=1 enables it.
bool g_use_pdl = []() { const char* v = getenv("MPDB_PDL"); return v && v[0] == '1'; }();
// Programmatic dependent launch between the persistent per-layer tensor-core kernels only (one CTA per SM, one wave): the next
// layer's CTA starts on an SM as soon as this layer's CTA there has exited — no grid-wide completion + launch latency — and
// runs its prologue (barriers, TMEM, first weight copies) up to griddepcontrol.wait. Kernels with more CTAs than fit at once
// (the guide at 512 trajectories) must NOT trigger their dependents early: the waiting CTAs take the slots their own later
// waves need (measured: PDL everywhere leaves cfg 5 unchanged and costs cfg 4 1 %; per-layer only: see profiles/README.md).
bool g_pdl_layers = []() { const char* v = getenv("MPDB_PDL_LAYERS"); return !(v && v[0] == '0'); }();
// The small-batch loop: cluster-kernel forwards (104 CTAs) and fused guide launches (one CTA per trajectory, one wave) trigger
// their dependents late — in the last layer's epilogue / the last evaluation's update pass — so the next kernel starts on the
// SMs that are idle or free up, runs its prologue and blocks in griddepcontrol.wait.
bool g_pdl_loop = []() { const char* v = getenv("MPDB_PDL_LOOP"); return !(v && v[0] == '0'); }();
void set_error(const std::string& msg) { g_error = msg; }

// TMA tensor map of one activation in the TC layout (plane[tile][C/8][132][8] fp16, hi plane followed by the lo plane at
// `plane_elems`): 5-D {8 elements, 132 rows, C/8 k-groups, tiles, 2 planes}, box = {8, 132, 4 * nch, 1, planes} — nch K-chunks
// of one tile, both planes or the hi plane alone, in one cp.async.bulk.tensor; lands as [plane][k-group][row][8], which is
// the tcgen05 no-swizzle K-major operand layout. cuTensorMapEncodeTiled is fetched through the runtime (no libcuda link).
int make_act_tensor_map(CUtensorMap* out, const unsigned short* hi_plane, long long plane_elems, int C, long long tiles, int nch, int planes) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MPDB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MPDB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[5] = {8, (cuuint64_t)TC_RT, (cuuint64_t)(C / 8), (cuuint64_t)tiles, 2};
    const cuuint64_t strides[4] = {16, (cuuint64_t)TC_RT * 16, (cuuint64_t)(C / 8) * TC_RT * 16, (cuuint64_t)plane_elems * 2};  // bytes, dims 1..4
    const cuuint32_t box[5] = {8, (cuuint32_t)TC_RT, (cuuint32_t)(4 * nch), 1, (cuuint32_t)planes};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<unsigned short*>(hi_plane), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPDB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

static int group_norm_n_groups(int c) {  // reference layers.py:389-395
    if (c < 8) return 1;
    for (int g = 8; g < 18; ++g)
        if (c % g == 0) return g;
    return 1;
}

struct Param {
    std::string name;
    long long numel = 0;
    long long offset = 0;  // floats into raw storage
    bool set = false;
};

struct ActBuf {
    std::string name;
    int C = 0, L = 0;
    bool exclusive = false;  // never shares storage (partially written buffers rely on their zero fill)
    long long offset = 0;  // floats into the workspace, per-sample stride = C * (L + 4)
};

struct PackJob {
    std::string src;
    long long dst;
    int CO, CI, K, transposed;  // K == 0 -> plain copy of CO floats
};

struct ConvOp {
    int mode = MODE_CONV5;
    // sources: buffer ids (-1 = external x in BLC layout, -2 = none)
    int in0 = -2, in1 = -2;
    int res0 = -2, res1 = -2;
    int out = -1;
    long long w = -1, bias = -1, gamma = -1, beta = -1, cond = -1, res_w = -1, res_bias = -1;  // packed offsets
    int CO = 0, L_in = 0, L_out = 0, gs = 4;
    bool gn = false;
    // tensor-core path (unet_tc.cu)
    bool tc_ok = false;
    int cin = 0, res_cin = 0;
    int tc_in0 = -2, tc_res0 = -2;  // buffers the tensor-core path reads instead of the external BLC trajectory
    int cin_tc = 0, res_cin_tc = 0; // input widths on the tensor-core path (trajectory padded to 32 channels)
    long long w_tc = -1, res_w_tc = -1;  // offsets (16-bit elements) into packed_tc
};

}  // namespace mpdb

using namespace mpdb;

struct mpdb_engine {
    mpdb_engine_config cfg;
    int device = 0;
    std::vector<int> dims;  // [D, C1, C2, ...]
    std::map<std::string, Param> params;
    float* raw = nullptr;     // parameters in reference layout
    long long raw_floats = 0;
    float* packed = nullptr;  // parameters / tables in kernel layout
    long long packed_floats = 0;
    std::vector<ActBuf> bufs;
    std::vector<ConvOp> ops;
    std::vector<PackJob> packs;
    std::vector<std::pair<std::string, long long>> cond_jobs;  // cond_mlp prefix -> table offset
    float* work = nullptr;    // activations (fp32, channel-major with halo)
    unsigned short* packed_tc = nullptr;  // fp16-split weights in tensor-core layout
    unsigned short* packed_tc_hi = nullptr;  // their hi halves alone (half the bytes): streamed by precision-1 steps
    long long packed_tc_elems = 0;
    unsigned short* work_tc = nullptr;    // activations in tensor-core layout (fp16 hi / scaled-lo planes)
    std::vector<long long> tc_off, tc_plane;  // per buffer: offset of the hi plane, elements per plane
    std::vector<long long> cm_off;            // per buffer: float offset of its (possibly shared) physical slot
    std::vector<CUtensorMap> tmap3, tmap1;    // per buffer: TMA tensor maps of its TC-layout copy (22-bit split: both planes, 2 K-chunks
    std::vector<int> tm_nch3, tm_nch1;        // per box; precision 1: hi plane, 4 K-chunks), and the K-chunks per box
    std::vector<char> need_cm;                // per buffer: some consumer reads the fp32 channel-major copy (identity residual, final
                                              // projection, a layer without a tensor-core path); otherwise the tensor-core kernels skip writing it
    int fuse_rtb = []() { const char* v = getenv("MPDB_FUSE_RTB"); return v ? atoi(v) : 1; }();  // cluster-fused residual blocks on the tensor-core path (see can_fuse_rtb)
    int sm_count = 148;
    int fuse_max_co = []() { const char* v = getenv("MPDB_FUSE_MAX_CO"); return v ? atoi(v) : 128; }();  // widest fused block (measured: 0 -> 11.85, 32 -> 11.67, 64 -> 11.55, 128 -> 11.47 ms per loop)
    // whole-forward persistent cluster kernel (unet_mega.cu)
    long long generation = 0;  // bumped whenever device buffers are reallocated or an option changes: launches captured
                               // by a caller (torch CUDA graph around mpdb_sample_loop) are stale afterwards
    int use_mega = []() { const char* v = getenv("MPDB_MEGA"); return v ? atoi(v) : 1; }();
    // the guide evaluations of a step in one launch: the trajectory stays in shared memory, the batch-global clip flag is
    // resolved per CTA and only an undecidable CTA waits for the grid (guide.cu); bit-identical to one launch per evaluation
    // (tested). Used whenever the batch is co-resident; otherwise one launch per evaluation.
    int fuse_guide = []() { const char* v = getenv("MPDB_FUSE_GUIDE"); return v ? atoi(v) : 1; }();
    int fuse_final = []() { const char* v = getenv("MPDB_FUSE_FINAL"); return v ? atoi(v) : 1; }();  // projection + DDPM update in the cluster kernel
    bool mega_ok = false;
    std::string mega_why;      // why the configuration cannot run as one launch (falls back to per-layer kernels)
    MegaProgram mega;
    long long* mega_dbg = nullptr;        // optional per-layer timeline (option "mega_timeline")
    unsigned short* mega_skip = nullptr;  // skip connections in the cluster-tiled layout
    int alias_buffers = 1;     // liveness-based reuse of activation storage (0: one buffer per layer, for debugging)
    int n_slots = 0;
    long long* dbg_buf = nullptr;  // optional per-op timeline stamps (option "timeline")
    int timeline = 0;
    int tc_mode = 1;           // 0 = exact fp32 FMA path only, 1 = auto (loop steps below tc_amp_limit), 2 = force
    // steps whose sqrt(1/abar - 1) exceeds this run the exact path. With the 22-bit scaled-fp16 operand split the tensor-core
    // path matches the fp32 FMA path even at t = T-1 (eps amplified 4602x: step error 3.5e-4 vs 4.4e-4 against the oracle,
    // profiles/r01c_precision_tlast.txt), so no step is excluded by default; a finite limit restores the carve-out.
    float tc_amp_limit = 3.0e38f;
    // Per-timestep precision policy of the loop (tensor-core path): a step whose eps-to-mean amplification
    // posterior_mean_coef1[t] * sqrt(1/abar_t - 1) (predict_epsilon) is at most this limit issues ONE fp16 product per MMA
    // step instead of the three of the 22-bit split (unet_mega.cu / unet_tc.cu "precision 1"): an eps error of ~1e-3
    // relative (measured 1.5e-3 .. 1.7e-3 at every t, tests/test_gpu_benched.py) then moves the posterior mean by
    // <= limit * 1.7e-3 = 3.6e-4 — the size of the fp32 rounding of the reference itself at t = T-1 (3.3e-4) and a third of
    // the 1e-3 per-step bar. Exponential schedule, T = 25: 0.01 at t = 0, 0.104 at t = 15, 0.2005 at t = 18, 0.257 at t = 19,
    // 0.34 at t = 20, 1.24 at t = 23, 1095 at t = 24: t <= 18 and the noise-free extra steps run one product. 0 disables.
    bool force_prec3 = false;  // set by run_unet_body for the duration of one forward
    float prec1_amp_limit = []() { const char* v = getenv("MPDB_PREC1_AMP"); return v ? (float)atof(v) : 0.21f; }();
    long long work_floats_per_sample = 0;
    int work_batch = 0;
    long long final_w = -1, final_b = -1;
    int final_in = -1;
    // schedule tables on device: [7][T]
    float* sched = nullptr;
    std::vector<float> sched_host;
    bool sched_set = false;
    bool finalized = false;
    // loop state
    float* xbuf[2] = {nullptr, nullptr};
    int* flags = nullptr;
    int n_flags = 0;
    int loop_batch = 0;
    // CUDA graph cache for the fused loop: one instantiated graph per configuration key (batch, guide handle + config hash,
    // loop parameters, options), a handful kept (a weight sweep alternates between guides; BASELINE config 3 has nine)
    struct LoopGraph { cudaGraphExec_t exec = nullptr; long long kernels = 0; long long last_use = 0; };
    std::map<std::string, LoopGraph> graphs;
    long long graph_clock = 0;
    float* g_noise = nullptr;   // staging owned by the engine (stable addresses for the graph)
    float* g_hc = nullptr;
    float* g_chain = nullptr;
    long long g_noise_floats = 0, g_hc_floats = 0, g_chain_floats = 0;
};

namespace mpdb {

static void drop_graphs(mpdb_engine* e) {
    for (auto& kv : e->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
}

static long long add_param(mpdb_engine* e, const std::string& name, long long numel) {
    Param p;
    p.name = name;
    p.numel = numel;
    p.offset = e->raw_floats;
    e->raw_floats += (numel + 3) / 4 * 4;
    e->params[name] = p;
    return p.offset;
}

static long long alloc_packed(mpdb_engine* e, long long numel) {
    long long off = e->packed_floats;
    e->packed_floats += (numel + 3) / 4 * 4;
    return off;
}

static int add_buf(mpdb_engine* e, const std::string& name, int C, int L) {
    ActBuf b;
    b.name = name;
    b.C = C;
    b.L = L;
    b.offset = e->work_floats_per_sample;
    e->work_floats_per_sample += (long long)C * (L + 2 * HALO);
    e->bufs.push_back(b);
    return (int)e->bufs.size() - 1;
}

struct PlanBuilder {
    mpdb_engine* e;
    std::vector<PackJob> packs;
    std::vector<std::pair<std::string, long long>> cond_jobs;  // cond_mlp prefix -> table offset

    long long conv_w(const std::string& name, int CO, int CI, int K, int transposed = 0) {
        add_param(e, name, (long long)CO * CI * K);
        long long dst = alloc_packed(e, (long long)CO * CI * K);
        packs.push_back({name, dst, CO, CI, K, transposed});
        return dst;
    }
    long long vec(const std::string& name, int n) {
        add_param(e, name, n);
        long long dst = alloc_packed(e, n);
        packs.push_back({name, dst, n, 1, 0, 0});
        return dst;
    }

    // ResidualTemporalBlock (layers.py:323-355) = two launches
    int rtb(const std::string& p, int in0, int in1, int cin, int cout, int L) {
        int T = e->cfg.n_diffusion_steps;
        int h1 = add_buf(e, p + ".blocks.0", cout, L);
        int out = add_buf(e, p, cout, L);
        ConvOp a;
        a.mode = MODE_CONV5;
        a.in0 = in0; a.in1 = in1;
        a.out = h1;
        a.w = conv_w(p + ".blocks.0.block.0.weight", cout, cin, 5);
        a.bias = vec(p + ".blocks.0.block.0.bias", cout);
        a.gamma = vec(p + ".blocks.0.block.2.weight", cout);
        a.beta = vec(p + ".blocks.0.block.2.bias", cout);
        add_param(e, p + ".cond_mlp.1.weight", (long long)cout * 32);
        add_param(e, p + ".cond_mlp.1.bias", cout);
        a.cond = alloc_packed(e, (long long)T * cout);
        cond_jobs.push_back({p + ".cond_mlp.1", a.cond});
        a.CO = cout; a.L_in = L; a.L_out = L; a.gn = true;
        a.gs = cout / group_norm_n_groups(cout);
        e->ops.push_back(a);

        ConvOp b;
        b.mode = MODE_CONV5;
        b.in0 = h1;
        b.out = out;
        b.w = conv_w(p + ".blocks.1.block.0.weight", cout, cout, 5);
        b.bias = vec(p + ".blocks.1.block.0.bias", cout);
        b.gamma = vec(p + ".blocks.1.block.2.weight", cout);
        b.beta = vec(p + ".blocks.1.block.2.bias", cout);
        b.res0 = in0; b.res1 = in1;
        if (cin != cout) {
            b.res_w = conv_w(p + ".residual_conv.weight", cout, cin, 1);
            b.res_bias = vec(p + ".residual_conv.bias", cout);
        }
        b.CO = cout; b.L_in = L; b.L_out = L; b.gn = true;
        b.gs = a.gs;
        e->ops.push_back(b);
        return out;
    }
};

static int build_plan(mpdb_engine* e, PlanBuilder& pb) {
    const mpdb_engine_config& c = e->cfg;
    e->dims.clear();
    e->dims.push_back(c.state_dim);
    for (int i = 0; i < c.n_levels; ++i) e->dims.push_back(c.unet_input_dim * c.dim_mults[i]);
    const int n = c.n_levels;
    int L = c.horizon;

    add_param(e, "time_mlp.encoder.1.weight", 128 * 32);
    add_param(e, "time_mlp.encoder.1.bias", 128);
    add_param(e, "time_mlp.encoder.3.weight", 32 * 128);
    add_param(e, "time_mlp.encoder.3.bias", 32);

    // tensor-core path only: the trajectory [B,H,D] re-laid out as a TC-layout activation padded to 32 channels
    const int c_in_pad = (c.state_dim + TC_KCH - 1) / TC_KCH * TC_KCH;
    const int input_buf = add_buf(e, "input", c_in_pad, L);
    e->bufs[input_buf].exclusive = true;  // channels >= state_dim are never written and must stay zero
    {
        ConvOp in;
        in.mode = MODE_INPUT;
        in.in0 = -1;
        in.out = input_buf;
        in.CO = c_in_pad; in.L_in = L; in.L_out = L;
        e->ops.push_back(in);
    }
    std::vector<int> skips;
    int cur = -1;  // external x (BLC)
    for (int i = 0; i < n; ++i) {
        const int cin = e->dims[i], cout = e->dims[i + 1];
        std::string p = "downs." + std::to_string(i);
        int a = pb.rtb(p + ".0", cur, -2, cin, cout, L);
        int b = pb.rtb(p + ".1", a, -2, cout, cout, L);
        skips.push_back(b);
        if (i < n - 1) {
            ConvOp d;
            d.mode = MODE_DOWN;
            d.in0 = b;
            d.out = add_buf(e, p + ".4", cout, L / 2);
            d.w = pb.conv_w(p + ".4.conv.weight", cout, cout, 3);
            d.bias = pb.vec(p + ".4.conv.bias", cout);
            d.CO = cout; d.L_in = L; d.L_out = L / 2;
            e->ops.push_back(d);
            cur = d.out;
            L /= 2;
        } else {
            cur = b;
        }
    }
    const int mid = e->dims[n];
    cur = pb.rtb("mid_block1", cur, -2, mid, mid, L);
    cur = pb.rtb("mid_block2", cur, -2, mid, mid, L);
    for (int i = 0; i < n - 1; ++i) {
        const int lv = n - 1 - i;
        const int cin = e->dims[lv], cout = e->dims[lv + 1];
        std::string p = "ups." + std::to_string(i);
        int a = pb.rtb(p + ".0", cur, skips[lv], 2 * cout, cin, L);
        int b = pb.rtb(p + ".1", a, -2, cin, cin, L);
        ConvOp u;
        u.mode = MODE_UP;
        u.in0 = b;
        u.out = add_buf(e, p + ".4", cin, L * 2);
        u.w = pb.conv_w(p + ".4.conv.weight", cin, cin, 4, /*transposed=*/1);
        u.bias = pb.vec(p + ".4.conv.bias", cin);
        u.CO = cin; u.L_in = L; u.L_out = L * 2;
        e->ops.push_back(u);
        cur = u.out;
        L *= 2;
    }
    {
        const int C = c.unet_input_dim;
        ConvOp f;
        f.mode = MODE_CONV5;
        f.in0 = cur;
        f.out = add_buf(e, "final_conv.0", C, L);
        f.w = pb.conv_w("final_conv.0.block.0.weight", C, e->dims[1], 5);
        f.bias = pb.vec("final_conv.0.block.0.bias", C);
        f.gamma = pb.vec("final_conv.0.block.2.weight", C);
        f.beta = pb.vec("final_conv.0.block.2.bias", C);
        f.CO = C; f.L_in = L; f.L_out = L; f.gn = true;
        f.gs = C / group_norm_n_groups(C);
        e->ops.push_back(f);
        e->final_in = f.out;
        add_param(e, "final_conv.1.weight", (long long)c.state_dim * C);
        add_param(e, "final_conv.1.bias", c.state_dim);
        e->final_w = alloc_packed(e, (long long)c.state_dim * C);
        e->final_b = alloc_packed(e, c.state_dim);
        pb.packs.push_back({"final_conv.1.weight", e->final_w, c.state_dim * C, 1, 0, 0});
        pb.packs.push_back({"final_conv.1.bias", e->final_b, c.state_dim, 1, 0, 0});
    }
    MPDB_REQUIRE(L == c.horizon, "internal: plan length mismatch");
    e->need_cm.assign(e->bufs.size(), 0);
    // which k=5 layers can run on the tensor cores
    auto chans = [&](int id0, int id1) {
        int cc = 0;
        if (id0 == -1) cc += c.state_dim; else if (id0 >= 0) cc += e->bufs[id0].C;
        if (id1 >= 0) cc += e->bufs[id1].C;
        return cc;
    };
    for (ConvOp& op : e->ops) {
        if (op.mode == MODE_INPUT) continue;
        op.cin = chans(op.in0, op.in1);
        op.res_cin = op.res_w >= 0 ? chans(op.res0, op.res1) : 0;
        // the external trajectory is read through its padded TC-layout copy
        op.tc_in0 = op.in0 == -1 ? input_buf : op.in0;
        op.tc_res0 = op.res0 == -1 ? input_buf : op.res0;
        op.cin_tc = (op.in0 == -1 ? c_in_pad : e->bufs[op.in0].C) + (op.in1 >= 0 ? e->bufs[op.in1].C : 0);
        op.res_cin_tc = op.res_w >= 0 ? (op.res0 == -1 ? c_in_pad : e->bufs[op.res0].C) + (op.res1 >= 0 ? e->bufs[op.res1].C : 0) : 0;
        const bool width_ok = op.cin_tc % TC_KCH == 0 && (op.in1 < 0 || e->bufs[op.tc_in0].C % TC_KCH == 0) && op.CO % TC_NT == 0 &&
                              op.L_in + 4 <= TC_RT;
        bool ok = false;
        int ntaps = 5;
        if (op.mode == MODE_CONV5) {
            const bool res_ok = op.res_w < 0 || (op.res_cin_tc % TC_KCH == 0 && (op.res1 < 0 || e->bufs[op.tc_res0].C % TC_KCH == 0));
            const bool gn_ok = op.gn && (op.gs == 4 || op.gs == 8 || op.gs == 16 || op.gs == 32);
            ok = width_ok && res_ok && gn_ok;
        } else if (op.mode == MODE_DOWN) {
            ok = width_ok; ntaps = 3;
        } else if (op.mode == MODE_UP) {
            ok = width_ok; ntaps = 4;
        }
        op.tc_ok = ok;
        if (op.tc_ok) {
            op.w_tc = e->packed_tc_elems;
            e->packed_tc_elems += 2LL * op.cin_tc * op.CO * ntaps;
            if (op.res_w >= 0) {
                op.res_w_tc = e->packed_tc_elems;
                e->packed_tc_elems += 2LL * op.res_cin_tc * op.CO;
            }
        }
    }
    for (const ConvOp& op : e->ops) {
        if (op.mode == MODE_INPUT) continue;
        if (op.res_w < 0 && op.res0 >= 0) e->need_cm[op.res0] = 1;  // identity residual: fp32 values
        if (!op.tc_ok)
            for (int id : {op.in0, op.in1, op.res0, op.res1})
                if (id >= 0) e->need_cm[id] = 1;
    }
    e->need_cm[e->final_in] = 1;
    return 0;
}

// Activation storage. Every layer output is a logical buffer; physical slots are shared between logical buffers of
// the same shape whose lifetimes do not overlap (greedy, in write order). Same shape => same halo / spare-row
// positions, which are never written, so the zero-padding invariant survives the reuse. Keeping the working set of a
// forward pass at a few tens of MB keeps weights and activations resident in the 126 MB L2.
static void build_mega(mpdb_engine* e, int B);

static int ensure_workspace(mpdb_engine* e, int B) {
    if (B <= e->work_batch) return 0;
    ++e->generation;
    drop_graphs(e);
    MPDB_CHECK_CUDA(cudaDeviceSynchronize());
    const size_t nb = e->bufs.size();
    // liveness: op index of the write, op index of the last read
    std::vector<int> wr(nb, -1), rd(nb, -1);
    for (size_t i = 0; i < e->ops.size(); ++i) {
        const ConvOp& op = e->ops[i];
        const int ins[6] = {op.in0, op.in1, op.res0, op.res1, op.tc_in0, op.res_w >= 0 ? op.tc_res0 : -2};
        for (int id : ins)
            if (id >= 0) rd[id] = (int)i;
        wr[op.out] = (int)i;
    }
    rd[e->final_in] = (int)e->ops.size() + 1;  // read by the fused final kernel
    struct Slot { int C, L, free_after; long long cm_off, tc_off, tc_plane; };
    std::vector<Slot> slots;
    e->cm_off.assign(nb, 0);
    e->tc_off.assign(nb, 0);
    e->tc_plane.assign(nb, 0);
    long long cm_total = 0, tc_total = 0;
    for (size_t k = 0; k < nb; ++k) {
        const ActBuf& b = e->bufs[k];
        int found = -1;
        if (e->alias_buffers && !b.exclusive)
            for (size_t sidx = 0; sidx < slots.size(); ++sidx)
                if (slots[sidx].C == b.C && slots[sidx].L == b.L && slots[sidx].free_after < wr[k]) { found = (int)sidx; break; }
        if (found < 0) {
            Slot sl;
            sl.C = b.C; sl.L = b.L;
            sl.cm_off = cm_total;
            cm_total += (long long)b.C * (b.L + 2 * HALO) * B;
            const int Lp = b.L + 2 * HALO;
            sl.tc_plane = 0; sl.tc_off = tc_total;
            if (Lp <= TC_RT && b.C % 8 == 0) {
                const int SPT = TC_RT / Lp;
                const long long tiles = (B + SPT - 1) / SPT;
                sl.tc_plane = tiles * (b.C / 8) * TC_RT * 8;
                tc_total += 2 * sl.tc_plane;
            }
            slots.push_back(sl);
            found = (int)slots.size() - 1;
        }
        slots[found].free_after = b.exclusive ? (1 << 30) : (rd[k] >= 0 ? rd[k] : wr[k]);
        e->cm_off[k] = slots[found].cm_off;
        e->tc_off[k] = slots[found].tc_off;
        e->tc_plane[k] = slots[found].tc_plane;
    }
    e->n_slots = (int)slots.size();
    if (e->work) cudaFree(e->work);
    e->work = nullptr;
    MPDB_CHECK_CUDA(cudaMalloc(&e->work, sizeof(float) * (size_t)cm_total));
    MPDB_CHECK_CUDA(cudaMemset(e->work, 0, sizeof(float) * (size_t)cm_total));  // halo columns stay zero forever
    if (e->work_tc) cudaFree(e->work_tc);
    e->work_tc = nullptr;
    MPDB_CHECK_CUDA(cudaMalloc(&e->work_tc, sizeof(unsigned short) * (size_t)(tc_total > 0 ? tc_total : 8)));
    MPDB_CHECK_CUDA(cudaMemset(e->work_tc, 0, sizeof(unsigned short) * (size_t)(tc_total > 0 ? tc_total : 8)));
    for (int k = 0; k < 2; ++k) {
        if (e->xbuf[k]) cudaFree(e->xbuf[k]);
        MPDB_CHECK_CUDA(cudaMalloc(&e->xbuf[k], sizeof(float) * (size_t)B * e->cfg.horizon * e->cfg.state_dim));
    }
    e->tmap3.assign(nb, CUtensorMap());
    e->tmap1.assign(nb, CUtensorMap());
    e->tm_nch3.assign(nb, 0);
    e->tm_nch1.assign(nb, 0);
    for (size_t k = 0; k < nb; ++k) {
        const ActBuf& b = e->bufs[k];
        if (e->tc_plane[k] == 0 || b.C % TC_KCH != 0) continue;
        const int Lp = b.L + 2 * HALO, SPT = TC_RT / Lp;
        const long long tiles = (B + SPT - 1) / SPT;
        const int chunks = b.C / TC_KCH;
        // K-chunks per box: ~37 KB stages either way (22-bit split: 1 chunk x 2 planes + 20 KB of weights; precision 1: 2 chunks x hi
        // plane + 2 x 10 KB), so the 211 KB ring is 5 stages deep — the L2 -> shared latency (~2 us) needs the depth more than the
        // copy engine needs fewer, larger copies (profiles/r02_tc_timeline_cfg5.txt)
        e->tm_nch3[k] = 1;
        e->tm_nch1[k] = chunks < 2 ? chunks : 2;
        if (make_act_tensor_map(&e->tmap3[k], e->work_tc + e->tc_off[k], e->tc_plane[k], b.C, tiles, e->tm_nch3[k], 2)) return 1;
        if (make_act_tensor_map(&e->tmap1[k], e->work_tc + e->tc_off[k], e->tc_plane[k], b.C, tiles, e->tm_nch1[k], 1)) return 1;
    }
    e->work_batch = B;
    build_mega(e, B);
    return 0;
}

static const float* buf_ptr(mpdb_engine* e, int id, int /*B_alloc*/) { return e->work + e->cm_off[id]; }

static ConvSrc make_src(mpdb_engine* e, int id0, int id1, const float* x_ext, int L) {
    ConvSrc s;
    memset(&s, 0, sizeof(s));
    s.L = L;
    if (id0 == -1) {
        s.p0 = x_ext; s.c0 = e->cfg.state_dim; s.blc = 1;
    } else if (id0 >= 0) {
        s.p0 = buf_ptr(e, id0, e->work_batch); s.c0 = e->bufs[id0].C;
    }
    if (id1 >= 0) { s.p1 = buf_ptr(e, id1, e->work_batch); s.c1 = e->bufs[id1].C; }
    return s;
}

// 1 (one fp16 product) or 3 (22-bit split) for a forward at uniform timestep t, see prec1_amp_limit
static int step_prec(const mpdb_engine* e, int t) {
    const int T = e->cfg.n_diffusion_steps;
    if (e->force_prec3 || e->tc_mode == 2 || e->prec1_amp_limit <= 0.f || !e->sched_set || t < 0 || t >= T) return 3;  // tc_mode 2 ("force") = full split everywhere
    const float c1 = e->sched_host[2 * (size_t)T + t], srm1 = e->sched_host[1 * (size_t)T + t];
    const float amp = e->cfg.predict_epsilon ? c1 * srm1 : c1;
    return amp <= e->prec1_amp_limit ? 1 : 3;
}

static void fill_tc_args(mpdb_engine* e, const ConvOp& op, const long long* t_dev, int t_uniform, int B, TcConvArgs& a) {
    auto hi = [&](int id) -> unsigned short* { return e->tc_plane[id] ? e->work_tc + e->tc_off[id] : nullptr; };
    auto lo = [&](int id) -> unsigned short* { return e->tc_plane[id] ? e->work_tc + e->tc_off[id] + e->tc_plane[id] : nullptr; };
    memset(&a, 0, sizeof(a));
    a.mode = op.mode == MODE_DOWN ? TCM_DOWN : op.mode == MODE_UP ? TCM_UP : TCM_CONV5;
    a.in0_hi = hi(op.tc_in0); a.in0_lo = lo(op.tc_in0); a.c0 = e->bufs[op.tc_in0].C;
    if (op.in1 >= 0) { a.in1_hi = hi(op.in1); a.in1_lo = lo(op.in1); a.c1 = e->bufs[op.in1].C; }
    a.w = e->packed_tc + op.w_tc;
    a.w_hi = e->packed_tc_hi + op.w_tc / 2;
    a.bias = e->packed + op.bias;
    if (op.gn) { a.gamma = e->packed + op.gamma; a.beta = e->packed + op.beta; }
    if (op.cond >= 0) { a.cond = e->packed + op.cond; a.t_dev = t_dev; a.t_uniform = t_uniform; }
    if (op.res_w >= 0) {
        a.r0_hi = hi(op.tc_res0); a.r0_lo = lo(op.tc_res0); a.rc0 = e->bufs[op.tc_res0].C;
        if (op.res1 >= 0) { a.r1_hi = hi(op.res1); a.r1_lo = lo(op.res1); a.rc1 = e->bufs[op.res1].C; }
        a.res_w = e->packed_tc + op.res_w_tc;
        a.res_w_hi = e->packed_tc_hi + op.res_w_tc / 2;
        a.res_bias = e->packed + op.res_bias;
    } else if (op.res0 >= 0) {
        a.res_cm = buf_ptr(e, op.res0, e->work_batch);
    }
    a.out_cm = (!e->alias_buffers || e->need_cm.empty() || e->need_cm[op.out]) ? const_cast<float*>(buf_ptr(e, op.out, e->work_batch)) : nullptr;
    a.out_hi = hi(op.out); a.out_lo = lo(op.out);
    a.CO = op.CO; a.L = op.L_in; a.B = B; a.gs = op.gs;
    a.prec = t_dev == nullptr ? step_prec(e, t_uniform) : 3;  // per-sample t (per-call entry points): always the full split
    const int src_ids[4] = {op.tc_in0, op.in1, op.res_w >= 0 ? op.tc_res0 : -2, op.res_w >= 0 ? op.res1 : -2};
    for (int i = 0; i < 4; ++i) {
        const int id = src_ids[i];
        if (id < 0 || e->tc_plane[id] == 0) continue;
        a.tm[i] = a.prec == 1 ? e->tmap1[id] : e->tmap3[id];
        a.tm_nch[i] = a.prec == 1 ? e->tm_nch1[id] : e->tm_nch3[id];
    }
}

// ---------------------------------------------------------------------------------------------------
// Layer program of the whole-forward cluster kernel (unet_mega.cu): one MegaLayer per op of the plan.
// ---------------------------------------------------------------------------------------------------
static bool mega_geom(int L, int G, MegaLayer& Ld) {
    Ld.L = L;
    Ld.Lp = L + 4;
    if (Ld.Lp > TC_RT || L % 4 != 0) return false;
    Ld.SPT = TC_RT / Ld.Lp < G ? TC_RT / Ld.Lp : G;
    Ld.MT = (G + Ld.SPT - 1) / Ld.SPT;
    Ld.RT = Ld.SPT * Ld.Lp;
    return Ld.SPT <= 12;
}

static bool try_build_mega(mpdb_engine* e, int B, int G, std::string& why) {
    MegaProgram& P = e->mega;
    memset(&P, 0, sizeof(P));
    P.G = G; P.B = B; P.H = e->cfg.horizon; P.D = e->cfg.state_dim;
    const int n_clusters = (B + G - 1) / G;
    if (e->ops.size() > (size_t)MEGA_MAX_LAYERS) { why = "too many layers"; return false; }
    // buffers read as the second (concatenated) source are skip connections: they go through global memory
    struct Skip { long long off; int C, L, ready; };
    std::map<int, Skip> skips;
    long long skip_elems = 0;
    for (const ConvOp& op : e->ops)
        if (op.in1 >= 0 && !skips.count(op.in1)) {
            const ActBuf& bf = e->bufs[op.in1];
            MegaLayer g; memset(&g, 0, sizeof(g));
            if (bf.C % TC_KCH != 0 || !mega_geom(bf.L, G, g)) { why = "skip tensor shape"; return false; }
            Skip sk; sk.off = skip_elems; sk.C = bf.C; sk.L = bf.L; sk.ready = -1;
            skip_elems += 2LL * n_clusters * g.MT * (bf.C / 8) * g.RT * 8;
            skips[op.in1] = sk;
        }
    if (e->mega_skip) { cudaFree(e->mega_skip); e->mega_skip = nullptr; }
    if (cudaMalloc(&e->mega_skip, sizeof(unsigned short) * (size_t)(skip_elems > 0 ? skip_elems : 8)) != cudaSuccess ||
        cudaMemset(e->mega_skip, 0, sizeof(unsigned short) * (size_t)(skip_elems > 0 ? skip_elems : 8)) != cudaSuccess) {
        why = "skip allocation failed"; cudaGetLastError(); return false;
    }
    int cur = -3, n = 0, a_bytes = 0;
    std::vector<int> out_of_layer;
    for (size_t i = 0; i < e->ops.size(); ++i, ++n) {
        const ConvOp& op = e->ops[i];
        MegaLayer& Ld = P.layers[n];
        if (!mega_geom(op.L_in, G, Ld)) { why = "row tiling"; return false; }
        if (op.mode == MODE_INPUT) {
            if (n != 0) { why = "input layer must be first"; return false; }
            Ld.type = MG_INPUT; Ld.NC = 1; Ld.CO = e->bufs[op.out].C;
            if (Ld.CO != TC_KCH || e->cfg.state_dim > TC_KCH) { why = "state_dim > 32"; return false; }
            if (Ld.MT > MEGA_CLUSTER) { why = "row tiles exceed the cluster"; return false; }
        } else {
            if (!op.tc_ok) { why = "layer not tensor-core capable"; return false; }
            if (op.tc_in0 != cur) { why = "layer chain is not sequential"; return false; }
            Ld.type = op.mode == MODE_DOWN ? MG_DOWN : op.mode == MODE_UP ? MG_UP : MG_CONV5;
            Ld.CO = op.CO; Ld.NC = op.CO / TC_NT; Ld.gs = op.gs;
            if (Ld.MT * Ld.NC > MEGA_CLUSTER) { why = "layer needs more than 8 CTA tiles"; return false; }
            const int Ca = e->bufs[cur].C;
            Ld.n_a = Ca / TC_KCH;
            Ld.a_plane = (Ca / 8) * Ld.RT * 16;
            if (2 * Ld.a_plane > a_bytes) a_bytes = 2 * Ld.a_plane;
            if (op.in1 >= 0) {
                const Skip& sk = skips[op.in1];
                if (sk.ready < 0 || sk.ready > n || sk.L != op.L_in) { why = "skip connection not produced before use"; return false; }
                MegaLayer g; memset(&g, 0, sizeof(g)); mega_geom(sk.L, G, g);
                const long long plane = (long long)n_clusters * g.MT * (sk.C / 8) * g.RT * 8;
                Ld.n_skip = sk.C / TC_KCH; Ld.skip_C = sk.C; Ld.skip_ready = sk.ready;
                Ld.skip_hi = e->mega_skip + sk.off; Ld.skip_lo = e->mega_skip + sk.off + plane;
            }
            Ld.w = e->packed_tc + op.w_tc;
            Ld.w_hi = e->packed_tc_hi + op.w_tc / 2;
            Ld.bias = e->packed + op.bias;
            if (op.gn) { Ld.gamma = e->packed + op.gamma; Ld.beta = e->packed + op.beta; }
            if (op.cond >= 0) Ld.cond = e->packed + op.cond;
            if (op.mode == MODE_CONV5 && op.cond >= 0) {
                // conv0 of a residual block: the block's 1x1 residual conv reads the same inputs, so it is issued here
                if (i + 1 >= e->ops.size()) { why = "dangling block"; return false; }
                const ConvOp& b = e->ops[i + 1];
                if (b.mode != MODE_CONV5 || b.in0 != op.out || b.in1 >= 0 || b.res0 != op.in0 || b.res1 != op.in1) { why = "unexpected block structure"; return false; }
                if (b.res_w >= 0) { Ld.n_res_a = Ld.n_a; Ld.n_res_skip = Ld.n_skip; Ld.res_w = e->packed_tc + b.res_w_tc; Ld.res_w_hi = e->packed_tc_hi + b.res_w_tc / 2; }
            }
            if (op.mode == MODE_CONV5 && op.res0 != -2) {
                if (n < 1 || P.layers[n - 1].type != MG_CONV5 || P.layers[n - 1].cond == nullptr) { why = "conv1 without conv0"; return false; }
                if (op.res_w >= 0) {
                    Ld.res_mode = 2; Ld.res_bias = e->packed + op.res_bias;
                } else {
                    // identity: the thread that owns (row, channels) of this output produced the same element of the
                    // previous block's output -> the fp32 values are still in its registers
                    Ld.res_mode = 1;
                    if (n < 2 || P.layers[n - 2].type != MG_CONV5 || P.layers[n - 2].res_mode == 0 || P.layers[n - 2].CO != Ld.CO ||
                        P.layers[n - 2].L != Ld.L || out_of_layer[n - 2] != op.res0 || op.res1 >= 0) { why = "identity residual across a layout change"; return false; }
                }
            }
            if (skips.count(op.out)) {
                Skip& sk = skips[op.out];
                if (Ld.type != MG_CONV5) { why = "skip produced by a strided layer"; return false; }
                const long long plane = (long long)n_clusters * Ld.MT * (sk.C / 8) * Ld.RT * 8;
                Ld.skip_out_hi = e->mega_skip + sk.off; Ld.skip_out_lo = e->mega_skip + sk.off + plane;
                // The writers fence their skip stores in the shadow of the NEXT layer's MMAs and arrive on their a_full after
                // that; their issuer warps then pass the a_free hand-off of layer n + 2, which every CTA of the cluster
                // observes before its own mma_progress reaches n + 3
                sk.ready = n + 3;
            }
            if (op.out == e->final_in) Ld.out_cm = const_cast<float*>(buf_ptr(e, op.out, e->work_batch));
        }
        out_of_layer.push_back(op.out);
        cur = op.out;
    }
    if (cur != e->final_in) { why = "plan does not end at final_conv.0"; return false; }
    P.n_layers = n;
    for (int k = 0; k + 1 < n; ++k) {
        MegaLayer& Ld = P.layers[k];
        const MegaLayer& nx = P.layers[k + 1];
        Ld.oSPT = nx.SPT; Ld.oLp = nx.Lp; Ld.oNC = nx.NC; Ld.oRT = nx.RT;
        if ((Ld.CO / 8) * nx.RT * 16 != nx.a_plane) { why = "internal: plane mismatch"; return false; }
        const int out_L = Ld.type == MG_DOWN ? Ld.L / 2 : Ld.type == MG_UP ? Ld.L * 2 : Ld.L;
        if (out_L != nx.L) { why = "internal: length mismatch"; return false; }
        // The lo plane sits at one fixed offset (the largest plane of the program), so layers of the same length share their
        // zero halo rows whatever their channel count: the A buffer is cleared only where the LENGTH changes (6 layers, not
        // 12), over the largest extent any layer of the following same-length run reads.
        Ld.zero_bytes = 0;
        if (k > 0 && nx.L != Ld.L) {
            int zb = 0;
            for (int m = k + 1; m < n && P.layers[m].L == nx.L; ++m) zb = std::max(zb, P.layers[m].a_plane);
            Ld.zero_bytes = zb;
        }
        // remote bytes per destination CTA (mirrors the delivery loop of unet_mega_kernel): producer CTA p = (row tile, chunk)
        // sends, per sample of its tile and output row, 4 column groups x (16 B hi + 16 B lo) to every CTA of the consuming
        // row tile; its own share is a plain local store and is not counted
        for (int p = 0; p < Ld.MT * Ld.NC; ++p) {
            const int mt = p / Ld.NC;
            for (int s = 0; s < Ld.SPT; ++s) {
                const int sg = mt * Ld.SPT + s;
                if (sg >= G) continue;
                const int mt2 = sg / Ld.oSPT;
                for (int j = 0; j < Ld.oNC; ++j) {
                    const int cta = mt2 * Ld.oNC + j;
                    if (cta >= MEGA_CLUSTER) { why = "internal: consumer tile outside the cluster"; return false; }
                    if (cta != p) Ld.tx_in[cta] += out_L * 4 * 32;
                }
            }
        }
        for (int c = 0; c < MEGA_CLUSTER; ++c)
            if (Ld.tx_in[c] >= (1 << 20)) { why = "internal: mbarrier tx-count range"; return false; }
    }
    for (int k = 0; k < n; ++k) {  // after the extents have been used above: every layer addresses the lo plane at the same offset
        P.layers[k].a_plane = a_bytes / 2;
        P.layers[k].o_plane = a_bytes / 2;
    }
    if ((a_bytes / 2) % 16 != 0) { why = "internal: plane alignment"; return false; }
    for (int k = 0; k < n; ++k) {
        MegaLayer& Ld = P.layers[k];
        auto inv = [](int d) { return d > 0 ? (65536 + d - 1) / d : 0; };  // exact for x < 512 and d < 512 (x = row, rank, sample index)
        Ld.inv_Lp = inv(Ld.Lp); Ld.inv_NC = inv(Ld.NC); Ld.inv_oSPT = inv(Ld.oSPT); Ld.inv_oNC = inv(Ld.oNC);
    }
    P.a_bytes = (a_bytes + 127) / 128 * 128;
    if (mega_smem_bytes(P.a_bytes) > 227 * 1024) { why = "shared memory budget"; return false; }
    return true;
}

static void build_mega(mpdb_engine* e, int B) {
    // structural build for the workspace batch (skip tensors are sized for it); whether a given launch uses the program is
    // decided per batch by mega_usable()
    e->mega_ok = false;
    static const int forced = []() { const char* v = getenv("MPDB_MEGA_G"); return v ? atoi(v) : 0; }();
    const int Gs[8] = {8, 7, 6, 5, 4, 3, 2, 1};
    for (int G : Gs) {
        if (forced > 0 && G != forced) continue;
        if (forced <= 0 && G != 8 && G != 4 && G != 2 && G != 1) continue;
        if (!try_build_mega(e, B, G, e->mega_why)) continue;
        e->mega_ok = true;
        e->mega_why.clear();
        return;
    }
}

// One wave only: with more clusters than can be resident at once the latency chain runs twice and the per-layer kernels are
// faster (measured: 128 trajectories = 16 clusters = 2 waves = 543 us vs 358 us). use_mega = 2 overrides.
static bool mega_usable(mpdb_engine* e, int B, std::string* why = nullptr) {
    if (!e->mega_ok) { if (why) *why = e->mega_why; return false; }
    if (!e->use_mega) { if (why) *why = "option mega = 0"; return false; }
    if (!e->alias_buffers) { if (why) *why = "option alias_buffers = 0 keeps per-layer buffers"; return false; }
    const int G = e->mega.G;
    const int n_clusters = (B + G - 1) / G;
    const int max_clusters = mega_max_active_clusters(e->mega.a_bytes);
    if (e->use_mega != 2 && n_clusters > max_clusters) {
        if (why) *why = "batch needs " + std::to_string(n_clusters) + " clusters of " + std::to_string(G) + " trajectories, " +
                        std::to_string(max_clusters) + " fit in one wave";
        return false;
    }
    if (why) why->clear();
    return true;
}

// Can ops[i], ops[i+1] (the two Conv1dBlocks of a ResidualTemporalBlock) run as one cluster-fused launch?
// fuse_rtb = 1 (default): only while the block's CTAs fit one wave — beyond that the persistent per-layer kernel (operand
// prefetch across work items, double-buffered accumulators) beats one cluster launch per block; 2: always; 0: never.
static bool can_fuse_rtb(mpdb_engine* e, size_t i, bool tc, int B) {
    if (!tc || !e->fuse_rtb || !e->alias_buffers || e->timeline || i + 1 >= e->ops.size()) return false;
    const ConvOp& a = e->ops[i];
    const ConvOp& b = e->ops[i + 1];
    if (e->fuse_rtb == 1) {
        const int SPT = TC_RT / (a.L_in + 4);
        if (SPT < 1 || (long long)((B + SPT - 1) / SPT) * (a.CO / TC_NT) > e->sm_count) return false;
    }
    return a.mode == MODE_CONV5 && b.mode == MODE_CONV5 && a.tc_ok && b.tc_ok && a.cond >= 0 && b.in0 == a.out && b.in1 < 0 &&
           a.res0 == -2 && a.CO == b.CO && a.CO <= 128 && a.CO <= e->fuse_max_co && a.L_in == b.L_in;
}

static int launch_rtb(mpdb_engine* e, size_t i, const long long* t_dev, int t_uniform, int B, cudaStream_t st) {
    TcRtbArgs r;
    fill_tc_args(e, e->ops[i], t_dev, t_uniform, B, r.c0);
    fill_tc_args(e, e->ops[i + 1], t_dev, t_uniform, B, r.c1);
    return launch_rtb_tc(r, st);
}

static int launch_op(mpdb_engine* e, const ConvOp& op, const float* x, const long long* t_dev, int t_uniform, int B,
                     cudaStream_t st, bool tc) {
    auto hi = [&](int id) -> unsigned short* { return e->tc_plane[id] ? e->work_tc + e->tc_off[id] : nullptr; };
    auto lo = [&](int id) -> unsigned short* { return e->tc_plane[id] ? e->work_tc + e->tc_off[id] + e->tc_plane[id] : nullptr; };
    if (op.mode == MODE_INPUT) {
        if (!tc) return 0;  // the exact path reads the trajectory directly
        return launch_blc_to_tc(x, hi(op.out), lo(op.out), B, e->cfg.horizon, e->cfg.state_dim, e->bufs[op.out].C, st);
    }
    if (tc && op.tc_ok) {
        TcConvArgs a;
        fill_tc_args(e, op, t_dev, t_uniform, B, a);
        if (e->timeline && e->dbg_buf) a.dbg = e->dbg_buf + (&op - e->ops.data()) * 16;
        return launch_conv5_tc(a, st);
    }
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.in = make_src(e, op.in0, op.in1, x, op.L_in);
    a.w = e->packed + op.w;
    a.bias = e->packed + op.bias;
    if (op.gn) { a.gamma = e->packed + op.gamma; a.beta = e->packed + op.beta; }
    if (op.cond >= 0) { a.cond = e->packed + op.cond; a.t_dev = t_dev; a.t_uniform = t_uniform; }
    if (op.res0 != -2) {
        a.res = make_src(e, op.res0, op.res1, x, op.L_out);
        if (op.res_w >= 0) { a.res_w = e->packed + op.res_w; a.res_bias = e->packed + op.res_bias; }
    }
    a.out = const_cast<float*>(buf_ptr(e, op.out, e->work_batch));
    if (tc) { a.out_hi = hi(op.out); a.out_lo = lo(op.out); }  // consumers on the tensor-core path read this copy
    a.CO = op.CO; a.L_out = op.L_out; a.B = B; a.gs = op.gs;
    choose_tile(op.mode, &a);
    return launch_conv(op.mode, a, st);
}

// Runs every layer up to (and including) final_conv.0; the 1x1 projection is fused into launch_final.
// `fin` (optional): the projection + DDPM update that follows the body. When the body runs as the cluster kernel it is
// executed in that kernel's last epilogue and *fused is set; otherwise the caller launches final_kernel.
static int run_unet_body(mpdb_engine* e, const float* x, const long long* t_dev, int t_uniform, int B, cudaStream_t st,
                         bool tc, const FinalArgs* fin = nullptr, bool* fused = nullptr, bool ddim = false) {
    e->force_prec3 = ddim;  // DDIM steps have no clamp / posterior damping: always the full 22-bit split
    if (fused) *fused = false;
    if (tc && t_dev == nullptr && !e->timeline && mega_usable(e, B)) {
        MegaProgram P = e->mega;  // one launch: every layer up to final_conv.0 inside thread-block clusters
        P.x = x; P.t = t_uniform; P.B = B;
        P.prec = step_prec(e, t_uniform);
        P.dbg = e->mega_dbg;
        static const int dbg_cluster = []() { const char* v = getenv("MPDB_MEGA_DBG_CLUSTER"); return v ? atoi(v) : 0; }();
        P.dbg_cluster = dbg_cluster;
        const long long fuse_bytes = ((((long long)e->cfg.state_dim * e->cfg.unet_input_dim + 3) & ~3LL) + 128LL * 4 * e->cfg.state_dim) * 4;
        if (fin && fused && e->fuse_final && fin->t_dev == nullptr && fin->x == x && fuse_bytes <= P.a_bytes) {
            P.fuse_final = 1;
            P.fin = *fin;
            *fused = true;
        }
        return launch_unet_mega(P, st);
    }
    for (size_t i = 0; i < e->ops.size(); ++i) {
        if (can_fuse_rtb(e, i, tc, B)) {
            if (launch_rtb(e, i, t_dev, t_uniform, B, st)) return 1;
            ++i;  // the second conv of the block ran inside the fused launch
            continue;
        }
        if (launch_op(e, e->ops[i], x, t_dev, t_uniform, B, st, tc)) return 1;
    }
    return 0;
}

static double op_flops(mpdb_engine* e, const ConvOp& op, int B) {
    auto chans = [&](int id0, int id1) {
        int c = 0;
        if (id0 == -1) c += e->cfg.state_dim; else if (id0 >= 0) c += e->bufs[id0].C;
        if (id1 >= 0) c += e->bufs[id1].C;
        return c;
    };
    if (op.mode == MODE_INPUT) return 0.0;
    const int cin = chans(op.in0, op.in1);
    double f;
    if (op.mode == MODE_UP) f = 2.0 * B * op.L_in * op.CO * (double)cin * 4;       // every input feeds 4 taps
    else f = 2.0 * B * op.L_out * op.CO * (double)cin * (op.mode == MODE_CONV5 ? 5 : op.mode == MODE_DOWN ? 3 : 1);
    if (op.res_w >= 0) f += 2.0 * B * op.L_out * op.CO * (double)chans(op.res0, op.res1);
    return f;
}

static void fill_final(mpdb_engine* e, FinalArgs& f, const float* x, const long long* t_dev, int t_uniform, int B) {
    memset(&f, 0, sizeof(f));
    f.h = buf_ptr(e, e->final_in, e->work_batch);
    f.C = e->cfg.unet_input_dim;
    f.w = e->packed + e->final_w;
    f.bias = e->packed + e->final_b;
    f.x = x;
    f.t_dev = t_dev;
    f.t_uniform = t_uniform;
    const int T = e->cfg.n_diffusion_steps;
    f.sr = e->sched + 0 * T; f.srm1 = e->sched + 1 * T; f.c1 = e->sched + 2 * T; f.c2 = e->sched + 3 * T;
    f.stdv = e->sched + 5 * T;
    f.predict_epsilon = e->cfg.predict_epsilon;
    f.clip_denoised = e->cfg.clip_denoised;
    f.B = B; f.L = e->cfg.horizon; f.D = e->cfg.state_dim;
}

}  // namespace mpdb

extern "C" const char* mpdb_last_error(void) { return mpdb::g_error.c_str(); }
extern "C" int mpdb_version(void) { return 100; }
extern "C" int64_t mpdb_launch_count(void) { return mpdb::g_launch_count.load(); }

extern "C" int mpdb_engine_create(const mpdb_engine_config* cfg, int device, mpdb_engine** out) {
    MPDB_REQUIRE(cfg && out, "mpdb_engine_create: null argument");
    MPDB_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= MPDB_MAX_LEVELS, "engine: bad n_levels");
    MPDB_REQUIRE(cfg->state_dim >= 1 && cfg->state_dim <= MPDB_MAX_STATE_DIM, "engine: state_dim must be in [1, 32]");
    MPDB_REQUIRE(cfg->unet_input_dim % 32 == 0 && cfg->unet_input_dim > 0,
                 "engine: unet_input_dim must be a multiple of 32 (GroupNorm group = 4k channels)");
    MPDB_REQUIRE(cfg->horizon > 0 && cfg->horizon % (8 << (cfg->n_levels - 1)) == 0,
                 "engine: horizon must be divisible by 8 * 2^(n_levels-1)");
    MPDB_REQUIRE(cfg->n_diffusion_steps >= 1, "engine: n_diffusion_steps must be >= 1");
    for (int i = 0; i < cfg->n_levels; ++i) MPDB_REQUIRE(cfg->dim_mults[i] >= 1, "engine: bad dim_mults");
    MPDB_ENTER_DEVICE(device);
    std::unique_ptr<mpdb_engine> e(new mpdb_engine());
    e->cfg = *cfg;
    e->device = device;
    MPDB_CHECK_CUDA(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device));
    PlanBuilder pb;
    pb.e = e.get();
    if (build_plan(e.get(), pb)) return 2;
    MPDB_CHECK_CUDA(cudaMalloc(&e->raw, sizeof(float) * (size_t)e->raw_floats));
    MPDB_CHECK_CUDA(cudaMalloc(&e->packed, sizeof(float) * (size_t)(e->packed_floats + 32 * cfg->n_diffusion_steps)));
    MPDB_CHECK_CUDA(cudaMemset(e->packed, 0, sizeof(float) * (size_t)(e->packed_floats + 32 * cfg->n_diffusion_steps)));
    MPDB_CHECK_CUDA(cudaMalloc(&e->sched, sizeof(float) * 7 * (size_t)cfg->n_diffusion_steps));
    MPDB_CHECK_CUDA(cudaMalloc(&e->packed_tc, sizeof(unsigned short) * (size_t)(e->packed_tc_elems > 0 ? e->packed_tc_elems : 8)));
    MPDB_CHECK_CUDA(cudaMalloc(&e->packed_tc_hi, sizeof(unsigned short) * (size_t)(e->packed_tc_elems > 0 ? e->packed_tc_elems / 2 : 8)));
    e->packs = pb.packs;
    e->cond_jobs = pb.cond_jobs;
    if (ensure_workspace(e.get(), cfg->max_batch > 0 ? cfg->max_batch : 1)) return 1;
    *out = e.release();
    return 0;
}

extern "C" void mpdb_engine_destroy(mpdb_engine* e) {
    if (!e) return;
    mpdb::DeviceGuard dg(e->device);
    cudaDeviceSynchronize();
    mpdb::drop_graphs(e);
    cudaFree(e->raw); cudaFree(e->packed); cudaFree(e->work); cudaFree(e->sched);
    cudaFree(e->packed_tc); cudaFree(e->packed_tc_hi); cudaFree(e->work_tc); cudaFree(e->dbg_buf); cudaFree(e->mega_skip); cudaFree(e->mega_dbg);
    cudaFree(e->xbuf[0]); cudaFree(e->xbuf[1]); cudaFree(e->flags);
    cudaFree(e->g_noise); cudaFree(e->g_hc); cudaFree(e->g_chain);
    delete e;
}

extern "C" int mpdb_engine_set_param(mpdb_engine* e, const char* name, const float* dev_ptr, int64_t numel, void* stream) {
    MPDB_REQUIRE(e && name && dev_ptr, "mpdb_engine_set_param: null argument");
    auto it = e->params.find(name);
    MPDB_REQUIRE(it != e->params.end(), std::string("unexpected parameter '") + name + "' for this TemporalUnet configuration");
    MPDB_REQUIRE(it->second.numel == numel, std::string("size mismatch for parameter '") + name + "': expected " +
                                                std::to_string(it->second.numel) + ", got " + std::to_string(numel));
    MPDB_ENTER_DEVICE(e->device);
    MPDB_CHECK_CUDA(cudaMemcpyAsync(e->raw + it->second.offset, dev_ptr, sizeof(float) * (size_t)numel,
                                    cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    it->second.set = true;
    e->finalized = false;
    return 0;
}

extern "C" int mpdb_engine_set_schedule(mpdb_engine* e, const float* sr, const float* srm1, const float* c1,
                                        const float* c2, const float* logvar, const float* stdv, const float* var) {
    MPDB_REQUIRE(e && sr && srm1 && c1 && c2 && logvar && stdv && var, "mpdb_engine_set_schedule: null argument");
    const int T = e->cfg.n_diffusion_steps;
    e->sched_host.resize(7 * (size_t)T);
    const float* src[7] = {sr, srm1, c1, c2, logvar, stdv, var};
    for (int k = 0; k < 7; ++k) memcpy(e->sched_host.data() + (size_t)k * T, src[k], sizeof(float) * T);
    MPDB_ENTER_DEVICE(e->device);
    MPDB_CHECK_CUDA(cudaMemcpy(e->sched, e->sched_host.data(), sizeof(float) * 7 * (size_t)T, cudaMemcpyHostToDevice));
    // captured loops bake schedule values (std, var, the per-step tensor-core decision) into kernel arguments
    drop_graphs(e);
    ++e->generation;
    e->sched_set = true;
    return 0;
}

extern "C" int mpdb_engine_step_precision(mpdb_engine* e, int32_t t) {
    if (!e) return 0;
    e->force_prec3 = false;
    return e->tc_mode == 0 ? 0 : mpdb::step_prec(e, t);
}

extern "C" int mpdb_limits_normalize(const float* x, int64_t n_rows, int32_t d_in, const float* mins, const float* range, float* out,
                                     int32_t d_out, int device, void* stream) {
    MPDB_REQUIRE(x && mins && range && out && n_rows > 0 && d_in > 0 && d_out >= d_in, "mpdb_limits_normalize: bad argument");
    MPDB_ENTER_DEVICE(device);
    return mpdb::launch_limits_normalize(x, (long long)n_rows, d_in, mins, range, out, d_out, (cudaStream_t)stream);
}

extern "C" int64_t mpdb_engine_generation(mpdb_engine* e) { return e ? (int64_t)e->generation : -1; }

extern "C" int mpdb_engine_set_option(mpdb_engine* e, const char* name, double value) {
    MPDB_REQUIRE(e && name, "mpdb_engine_set_option: null argument");
    ++e->generation;
    const std::string n(name);
    if (n == "tc_mode") {
        MPDB_REQUIRE(value == 0 || value == 1 || value == 2, "tc_mode must be 0 (off), 1 (auto) or 2 (force)");
        e->tc_mode = (int)value;
    } else if (n == "fuse_rtb") {
        MPDB_REQUIRE(value == 0 || value == 1 || value == 2, "fuse_rtb must be 0 (never), 1 (while the block fits one wave) or 2 (always)");
        e->fuse_rtb = (int)value;
        drop_graphs(e);
    } else if (n == "mega_timeline") {
        if (value != 0 && !e->mega_dbg) {
            MPDB_CHECK_CUDA(cudaMalloc(&e->mega_dbg, sizeof(long long) * MEGA_DBG * MEGA_CLUSTER * MEGA_MAX_LAYERS));
            MPDB_CHECK_CUDA(cudaMemset(e->mega_dbg, 0, sizeof(long long) * MEGA_DBG * MEGA_CLUSTER * MEGA_MAX_LAYERS));
        } else if (value == 0 && e->mega_dbg) {
            cudaFree(e->mega_dbg); e->mega_dbg = nullptr;
        }
        drop_graphs(e);
    } else if (n == "fuse_guide") {
        e->fuse_guide = value != 0;
        drop_graphs(e);
    } else if (n == "fuse_final") {
        e->fuse_final = value != 0;
        drop_graphs(e);
    } else if (n == "mega") {
        MPDB_REQUIRE(value == 0 || value == 1 || value == 2, "mega must be 0 (off), 1 (when the batch fits one wave) or 2 (always)");
        e->use_mega = (int)value;
        drop_graphs(e);
    } else if (n == "alias_buffers") {
        if (e->alias_buffers != (value != 0)) {
            e->alias_buffers = value != 0;
            e->work_batch = 0;  // re-plan the workspace on the next call
            drop_graphs(e);
        }
    } else if (n == "timeline") {
        e->timeline = value != 0;
        if (e->timeline && !e->dbg_buf) {
            MPDB_CHECK_CUDA(cudaMalloc(&e->dbg_buf, sizeof(long long) * 16 * e->ops.size()));
            MPDB_CHECK_CUDA(cudaMemset(e->dbg_buf, 0, sizeof(long long) * 16 * e->ops.size()));
        }
    } else if (n == "tc_amp_limit") {
        e->tc_amp_limit = (float)value;
    } else if (n == "prec1_amp_limit") {
        MPDB_REQUIRE(value >= 0, "prec1_amp_limit must be >= 0 (0 disables the one-product steps)");
        e->prec1_amp_limit = (float)value;
        drop_graphs(e);
    } else {
        MPDB_REQUIRE(false, "unknown option '" + n + "'");
    }
    return 0;
}

extern "C" int mpdb_engine_finalize(mpdb_engine* e, void* stream) {
    MPDB_REQUIRE(e, "mpdb_engine_finalize: null engine");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    for (auto& kv : e->params)
        MPDB_REQUIRE(kv.second.set, std::string("missing parameter '") + kv.first + "' (load_state_dict incomplete)");
    MPDB_REQUIRE(e->sched_set, "schedule tables not set");
    drop_graphs(e);
    for (const PackJob& j : e->packs) {
        const float* src = e->raw + e->params[j.src].offset;
        float* dst = e->packed + j.dst;
        if (j.K == 0) {
            MPDB_CHECK_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)j.CO, cudaMemcpyDeviceToDevice, st));
        } else {
            if (launch_repack_conv(src, dst, j.CO, j.CI, j.K, j.transposed, st)) return 1;
        }
    }
    for (const ConvOp& op : e->ops) {
        if (!op.tc_ok) continue;
        // destination tap -> source tap: identity for Conv1d; ConvTranspose1d packs [W1, W3 | W0, W2] (even | odd outputs)
        const int ntaps = op.mode == MODE_DOWN ? 3 : op.mode == MODE_UP ? 4 : 5;
        const unsigned perm = op.mode == MODE_UP ? 0x2031u : 0x43210u;
        if (launch_pack_tc_weights(e->packed + op.w, e->packed_tc + op.w_tc, e->packed_tc_hi + op.w_tc / 2, op.cin_tc, op.cin, op.CO, ntaps, perm, st)) return 1;
        if (op.res_w >= 0 && launch_pack_tc_weights(e->packed + op.res_w, e->packed_tc + op.res_w_tc, e->packed_tc_hi + op.res_w_tc / 2,
                                                    op.res_cin_tc, op.res_cin, op.CO, 1, 0u, st))
            return 1;
    }
    const int T = e->cfg.n_diffusion_steps;
    float* temb = e->packed + e->packed_floats;  // [T][32] scratch behind the packed parameters
    if (launch_time_tables(e->raw + e->params["time_mlp.encoder.1.weight"].offset,
                           e->raw + e->params["time_mlp.encoder.1.bias"].offset,
                           e->raw + e->params["time_mlp.encoder.3.weight"].offset,
                           e->raw + e->params["time_mlp.encoder.3.bias"].offset, temb, T, st))
        return 1;
    for (auto& cj : e->cond_jobs) {
        const Param& w = e->params[cj.first + ".weight"];
        const Param& b = e->params[cj.first + ".bias"];
        if (launch_cond_table(e->raw + w.offset, e->raw + b.offset, temb, e->packed + cj.second, T, (int)b.numel, st))
            return 1;
    }
    MPDB_CHECK_CUDA(cudaStreamSynchronize(st));
    e->finalized = true;
    ++e->generation;
    return 0;
}

// Per-call entry points (TemporalUnet.forward, p_mean_variance: per-sample t on the device): tensor cores unless the exact
// path is requested or a finite tc_amp_limit carve-out is configured (t is not known on the host, so it cannot be applied
// per step here).
static bool percall_tc(const mpdb_engine* e) { return e->tc_mode == 2 || (e->tc_mode == 1 && e->tc_amp_limit >= 1.0e30f); }

extern "C" int mpdb_unet_forward(mpdb_engine* e, const float* x, const int64_t* t, float* eps, int32_t B, void* stream) {
    MPDB_REQUIRE(e && x && t && eps && B > 0, "mpdb_unet_forward: bad argument");
    MPDB_REQUIRE(e->finalized, "engine not finalized (call mpdb_engine_finalize after loading parameters)");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    if (ensure_workspace(e, B)) return 1;
    if (run_unet_body(e, x, (const long long*)t, 0, B, st, percall_tc(e))) return 1;
    FinalArgs f;
    fill_final(e, f, x, (const long long*)t, 0, B);
    f.mode = 0;
    f.out = eps;
    return launch_final(f, st);
}

// The forward exactly as the fused loop runs it (one uniform timestep, tensor-core policy of the loop:
// tc_mode 0 -> exact, otherwise tensor cores; whole-forward cluster kernel when enabled and supported).
extern "C" int mpdb_unet_forward_uniform(mpdb_engine* e, const float* x, int32_t t, float* eps, int32_t B, void* stream) {
    MPDB_REQUIRE(e && x && eps && B > 0, "mpdb_unet_forward_uniform: bad argument");
    MPDB_REQUIRE(t >= 0 && t < e->cfg.n_diffusion_steps, "mpdb_unet_forward_uniform: t out of range");
    MPDB_REQUIRE(e->finalized, "engine not finalized (call mpdb_engine_finalize after loading parameters)");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    if (ensure_workspace(e, B)) return 1;
    if (run_unet_body(e, x, nullptr, t, B, st, e->tc_mode != 0)) return 1;
    FinalArgs f;
    fill_final(e, f, x, nullptr, t, B);
    f.mode = 0;
    f.out = eps;
    return launch_final(f, st);
}

// Is the whole-forward cluster kernel in use for batch B? Returns 1/0; fills samples per cluster, layers, A-buffer and
// shared-memory bytes; `why` receives the reason when it is not.
extern "C" int mpdb_engine_mega_info(mpdb_engine* e, int32_t B, int32_t* G, int32_t* n_layers, int32_t* a_bytes,
                                     int32_t* smem_bytes, char* why, int why_cap) {
    if (!e || B <= 0) return 0;
    mpdb::DeviceGuard dg(e->device);
    if (!dg.ok || ensure_workspace(e, B)) return 0;
    if (G) *G = e->mega.G;
    if (n_layers) *n_layers = e->mega.n_layers;
    if (a_bytes) *a_bytes = e->mega.a_bytes;
    if (smem_bytes) *smem_bytes = (int32_t)mega_smem_bytes(e->mega.a_bytes);
    std::string reason;
    const bool ok = mega_usable(e, B, &reason);
    if (why && why_cap > 0) { strncpy(why, reason.c_str(), why_cap - 1); why[why_cap - 1] = 0; }
    return ok ? 1 : 0;
}

// Device time of the UNet body (every layer up to final_conv.0, as the loop runs it) per forward: CUDA events on
// `stream` around `reps` back-to-back forwards. launches_out = kernels per forward.
extern "C" int mpdb_profile_unet_body(mpdb_engine* e, const float* x, int32_t t, int32_t B, int32_t reps, float* ms_out,
                                      double* flops_out, int32_t* launches_out, void* stream) {
    MPDB_REQUIRE(e && x && ms_out && B > 0 && reps > 0, "mpdb_profile_unet_body: bad argument");
    MPDB_REQUIRE(e->finalized, "engine not finalized");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    if (ensure_workspace(e, B)) return 1;
    const bool tc = e->tc_mode != 0;
    cudaEvent_t ev0, ev1;
    MPDB_CHECK_CUDA(cudaEventCreate(&ev0));
    MPDB_CHECK_CUDA(cudaEventCreate(&ev1));
    const long long before = mpdb::g_launch_count.load();
    if (run_unet_body(e, x, nullptr, t, B, st, tc)) return 1;  // warm-up
    const long long per = mpdb::g_launch_count.load() - before;
    MPDB_CHECK_CUDA(cudaEventRecord(ev0, st));
    for (int r = 0; r < reps; ++r)
        if (run_unet_body(e, x, nullptr, t, B, st, tc)) return 1;
    MPDB_CHECK_CUDA(cudaEventRecord(ev1, st));
    MPDB_CHECK_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    MPDB_CHECK_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    *ms_out = ms / reps;
    if (flops_out) {
        double f = 0.0;
        for (const ConvOp& op : e->ops) f += op_flops(e, op, B);
        *flops_out = f;
    }
    if (launches_out) *launches_out = (int32_t)per;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return 0;
}

extern "C" int mpdb_p_mean(mpdb_engine* e, const float* x, const int64_t* t, float* mean, int32_t B, void* stream) {
    MPDB_REQUIRE(e && x && t && mean && B > 0, "mpdb_p_mean: bad argument");
    MPDB_REQUIRE(e->finalized, "engine not finalized (call mpdb_engine_finalize after loading parameters)");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    if (ensure_workspace(e, B)) return 1;
    if (run_unet_body(e, x, (const long long*)t, 0, B, st, percall_tc(e))) return 1;
    FinalArgs f;
    fill_final(e, f, x, (const long long*)t, 0, B);
    f.mode = 1;
    f.out = mean;
    return launch_final(f, st);
}

extern "C" int mpdb_add_noise(mpdb_engine* e, float* x, const int64_t* t, const float* noise, float noise_std,
                              int32_t B, void* stream) {
    MPDB_REQUIRE(e && x && t && noise && B > 0, "mpdb_add_noise: bad argument");
    MPDB_REQUIRE(e->sched_set, "schedule tables not set");
    MPDB_ENTER_DEVICE(e->device);
    const int T = e->cfg.n_diffusion_steps;
    return launch_add_noise(x, (const long long*)t, e->sched + 5 * T, noise, noise_std, B,
                            e->cfg.horizon * e->cfg.state_dim, (cudaStream_t)stream);
}

namespace mpdb {

// Enqueues the whole reverse loop on `st`. All pointers must stay valid until the stream drains.
static int enqueue_loop(mpdb_engine* e, mpdb_guide* g, const mpdb_loop_params* p, const float* noise,
                        const float* hc_vals, float* x_out, float* chain, long long chain_step_stride,
                        long long chain_batch_stride, int B, cudaStream_t st) {
    const int T = e->cfg.n_diffusion_steps, H = e->cfg.horizon, D = e->cfg.state_dim;
    const long long n = (long long)B * H * D;
    const int n_iters = T + p->n_steps_without_noise;
    // clip flags [n_iters][n_guide + 1] + 1, then the grid-barrier counters of the fused guide launches [n_iters][n_guide]
    const int n_flag_ints = n_iters * (p->n_guide_steps + 1) + 1;
    const int n_flags_needed = n_flag_ints + n_iters * (p->n_guide_steps > 0 ? p->n_guide_steps : 1);
    if (g && p->n_guide_steps > 0) {
        MPDB_REQUIRE(e->n_flags >= n_flags_needed, "internal: flag scratch not allocated");
        MPDB_CHECK_CUDA(cudaMemsetAsync(e->flags, 0, sizeof(int) * (size_t)n_flags_needed, st));
    }
    // x_T = noise[0] with hard conditions (diffusion_model_base.py:165-166)
    float* cur = e->xbuf[0];
    MPDB_CHECK_CUDA(cudaMemcpyAsync(cur, noise, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    if (launch_copy_hc(cur, chain, chain_batch_stride, p->n_hard_conds, p->hard_cond_rows, hc_vals, B, H, D, st)) return 1;

    int it = 0;
    for (int i = T - 1; i >= -p->n_steps_without_noise; --i, ++it) {
        const int t = i < 0 ? 0 : i;  // sample_functions.py:28-30
        const bool last = (i == -p->n_steps_without_noise);
        float* nxt = last ? x_out : e->xbuf[(it + 1) & 1];
        float* chain_slot = chain ? chain + (long long)(it + 1) * chain_step_stride : nullptr;
        const float* step_noise = noise + (long long)(it + 1) * n;
        const bool guided = g != nullptr && p->n_guide_steps > 0 && (long long)i < (long long)p->t_start_guide;
        const float ns = p->noise_std ? p->noise_std[it] : 1.0f;

        // condition-aware precision: the fp16-split tensor-core path everywhere except where the schedule
        // amplifies eps beyond tc_amp_limit (no step by default, see tc_amp_limit), which runs the exact fp32 FMA path
        const bool tc = e->tc_mode == 2 || (e->tc_mode == 1 && e->sched_host[1 * (size_t)T + t] <= e->tc_amp_limit);
        // the projection + DDPM update of this step: fused into the cluster kernel's last epilogue when the body runs as
        // one launch, a separate final_kernel launch otherwise
        FinalArgs f;
        fill_final(e, f, cur, nullptr, t, B);
        f.n_hc = p->n_hard_conds;
        for (int k = 0; k < p->n_hard_conds; ++k) f.hc_rows[k] = p->hard_cond_rows[k];
        f.hc_vals = hc_vals;
        int* fl = e->flags + (long long)it * (p->n_guide_steps + 1);
        if (!guided) {
            f.mode = 2;
            f.noise = step_noise;
            f.noise_std = ns;
            f.out = nxt;
            f.out2 = chain_slot;
            f.out2_bstride = chain_batch_stride;
        } else {
            f.mode = 1;
            f.out = nxt;  // model mean; guided in place below
            f.flag_out = fl;
        }
        bool fused = false;
        if (run_unet_body(e, cur, nullptr, t, B, st, tc, &f, &fused)) return 1;
        if (!fused && launch_final(f, st)) return 1;
        if (guided && e->fuse_guide && p->n_guide_steps > 1 && B <= guide_max_coresident(g, H)) {
            // the n_guide_steps evaluations of this step in ONE launch: the trajectory stays in shared memory, the
            // batch-global clip flag goes through a grid barrier (guide.cu)
            GuideStepArgs a;
            memset(&a, 0, sizeof(a));
            a.x_in = nxt;
            a.x_out = nxt;
            a.flag_in = fl;
            a.n_iters = p->n_guide_steps;
            a.iter_flags = fl;
            a.iter_counters = reinterpret_cast<unsigned int*>(e->flags + n_flag_ints + (long long)it * p->n_guide_steps);
            if (p->scale_grad_by_std) { a.use_var_uniform = 1; a.var_uniform = e->sched_host[6 * (size_t)T + t]; }
            a.n_hc = p->n_hard_conds;
            for (int q = 0; q < p->n_hard_conds; ++q) a.hc_rows[q] = p->hard_cond_rows[q];
            a.hc_vals = hc_vals;
            if (t != 0) {  // noise[t == 0] = 0 (sample_functions.py:52)
                a.noise = step_noise;
                a.noise_sd = e->sched_host[5 * (size_t)T + t];
                a.noise_mult = ns;
            }
            a.out2 = chain_slot;
            a.out2_bstride = chain_batch_stride;
            a.B = B;
            a.H = H;
            a.pdl = (g_pdl_loop && fused) ? 1 : 0;  // next to cluster-kernel forwards the launch is a single wave
            if (guide_launch_step(g, a, st)) return 1;
        } else if (guided) {
            for (int k = 0; k < p->n_guide_steps; ++k) {
                const bool klast = (k == p->n_guide_steps - 1);
                GuideStepArgs a;
                memset(&a, 0, sizeof(a));
                a.x_in = nxt;
                a.x_out = nxt;
                a.flag_in = fl + k;
                a.flag_out = klast ? nullptr : fl + k + 1;
                if (p->scale_grad_by_std) { a.use_var_uniform = 1; a.var_uniform = e->sched_host[6 * (size_t)T + t]; }
                a.n_hc = p->n_hard_conds;
                for (int q = 0; q < p->n_hard_conds; ++q) a.hc_rows[q] = p->hard_cond_rows[q];
                a.hc_vals = hc_vals;
                if (klast) {
                    if (t != 0) {  // noise[t == 0] = 0 (sample_functions.py:52)
                        a.noise = step_noise;
                        a.noise_sd = e->sched_host[5 * (size_t)T + t];
                        a.noise_mult = ns;
                    }
                    a.out2 = chain_slot;
                    a.out2_bstride = chain_batch_stride;
                }
                a.B = B;
                a.H = H;
                if (guide_launch_step(g, a, st)) return 1;
            }
        }
        cur = nxt;
    }
    return 0;
}

static int ensure_flags(mpdb_engine* e, int n_flags_needed) {
    if (e->n_flags >= n_flags_needed) return 0;
    drop_graphs(e);
    MPDB_CHECK_CUDA(cudaDeviceSynchronize());
    if (e->flags) cudaFree(e->flags);
    e->flags = nullptr;
    MPDB_CHECK_CUDA(cudaMalloc(&e->flags, sizeof(int) * (size_t)n_flags_needed));
    e->n_flags = n_flags_needed;
    ++e->generation;
    return 0;
}

static int ensure_staging(float** buf, long long* have, long long need) {
    if (need <= *have) return 0;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    MPDB_CHECK_CUDA(cudaMalloc(buf, sizeof(float) * (size_t)need));
    *have = need;
    return 0;
}

}  // namespace mpdb

extern "C" int mpdb_sample_loop(mpdb_engine* e, mpdb_guide* g, const mpdb_loop_params* p, const float* noise,
                                float* x_out, float* chain_out, int64_t chain_step_stride, int64_t chain_batch_stride,
                                int32_t B, void* stream) {
    MPDB_REQUIRE(e && p && noise && x_out && B > 0, "mpdb_sample_loop: bad argument");
    MPDB_REQUIRE(e->finalized, "engine not finalized (call mpdb_engine_finalize after loading parameters)");
    MPDB_REQUIRE(p->n_hard_conds >= 0 && p->n_hard_conds <= MPDB_MAX_HARD_CONDS, "too many hard conditions");
    MPDB_REQUIRE(p->n_hard_conds == 0 || p->hard_cond_vals, "hard_cond_vals is null");
    MPDB_REQUIRE(p->n_steps_without_noise >= 0 && p->n_guide_steps >= 0, "negative step count");
    MPDB_REQUIRE(!g || guide_state_dim(g) == e->cfg.state_dim, "guide and model disagree on state_dim");
    MPDB_REQUIRE(!g || guide_device(g) == e->device, "guide and model live on different devices");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    if (ensure_workspace(e, B)) return 1;
    const int T = e->cfg.n_diffusion_steps, H = e->cfg.horizon, D = e->cfg.state_dim;
    const int n_iters = T + p->n_steps_without_noise;
    const long long n = (long long)B * H * D;

    MPDB_REQUIRE(p->horizon == H && p->state_dim == D,
                 "mpdb_sample_loop: tensors of horizon " + std::to_string(p->horizon) + " x state_dim " + std::to_string(p->state_dim) +
                     " given to an engine built for " + std::to_string(H) + " x " + std::to_string(D));
    MPDB_REQUIRE(!chain_out || (chain_batch_stride >= (int64_t)H * D && chain_step_stride >= (int64_t)H * D),
                 "mpdb_sample_loop: chain strides smaller than one trajectory");
    if (ensure_flags(e, n_iters * (p->n_guide_steps + 1) + 1 + n_iters * (p->n_guide_steps > 0 ? p->n_guide_steps : 1))) return 1;
    if (!p->use_cuda_graph || (g && guide_recording(g)))  // recording decisions assigns buffer slots per launch: no replay
        return enqueue_loop(e, g, p, noise, p->hard_cond_vals, x_out, chain_out, chain_step_stride, chain_batch_stride,
                            B, st);

    // ---- CUDA-graph path: the loop is captured once per configuration over engine-owned staging buffers ----
    const long long noise_floats = (long long)(n_iters + 1) * n;
    const long long hc_floats = (long long)p->n_hard_conds * B * D;
    const long long chain_floats = chain_out ? (long long)(n_iters + 1) * n : 0;
    const bool grow = noise_floats > e->g_noise_floats || hc_floats > e->g_hc_floats || chain_floats > e->g_chain_floats;
    if (grow) drop_graphs(e);
    if (ensure_staging(&e->g_noise, &e->g_noise_floats, noise_floats)) return 1;
    if (ensure_staging(&e->g_hc, &e->g_hc_floats, hc_floats > 0 ? hc_floats : 1)) return 1;
    if (ensure_staging(&e->g_chain, &e->g_chain_floats, chain_floats > 0 ? chain_floats : 1)) return 1;

    std::string key = std::to_string(B) + "|" + std::to_string((long long)(uintptr_t)g) + "|" +
                      std::to_string(p->n_steps_without_noise) + "|" + std::to_string(p->t_start_guide) + "|" +
                      std::to_string(p->n_guide_steps) + "|" + std::to_string(p->scale_grad_by_std) + "|" +
                      std::to_string(chain_out != nullptr) + "|" + std::to_string(p->n_hard_conds) + "|tc" +
                      std::to_string(e->tc_mode) + "/" + std::to_string(e->tc_amp_limit) + "/" + std::to_string(e->fuse_rtb) + "/" + std::to_string(e->use_mega) + "/" + std::to_string(e->fuse_final) + "/" + std::to_string(e->fuse_guide) + "/" + std::to_string(e->prec1_amp_limit);
    for (int k = 0; k < p->n_hard_conds; ++k) key += "," + std::to_string(p->hard_cond_rows[k]);
    for (int k = 0; k < n_iters; ++k) {
        float v = p->noise_std ? p->noise_std[k] : 1.0f;
        uint32_t bits;
        memcpy(&bits, &v, 4);
        key += ":" + std::to_string(bits);
    }
    if (g) {  // the guide configuration is baked into kernel arguments
        const mpdb_guide_config* gc = reinterpret_cast<const mpdb_guide_config*>(g);
        const unsigned char* bytes = reinterpret_cast<const unsigned char*>(gc);
        unsigned long long hsh = 1469598103934665603ull;
        for (size_t k = 0; k < sizeof(mpdb_guide_config); ++k) { hsh ^= bytes[k]; hsh *= 1099511628211ull; }
        key += "|" + std::to_string(hsh);
    }

    auto git = e->graphs.find(key);
    if (git == e->graphs.end()) {
        if (e->graphs.size() >= 16) {  // evict the least recently used entry
            auto victim = e->graphs.begin();
            for (auto it2 = e->graphs.begin(); it2 != e->graphs.end(); ++it2)
                if (it2->second.last_use < victim->second.last_use) victim = it2;
            if (victim->second.exec) cudaGraphExecDestroy(victim->second.exec);
            e->graphs.erase(victim);
        }
        // internal chain staging is always [S][B][H][D]
        cudaStream_t cs;
        MPDB_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        MPDB_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        const long long launches_before = mpdb::g_launch_count.load();
        int rc = enqueue_loop(e, g, p, e->g_noise, e->g_hc, e->xbuf[(n_iters) & 1] /* placeholder, fixed below */,
                              chain_out ? e->g_chain : nullptr, n, (long long)H * D, B, cs);
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        if (rc != 0 || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaStreamDestroy(cs);
            if (rc == 0) mpdb::set_error(std::string("graph capture failed: ") + cudaGetErrorString(ce));
            return 1;
        }
        mpdb_engine::LoopGraph lg;
        ce = cudaGraphInstantiate(&lg.exec, graph, 0);
        cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        if (ce != cudaSuccess) {
            mpdb::set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
            return 1;
        }
        lg.kernels = mpdb::g_launch_count.load() - launches_before;
        mpdb::g_launch_count.store(launches_before);  // capture enqueued nothing; replays are counted below
        git = e->graphs.emplace(key, lg).first;
    }
    git->second.last_use = ++e->graph_clock;
    MPDB_CHECK_CUDA(cudaMemcpyAsync(e->g_noise, noise, sizeof(float) * (size_t)noise_floats, cudaMemcpyDeviceToDevice, st));
    if (hc_floats > 0)
        MPDB_CHECK_CUDA(cudaMemcpyAsync(e->g_hc, p->hard_cond_vals, sizeof(float) * (size_t)hc_floats,
                                        cudaMemcpyDeviceToDevice, st));
    MPDB_CHECK_CUDA(cudaGraphLaunch(git->second.exec, st));
    mpdb::g_launch_count.fetch_add(git->second.kernels);
    // result: the captured loop wrote its last step into xbuf[n_iters & 1]
    MPDB_CHECK_CUDA(cudaMemcpyAsync(x_out, e->xbuf[(n_iters) & 1], sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    if (chain_out) {
        if (chain_step_stride == n && chain_batch_stride == (long long)H * D) {
            MPDB_CHECK_CUDA(cudaMemcpyAsync(chain_out, e->g_chain, sizeof(float) * (size_t)chain_floats,
                                            cudaMemcpyDeviceToDevice, st));
        } else {
            // strided destination: one 2D copy per step ([B] rows of H*D floats)
            for (int s = 0; s <= n_iters; ++s)
                MPDB_CHECK_CUDA(cudaMemcpy2DAsync(chain_out + (long long)s * chain_step_stride,
                                                  sizeof(float) * (size_t)chain_batch_stride, e->g_chain + (long long)s * n,
                                                  sizeof(float) * (size_t)H * D, sizeof(float) * (size_t)H * D, B,
                                                  cudaMemcpyDeviceToDevice, st));
        }
    }
    return 0;
}

// ddim_sample (diffusion_model_base.py:184-259), fused: per (time, time_next) pair one UNet forward whose last epilogue (or
// final_kernel) applies the DDIM update, then the guide evaluations when time_next < t_start_guide. No host synchronisation.
extern "C" int mpdb_ddim_loop(mpdb_engine* e, mpdb_guide* g, const mpdb_ddim_params* p, const float* x_init, float* x_out,
                              float* chain_out, int64_t chain_step_stride, int64_t chain_batch_stride, int32_t B, void* stream) {
    MPDB_REQUIRE(e && p && x_init && x_out && B > 0, "mpdb_ddim_loop: bad argument");
    MPDB_REQUIRE(e->finalized, "engine not finalized (call mpdb_engine_finalize after loading parameters)");
    MPDB_REQUIRE(p->n_steps >= 1 && p->times && p->times_next && p->sqrt_alpha_next && p->coef_noise, "mpdb_ddim_loop: missing step tables");
    MPDB_REQUIRE(p->n_hard_conds >= 0 && p->n_hard_conds <= MPDB_MAX_HARD_CONDS, "too many hard conditions");
    MPDB_REQUIRE(p->n_hard_conds == 0 || p->hard_cond_vals, "hard_cond_vals is null");
    MPDB_REQUIRE(p->n_guide_steps >= 0, "negative step count");
    MPDB_REQUIRE(!g || guide_state_dim(g) == e->cfg.state_dim, "guide and model disagree on state_dim");
    MPDB_REQUIRE(!g || guide_device(g) == e->device, "guide and model live on different devices");
    const int T = e->cfg.n_diffusion_steps, H = e->cfg.horizon, D = e->cfg.state_dim;
    MPDB_REQUIRE(p->horizon == H && p->state_dim == D,
                 "mpdb_ddim_loop: tensors of horizon " + std::to_string(p->horizon) + " x state_dim " + std::to_string(p->state_dim) +
                     " given to an engine built for " + std::to_string(H) + " x " + std::to_string(D));
    MPDB_REQUIRE(!chain_out || (chain_batch_stride >= (int64_t)H * D && chain_step_stride >= (int64_t)H * D),
                 "mpdb_ddim_loop: chain strides smaller than one trajectory");
    for (int k = 0; k < p->n_steps; ++k) {
        MPDB_REQUIRE(p->times[k] >= 0 && p->times[k] < T && p->times_next[k] < T, "mpdb_ddim_loop: time index out of range");
        MPDB_REQUIRE(p->times_next[k] >= 0 || k == p->n_steps - 1, "mpdb_ddim_loop: only the last pair may end at time_next < 0");
    }
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    if (ensure_workspace(e, B)) return 1;
    const int ng = g ? p->n_guide_steps : 0;
    if (ensure_flags(e, p->n_steps * (ng + 1) + 1)) return 1;
    if (ng > 0) MPDB_CHECK_CUDA(cudaMemsetAsync(e->flags, 0, sizeof(int) * (size_t)(p->n_steps * (ng + 1) + 1), st));
    const long long n = (long long)B * H * D;
    float* cur = e->xbuf[0];
    MPDB_CHECK_CUDA(cudaMemcpyAsync(cur, x_init, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    if (launch_copy_hc(cur, chain_out, chain_batch_stride, p->n_hard_conds, p->hard_cond_rows, p->hard_cond_vals, B, H, D, st)) return 1;
    for (int k = 0; k < p->n_steps; ++k) {
        const int t = p->times[k], t_next = p->times_next[k];
        const bool last = (k == p->n_steps - 1);
        float* nxt = last ? x_out : e->xbuf[(k + 1) & 1];
        float* chain_slot = chain_out ? chain_out + (long long)(k + 1) * chain_step_stride : nullptr;
        const bool guided = ng > 0 && t_next >= 0 && (long long)t_next < (long long)p->t_start_guide;
        // the DDIM update has no clamp and no posterior damping (eps reaches x scaled by |c - sqrt_recipm1[t] * sqrt(alpha_next)|,
        // ~3600 at the first pair): every DDIM forward uses the full 22-bit split (run_unet_body(..., ddim = true))
        const bool tc = e->tc_mode == 2 || (e->tc_mode == 1 && e->sched_host[1 * (size_t)T + t] <= e->tc_amp_limit);
        FinalArgs f;
        fill_final(e, f, cur, nullptr, t, B);
        f.n_hc = p->n_hard_conds;
        for (int q = 0; q < p->n_hard_conds; ++q) f.hc_rows[q] = p->hard_cond_rows[q];
        f.hc_vals = p->hard_cond_vals;
        f.ddim_san = p->sqrt_alpha_next[k];
        f.ddim_c = p->coef_noise[k];
        f.ddim_last = t_next < 0 ? 1 : 0;
        f.out = nxt;
        int* fl = e->flags + (long long)k * (ng + 1);
        if (guided) {
            f.mode = 4;
            f.flag_out = fl;
        } else {
            f.mode = 3;
            f.out2 = chain_slot;
            f.out2_bstride = chain_batch_stride;
        }
        bool fused = false;
        if (run_unet_body(e, cur, nullptr, t, B, st, tc, &f, &fused, /*ddim=*/true)) return 1;
        if (!fused && launch_final(f, st)) return 1;
        if (guided) {
            for (int it = 0; it < ng; ++it) {
                const bool klast = (it == ng - 1);
                GuideStepArgs a;
                memset(&a, 0, sizeof(a));
                a.x_in = nxt;
                a.x_out = nxt;
                a.flag_in = fl + it;
                a.flag_out = klast ? nullptr : fl + it + 1;
                a.n_hc = p->n_hard_conds;
                for (int q = 0; q < p->n_hard_conds; ++q) a.hc_rows[q] = p->hard_cond_rows[q];
                a.hc_vals = p->hard_cond_vals;
                if (klast) { a.out2 = chain_slot; a.out2_bstride = chain_batch_stride; }  // + sigma * noise with sigma = 0 (eta = 0)
                a.B = B;
                a.H = H;
                if (guide_launch_step(g, a, st)) return 1;
            }
        }
        cur = nxt;
    }
    return 0;
}

extern "C" int mpdb_engine_num_buffers(mpdb_engine* e) { return e ? (int)e->bufs.size() : 0; }

extern "C" int mpdb_engine_buffer_info(mpdb_engine* e, int idx, char* name, int name_cap, int32_t* channels,
                                       int32_t* length) {
    MPDB_REQUIRE(e && idx >= 0 && idx < (int)e->bufs.size(), "mpdb_engine_buffer_info: bad index");
    if (name && name_cap > 0) {
        strncpy(name, e->bufs[idx].name.c_str(), name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (channels) *channels = e->bufs[idx].C;
    if (length) *length = e->bufs[idx].L;
    return 0;
}

extern "C" int mpdb_engine_read_buffer(mpdb_engine* e, int idx, float* dev_out, int32_t B, void* stream) {
    MPDB_REQUIRE(e && dev_out && idx >= 0 && idx < (int)e->bufs.size(), "mpdb_engine_read_buffer: bad argument");
    MPDB_REQUIRE(B > 0 && B <= e->work_batch, "mpdb_engine_read_buffer: batch larger than the workspace");
    MPDB_REQUIRE(!e->alias_buffers, "mpdb_engine_read_buffer: set option alias_buffers = 0 before the forward pass");
    MPDB_ENTER_DEVICE(e->device);
    return launch_cm_to_bcl(buf_ptr(e, idx, e->work_batch), dev_out, B, e->bufs[idx].C, e->bufs[idx].L,
                            (cudaStream_t)stream);
}

// Per-layer device timing of one UNet forward (CUDA events on `stream`), for bench.py's roofline object.
// ms_out/flops_out/mode_out: arrays of at least mpdb_engine_num_ops(e) entries (ops + the fused final kernel,
// which is reported with mode 4).
extern "C" int mpdb_engine_num_ops(mpdb_engine* e) { return e ? (int)e->ops.size() + 1 : 0; }

extern "C" int mpdb_profile_forward(mpdb_engine* e, const float* x, int32_t t, int32_t B, int32_t reps, float* ms_out,
                                    double* flops_out, int32_t* mode_out, void* stream) {
    MPDB_REQUIRE(e && x && ms_out && flops_out && mode_out && B > 0 && reps > 0, "mpdb_profile_forward: bad argument");
    MPDB_REQUIRE(e->finalized, "engine not finalized");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(e->device);
    if (ensure_workspace(e, B)) return 1;
    cudaEvent_t ev0, ev1;
    MPDB_CHECK_CUDA(cudaEventCreate(&ev0));
    MPDB_CHECK_CUDA(cudaEventCreate(&ev1));
    const bool tc = e->tc_mode != 0;
    if (run_unet_body(e, x, nullptr, t, B, st, tc)) return 1;  // warm-up, fills every buffer
    int k = 0;
    for (size_t i = 0; i < e->ops.size(); ++i) {
        const ConvOp& op = e->ops[i];
        const bool fused = can_fuse_rtb(e, i, tc, B);
        MPDB_CHECK_CUDA(cudaEventRecord(ev0, st));
        for (int r = 0; r < reps; ++r) {
            if (fused) { if (launch_rtb(e, i, nullptr, t, B, st)) return 1; }
            else if (launch_op(e, op, x, nullptr, t, B, st, tc)) return 1;
        }
        MPDB_CHECK_CUDA(cudaEventRecord(ev1, st));
        MPDB_CHECK_CUDA(cudaEventSynchronize(ev1));
        float ms = 0.f;
        MPDB_CHECK_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        ms_out[k] = ms / reps;
        flops_out[k] = op_flops(e, op, B);
        mode_out[k] = op.mode == MODE_INPUT ? 6 : (tc && op.tc_ok) ? 5 : op.mode;  // 5 = tcgen05 conv, 6 = layout conversion
        ++k;
        if (fused) {  // the pair ran as one launch: its time and FLOPs are reported on the first entry (mode 8)
            flops_out[k - 1] += op_flops(e, e->ops[i + 1], B);
            mode_out[k - 1] = 8;
            ms_out[k] = 0.f; flops_out[k] = 0.0; mode_out[k] = 9;
            ++k;
            ++i;
        }
    }
    {
        FinalArgs f;
        fill_final(e, f, x, nullptr, t, B);
        f.mode = 1;
        f.out = e->xbuf[1];
        MPDB_CHECK_CUDA(cudaEventRecord(ev0, st));
        for (int r = 0; r < reps; ++r)
            if (launch_final(f, st)) return 1;
        MPDB_CHECK_CUDA(cudaEventRecord(ev1, st));
        MPDB_CHECK_CUDA(cudaEventSynchronize(ev1));
        float ms = 0.f;
        MPDB_CHECK_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        ms_out[k] = ms / reps;
        flops_out[k] = 2.0 * B * e->cfg.horizon * e->cfg.state_dim * (double)e->cfg.unet_input_dim;
        mode_out[k] = 4;
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return 0;
}

// Unit-test hook for the tcgen05 implicit-GEMM core (no epilogue): raw accumulators of a k=5 convolution.
//   x_cm : device fp32 [B][CI][L+4] (zero halo),  w : device fp32 [CO][CI][5]
//   raw  : device fp32 [ceil(B/SPT)][CO/32][128][32], SPT = 132 / (L+4); row r of a tile = (b % SPT)*(L+4) + l
extern "C" int mpdb_debug_tc_conv5(const float* x_cm, const float* w, float* raw, int32_t B, int32_t CI, int32_t CO,
                                   int32_t L, void* stream) {
    MPDB_REQUIRE(x_cm && w && raw && B > 0, "mpdb_debug_tc_conv5: bad argument");
    MPDB_REQUIRE(CI % TC_KCH == 0 && CO % TC_NT == 0 && L % 4 == 0 && L + 4 <= TC_RT, "mpdb_debug_tc_conv5: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int SPT = TC_RT / (L + 4);
    const int tiles = (B + SPT - 1) / SPT;
    const size_t plane = (size_t)tiles * (CI / 8) * TC_RT * 8;
    float* wp = nullptr;
    unsigned short *wt = nullptr, *xh = nullptr, *xl = nullptr;
    MPDB_CHECK_CUDA(cudaMalloc(&wp, sizeof(float) * (size_t)CI * CO * 5));
    MPDB_CHECK_CUDA(cudaMalloc(&wt, 2 * 2 * (size_t)CI * CO * 5));
    MPDB_CHECK_CUDA(cudaMalloc(&xh, 2 * 2 * plane));  // hi plane followed by the lo plane (one tensor map covers both)
    xl = xh + plane;
    MPDB_CHECK_CUDA(cudaMemsetAsync(xh, 0, 2 * 2 * plane, st));
    int rc = launch_repack_conv(w, wp, CO, CI, 5, 0, st);
    if (!rc) rc = launch_pack_tc_weights(wp, wt, nullptr, CI, CI, CO, 5, 0x43210u, st);
    if (!rc) rc = launch_cm_to_tc(x_cm, xh, xl, B, CI, L, st);
    if (!rc) {
        TcConvArgs a;
        memset(&a, 0, sizeof(a));
        a.in0_hi = xh; a.in0_lo = xl; a.c0 = CI;
        a.w = wt;
        a.raw_out = raw;
        a.CO = CO; a.L = L; a.B = B; a.gs = 32; a.mode = TCM_CONV5;
        a.prec = 3;
        a.tm_nch[0] = 1;
        rc = make_act_tensor_map(&a.tm[0], xh, (long long)plane, CI, tiles, a.tm_nch[0], 2);
        float* dummy = wp;  // gamma/beta/bias are not read in raw mode but must be non-null for the launch checks
        a.gamma = dummy; a.beta = dummy; a.bias = dummy;
        if (!rc) rc = launch_conv5_tc(a, st);
    }
    cudaError_t e1 = cudaStreamSynchronize(st);
    cudaFree(wp); cudaFree(wt); cudaFree(xh);
    if (rc) return rc;
    MPDB_CHECK_CUDA(e1);
    return 0;
}

// Debug: per-layer clock64 stamps of the whole-forward kernel (option "mega_timeline"): [layers][8 ranks][4], plus the
// layer descriptors' (type, L, CO, MT*NC) as 4 ints per layer in `desc_out`.
extern "C" int mpdb_engine_read_mega_timeline(mpdb_engine* e, int64_t* host_out, int32_t* desc_out, int32_t max_layers) {
    MPDB_REQUIRE(e && host_out && e->mega_dbg, "mpdb_engine_read_mega_timeline: timeline not enabled");
    MPDB_ENTER_DEVICE(e->device);
    MPDB_CHECK_CUDA(cudaDeviceSynchronize());
    const int n = e->mega.n_layers < max_layers ? e->mega.n_layers : max_layers;
    // the slots behind the layers (44..47) hold per-cluster {entry, setup done, exit} stamps on the GPU-wide ns timer
    const int n_copy = max_layers < MEGA_MAX_LAYERS ? max_layers : MEGA_MAX_LAYERS;
    MPDB_CHECK_CUDA(cudaMemcpy(host_out, e->mega_dbg, sizeof(long long) * MEGA_DBG * MEGA_CLUSTER * n_copy, cudaMemcpyDeviceToHost));
    if (desc_out)
        for (int k = 0; k < n; ++k) {
            const MegaLayer& L = e->mega.layers[k];
            desc_out[4 * k + 0] = L.type; desc_out[4 * k + 1] = L.L; desc_out[4 * k + 2] = L.CO; desc_out[4 * k + 3] = L.MT * L.NC;
        }
    return n;
}

// Debug: clock64 stamps (16 per op) of CTA (0,0) of every tensor-core conv of the last forward (option "timeline").
extern "C" int mpdb_engine_read_timeline(mpdb_engine* e, int64_t* host_out, int32_t max_ops) {
    MPDB_REQUIRE(e && host_out && e->dbg_buf, "mpdb_engine_read_timeline: timeline not enabled");
    MPDB_ENTER_DEVICE(e->device);
    MPDB_CHECK_CUDA(cudaDeviceSynchronize());
    int n = (int)e->ops.size() < max_ops ? (int)e->ops.size() : max_ops;
    MPDB_CHECK_CUDA(cudaMemcpy(host_out, e->dbg_buf, sizeof(long long) * 16 * n, cudaMemcpyDeviceToHost));
    return 0;
}
