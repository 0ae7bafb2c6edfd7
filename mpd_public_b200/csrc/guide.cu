// Cost-gradient guide: GuideManagerTrajectoriesWithVelocity.forward (reference guides.py:173-236) with the
// restated CostComposite([CostCollision(field)...], CostGPTrajectory) of SURVEY.md Appendix C, and the
// update x <- x + guide(x) of guide_gradient_steps (sample_functions.py:65-83), as ONE kernel for the
// n_guide_steps evaluations of a reverse step:
//
//   unnormalise (LimitsNormalizer.unnormalize incl. the batch-global clip flag)  ->  linear interpolation
//   H -> n_interp  ->  FK of the collision spheres  ->  per collision cost {nearest-texel SDF lookup | workspace box |
//   pairwise self-collision distances}  ->  hinge  ->  hand-derived adjoint (J^T through the chain, two-tap scatter of
//   the interpolation written as a gather)  ->  per-cost 1/sigma^2, clip-by-norm, endpoint zeroing, weighting  ->
//   GP-prior 3-tap stencil gradient  ->  x + grad, hard conditioning [-> + std * noise * noise_std, hard conditioning].
//
// One CTA (512 threads) per trajectory; the trajectory, its unnormalised copy and every per-cost gradient stay in
// shared memory for all evaluations of the launch. The only global traffic is x in/out (+noise) and the SDF texel gathers.
//
// Thread mapping (round 2): an interpolated row belongs to a QUAD of lanes of one warp. Lane k of the quad takes the
// sines / cosines of joints k and k + 4 (exchanged with shuffles), lanes 0-2 run one matrix row of the kinematic chain
// each, the frames go through a per-row shared-memory scratch that only this quad touches (__syncwarp, no block barrier),
// then lane k handles spheres k, k + 4, ... for every collision cost and the quad adds its four partial J^T sums with
// xor-shuffles in a fixed order. Interpolation, FK, lookups and J^T therefore run between two block barriers; one
// evaluation needs four (unnormalise | rows | interpolation adjoint + clip | GP + update).
//
// The batch-global clip flag of LimitsNormalizer.unnormalize (normalization.py:160) couples the trajectories of a batch
// between two evaluations. A CTA resolves it locally whenever it can: with an element beyond 1 + 1e-4 the flag is set
// whatever the others do; with no element beyond 1 the clamp is the identity on this trajectory whatever the flag. Only a
// trajectory with an element in (1, 1 + 1e-4] and none beyond needs the others: it waits on a grid-wide counter that every
// CTA bumps without waiting (all CTAs co-resident, checked by the launcher). Same values as the reference in every case.
#include <math.h>

#include <vector>

#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

struct mpdb_guide {
    mpdb_guide_config cfg;
    int device;
    int* flags;  // device scratch [64]: [0,2) flags of the standalone entry points, [8] batch-dependent-clamp counter,
                 // [16,37) per-evaluation flags and [40,60) grid counters of a fused mpdb_guide_steps launch
    // parity instrumentation (mpdb_guide_record_decisions)
    int32_t* dec_buf = nullptr;
    long long dec_capacity = 0, dec_count = 0;
    int dec_batch = 0;
    int coresident_H = -1, coresident = 0;  // cached guide_max_coresident(H)
};

namespace mpdb {

constexpr int GUIDE_THREADS = 512;
constexpr int FK_ROWS = 128;   // interpolated rows per pass (one quad of lanes each)
constexpr int MAXC = MPDB_MAX_GRID_FIELDS + 2;  // collision costs: grid fields, workspace border, self-collision

struct GuideDev {
    int robot_kind, q_dim, ws_dim, n_spheres, D;
    int sphere_frame[MPDB_MAX_SPHERES];
    float sphere_off[MPDB_MAX_SPHERES][3];
    float sphere_r[MPDB_MAX_SPHERES];
    int frame_begin[10];                   // spheres attached to frame f (1..8) are frame_sphere[frame_begin[f-1] .. frame_begin[f])
    int frame_sphere[MPDB_MAX_SPHERES];    // sphere indices sorted by frame (stable)
    float mins[MPDB_MAX_STATE_DIM], range[MPDB_MAX_STATE_DIM];
    int n_grid, has_border, has_self, n_coll;
    const float* tex[MPDB_MAX_GRID_FIELDS];
    int gshape[MPDB_MAX_GRID_FIELDS][3];
    float glo[MPDB_MAX_GRID_FIELDS][3];
    float cell[MPDB_MAX_GRID_FIELDS];
    float blo[3], bhi[3];
    unsigned self_pairs[MPDB_MAX_SPHERES];
    // per collision cost, kernel order: grid fields, border, self
    float margin[MAXC], isig2[MAXC], weight[MAXC];
    float dt, w_gp;
    int use_gp, clip;
    float max_norm;
    int n_interp, vel_fd;
    int simple_spheres;  // sphere i rides on frame i + 1 with zero offset (the default Panda model): centres = frame origins
    float gp_a, gp_b, gp_c;
    float joint_xyz[7][3];
    float joint_cr[7], joint_sr[7];
    float flange[3];
};

// Panda chain constants: public Franka URDF values
static const double kPandaXYZ[7][3] = {{0.0, 0.0, 0.333}, {0.0, 0.0, 0.0},  {0.0, -0.316, 0.0}, {0.0825, 0.0, 0.0},
                                       {-0.0825, 0.384, 0.0}, {0.0, 0.0, 0.0}, {0.088, 0.0, 0.0}};
static const double kPandaRoll[7] = {0.0, -M_PI / 2, M_PI / 2, M_PI / 2, -M_PI / 2, M_PI / 2, M_PI / 2};
static const double kPandaFlange[3] = {0.0, 0.0, 0.107};

static GuideDev make_dev(const mpdb_guide_config& c) {
    GuideDev d;
    memset(&d, 0, sizeof(d));
    d.robot_kind = c.robot_kind;
    d.q_dim = c.q_dim;
    d.ws_dim = c.ws_dim;
    d.n_spheres = c.n_spheres;
    d.D = 2 * c.q_dim;
    for (int i = 0; i < c.n_spheres; ++i) {
        d.sphere_frame[i] = c.sphere_frame[i];
        for (int k = 0; k < 3; ++k) d.sphere_off[i][k] = c.sphere_offset[i][k];
        d.sphere_r[i] = c.sphere_radius[i];
        d.self_pairs[i] = c.self_pairs[i];
    }
    {
        int n = 0;
        for (int f = 1; f <= 8; ++f) {
            d.frame_begin[f - 1] = n;
            for (int i = 0; i < c.n_spheres; ++i)
                if (c.sphere_frame[i] == f) d.frame_sphere[n++] = i;
        }
        d.frame_begin[8] = d.frame_begin[9] = n;
    }
    d.simple_spheres = c.robot_kind == 1 && c.n_spheres <= 8;
    for (int i = 0; i < c.n_spheres; ++i)
        if (c.sphere_frame[i] != i + 1 || c.sphere_offset[i][0] != 0.f || c.sphere_offset[i][1] != 0.f || c.sphere_offset[i][2] != 0.f)
            d.simple_spheres = 0;
    for (int i = 0; i < d.D; ++i) {
        d.mins[i] = c.mins[i];
        d.range[i] = c.maxs[i] - c.mins[i];  // fp32 subtraction, as `self.maxs - self.mins`
    }
    d.n_grid = c.n_grid_fields;
    d.has_border = c.has_border ? 1 : 0;
    d.has_self = (c.has_self && c.robot_kind == 1) ? 1 : 0;
    d.n_coll = d.n_grid + d.has_border + d.has_self;
    auto isig2 = [](float s) { return (float)(1.0 / ((double)s * (double)s)); };
    for (int i = 0; i < c.n_grid_fields; ++i) {
        d.tex[i] = c.grid_texels[i];
        for (int k = 0; k < 3; ++k) { d.gshape[i][k] = c.grid_shape[i][k]; d.glo[i][k] = c.grid_lo[i][k]; }
        d.cell[i] = c.grid_cell[i];
        d.margin[i] = c.margin_grid[i];
        d.isig2[i] = isig2(c.sigma_grid[i]);
        d.weight[i] = c.weight_grid[i];
    }
    int k = d.n_grid;
    if (d.has_border) { d.margin[k] = c.margin_border; d.isig2[k] = isig2(c.sigma_border); d.weight[k] = c.weight_border; ++k; }
    if (d.has_self) { d.margin[k] = c.margin_self; d.isig2[k] = isig2(c.sigma_self); d.weight[k] = c.weight_self; ++k; }
    for (int q3 = 0; q3 < 3; ++q3) { d.blo[q3] = c.border_lo[q3]; d.bhi[q3] = c.border_hi[q3]; }
    d.dt = c.dt;
    d.w_gp = c.weight_gp;
    d.use_gp = c.use_gp;
    d.clip = c.clip_grad;
    d.max_norm = c.max_grad_norm;
    d.n_interp = c.n_interp;
    d.vel_fd = c.vel_from_fd ? 1 : 0;
    const double dt = (double)c.dt, s = 1.0 / ((double)c.sigma_gp * (double)c.sigma_gp);
    d.gp_a = (float)(12.0 / (dt * dt * dt) * s);
    d.gp_b = (float)(-6.0 / (dt * dt) * s);
    d.gp_c = (float)(4.0 / dt * s);
    for (int i = 0; i < 7; ++i) {
        for (int q3 = 0; q3 < 3; ++q3) d.joint_xyz[i][q3] = (float)kPandaXYZ[i][q3];
        d.joint_cr[i] = (float)cos(kPandaRoll[i]);
        d.joint_sr[i] = (float)sin(kPandaRoll[i]);
    }
    for (int q3 = 0; q3 < 3; ++q3) d.flange[q3] = (float)kPandaFlange[q3];
    return d;
}

// Forward kinematics of one interpolated row: joint origins (3 x 7), joint axes (3 x 7) and collision-sphere centres
// (3 x n_spheres) into per-row scratch with stride FK_ROWS. Panda chain: T_i = T_{i-1} Trans(xyz_i) Rx(roll_i) Rz(q_i)
// (SURVEY Appendix E); point mass: centre = q.
// One matrix row of the chain: row r3 of R and component r3 of the origin depend only on row r3 of the previous frame,
// so three lanes per interpolated row run the chain independently (3x shorter dependency chain).
__device__ __forceinline__ void fk_chain_row(const GuideDev& g, int r3, const float (&sq)[7], const float (&cq)[7], float* sc, float* cen) {
    float c0 = r3 == 0 ? 1.f : 0.f, c1 = r3 == 1 ? 1.f : 0.f, c2 = r3 == 2 ? 1.f : 0.f;  // row r3 of R
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        o += c0 * g.joint_xyz[j][0] + c1 * g.joint_xyz[j][1] + c2 * g.joint_xyz[j][2];
        const float cr = g.joint_cr[j], sr = g.joint_sr[j];
        const float a1 = c1 * cr + c2 * sr;   // (R Rx) column 1
        const float a2 = -c1 * sr + c2 * cr;  // (R Rx) column 2
        const float n0 = cq[j] * c0 + sq[j] * a1;
        const float n1 = -sq[j] * c0 + cq[j] * a1;
        c0 = n0; c1 = n1; c2 = a2;
        sc[(j * 3 + r3) * FK_ROWS] = o;
        sc[(21 + j * 3 + r3) * FK_ROWS] = c2;  // joint axis = third column
        if (g.simple_spheres) {
            if (j < g.n_spheres) cen[(j * 3 + r3) * FK_ROWS] = o;  // sphere j sits on this frame's origin
        } else {
            // the spheres attached to this frame (host-sorted list: no scan over all spheres per joint)
            for (int t = g.frame_begin[j]; t < g.frame_begin[j + 1]; ++t) {
                const int s = g.frame_sphere[t];
                cen[(s * 3 + r3) * FK_ROWS] = o + c0 * g.sphere_off[s][0] + c1 * g.sphere_off[s][1] + c2 * g.sphere_off[s][2];
            }
        }
    }
    o += c0 * g.flange[0] + c1 * g.flange[1] + c2 * g.flange[2];
    if (g.simple_spheres) {
        if (g.n_spheres > 7) cen[(7 * 3 + r3) * FK_ROWS] = o;
        return;
    }
    for (int t = g.frame_begin[7]; t < g.frame_begin[8]; ++t) {
        const int s = g.frame_sphere[t];
        cen[(s * 3 + r3) * FK_ROWS] = o + c0 * g.sphere_off[s][0] + c1 * g.sphere_off[s][1] + c2 * g.sphere_off[s][2];
    }
}

// whole chain by one thread (post-sampling evaluation, FK unit test)
__device__ __forceinline__ void fk_row(const GuideDev& g, const float (&qv)[7], float* sc, float* cen) {
    if (g.robot_kind == 1) {
        float sq[7], cq[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) sincosf(qv[j], &sq[j], &cq[j]);
        for (int r3 = 0; r3 < 3; ++r3) fk_chain_row(g, r3, sq, cq, sc, cen);
    } else {
        for (int s = 0; s < g.n_spheres; ++s)
            for (int r3 = 0; r3 < 3; ++r3) cen[(s * 3 + r3) * FK_ROWS] = r3 < g.ws_dim ? qv[r3] : 0.f;
    }
}

// Nearest-texel index of point p on grid field f (Appendix C.5): round((p - lo) / cell), clamped
__device__ __forceinline__ long long grid_flat_index(const GuideDev& g, int f, const float (&p)[3]) {
    long long flat = 0;
    for (int d = 0; d < g.ws_dim; ++d) {
        float uu = rintf(__fdiv_rn(__fsub_rn(p[d], g.glo[f][d]), g.cell[f]));
        uu = fminf(fmaxf(uu, 0.f), (float)(g.gshape[f][d] - 1));
        flat = flat * g.gshape[f][d] + (long long)uu;
    }
    return flat;
}
__device__ __forceinline__ void grid_fetch(const GuideDev& g, int f, long long flat, float& sdf, float (&gr)[3]) {
    if (g.ws_dim == 3) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(g.tex[f]) + flat);
        sdf = t.x; gr[0] = t.y; gr[1] = t.z; gr[2] = t.w;
    } else {
        const float* t = g.tex[f] + flat * 3;
        sdf = __ldg(t); gr[0] = __ldg(t + 1); gr[1] = __ldg(t + 2); gr[2] = 0.f;
    }
}
// workspace-boundary box (C.4): distance to the nearest wall, positive inside; code = (axis << 1) | (low wall)
__device__ __forceinline__ void border_lookup(const GuideDev& g, const float (&p)[3], float& sdf, float (&gr)[3], int& code) {
    gr[0] = gr[1] = gr[2] = 0.f;
    int arg = 0; float sgn = 1.f, best = 3.4e38f;
    for (int d = 0; d < g.ws_dim; ++d) {
        const float lo = __fsub_rn(p[d], g.blo[d]), hi = __fsub_rn(g.bhi[d], p[d]);
        const float m = fminf(lo, hi);
        if (m < best) { best = m; arg = d; sgn = (lo <= hi) ? 1.f : -1.f; }
    }
    sdf = best;
    gr[arg] = sgn;
    code = (arg << 1) | (sgn > 0.f ? 1 : 0);
}

__device__ __forceinline__ float clip_scale(float n, float max_norm) {
    // torch.clip(n, 0, max) / n
    return fminf(fmaxf(n, 0.f), max_norm) / n;
}

// KIND: 1 = Panda, 0 = point mass (compile-time copy of g.robot_kind): the other robot's code is not instantiated (the kernel
// runs every instruction once per evaluation, so its size is its cost).
// OCC: CTAs per SM the register allocation aims at: 1 = ~92 registers, nothing spilled (a batch that fits one CTA per SM is
// latency-bound per trajectory), 2 = 64 registers with ~150 B of spills, two trajectories per SM in flight (large batches).
template <int KIND, int OCC>
__global__ void __launch_bounds__(GUIDE_THREADS, OCC) guide_step_kernel(const __grid_constant__ GuideDev g, const __grid_constant__ GuideStepArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, D = g.D, q = g.q_dim, NI = g.n_interp, NTH = GUIDE_THREADS;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n_coll = g.n_coll, S = g.n_spheres;

    float* xn = smem;                      // [H][D] normalised trajectory (updated in place between evaluations)
    float* xu = xn + H * D;                // [H][D] unnormalised
    float* gq = xu + H * D;                // [n_coll][NI][8] d cost_c / d q_interp
    float* tapw = gq + n_coll * NI * 8;    // [H][8] taps of the interpolation adjoint: weight ...
    int* tapi = reinterpret_cast<int*>(tapw + H * 8);   // [H][8] ... and interpolated row (built once per launch)
    float* gvs = tapw + 2 * H * 8;         // [H][8] scratch plane (finite-difference velocity mode)
    float* w1s = gvs + H * 8;                           // [NI] interpolation weight of the upper tap
    int* i0s = reinterpret_cast<int*>(w1s + NI);        // [NI] lower tap
    float* fk = reinterpret_cast<float*>(i0s + NI);     // [42 + 3 S][FK_ROWS] FK scratch of one pass
    __shared__ int s_flag;

    long long* dbg = (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) ? a.dbg : nullptr;
    if (dbg) dbg[0] = clock64();
    const int n_it = a.n_iters > 1 ? a.n_iters : 1;
    const bool pos_only = a.vel_io != nullptr || g.vel_fd;  // x holds positions only (GuideManagerTrajectories)
    const int Dio = pos_only ? q : D;                       // columns of x_in / x_out
    int flag = 0;  // read after the dependency wait below
    const float ratio = NI > 1 ? (float)(H - 1) / (float)(NI - 1) : 0.f;  // align_corners=True
    const float inv_ratio = ratio > 0.f ? 1.f / ratio : 0.f;
    // Adjoint of the linear interpolation as a gather: support row h receives from the interpolated rows i whose lower / upper
    // tap is h. Which rows, and with which weight, depends only on (H, n_interp): the (row, weight) pairs are tabulated once per
    // launch, in ascending i (the summation order of the on-the-fly form, so the sums are bit-identical), instead of being
    // re-derived by every thread in every evaluation (22 % of the kernel's instructions, profiles/r02e_ncu_guide_hotspots.txt).
    // Up to 8 taps per row (5 at n_interp = 2 H); a denser interpolation falls back to the on-the-fly loop.
    __shared__ int s_tap_overflow;
    if (tid == 0) s_tap_overflow = 0;
    __syncthreads();
    for (int h = tid; h < H; h += NTH) {
        int lo_i = (int)floorf((float)(h - 1) * inv_ratio), hi_i = (int)ceilf((float)(h + 1) * inv_ratio);
        if (lo_i < 0) lo_i = 0;
        if (hi_i > NI - 1) hi_i = NI - 1;
        int cnt = 0;
        for (int i = lo_i; i <= hi_i; ++i) {
            const float r = ratio * (float)i;  // the same expressions as in the row pass below
            int i0 = (int)r;
            if (i0 > H - 1) i0 = H - 1;
            const float l1 = fminf(fmaxf(r - (float)i0, 0.f), 1.f);
            const int i1 = i0 + (i0 < H - 1 ? 1 : 0);
            const float cw = (i0 == h ? 1.f - l1 : 0.f) + (i1 == h ? l1 : 0.f);
            if (cw != 0.f) {
                if (cnt < 8) { tapw[h * 8 + cnt] = cw; tapi[h * 8 + cnt] = i; }
                ++cnt;
            }
        }
        if (cnt > 8) s_tap_overflow = 1;
        for (int c = cnt; c < 8; ++c) { tapw[h * 8 + c] = 0.f; tapi[h * 8 + c] = 0; }
    }
    __syncthreads();
    const bool use_taps = s_tap_overflow == 0;
    pdl_wait();  // x and the clip flag come from the previous kernel; everything above depends on the launch parameters only
    flag = a.flag_in ? *a.flag_in : 0;

    for (int it = 0; it < n_it; ++it) {
        const bool last_it = it == n_it - 1;
        // ---------------- (A) load / unnormalise ----------------
        {
            const float* xin = a.x_in + (long long)b * H * Dio;
            for (int i = tid; i < H * D; i += NTH) {
                const int d = i % D, hrow = i / D;
                if (pos_only && d >= q) {  // velocity half of the state: already unnormalised, not part of x
                    xn[i] = 0.f;
                    if (!g.vel_fd) xu[i] = a.vel_io[((long long)b * H + hrow) * q + (d - q)];
                    continue;
                }
                const float v = it == 0 ? xin[hrow * Dio + d] : xn[i];  // later evaluations continue from shared memory
                xn[i] = v;
                const float vc = flag ? fminf(fmaxf(v, -1.f), 1.f) : v;
                // ((x + 1) / 2) * (maxs - mins) + mins, reference operation order (normalization.py:165-167)
                xu[i] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(vc, 1.f), 0.5f), g.range[d]), g.mins[d]);
            }
            __syncthreads();
            if (flag && a.dep_count != nullptr) {
                // bookkeeping for the shard-equivalence tests: the batch's clip flag is set; was THIS trajectory's clamp decided
                // by the others (elements in (1, 1 + 1e-4] and none beyond)? Then a different batch composition can change it.
                bool hi = false, mid = false;
                for (int i = tid; i < H * D; i += NTH) {
                    if (pos_only && (i % D) >= q) continue;
                    const float v = xn[i];
                    hi |= (v > 1.0001f) || (v < -1.0001f);
                    mid |= (v > 1.f) || (v < -1.f);
                }
                const int any_hi = __syncthreads_or(hi ? 1 : 0), any_mid = __syncthreads_or(mid ? 1 : 0);
                if (!any_hi && any_mid && tid == 0) atomicAdd(a.dep_count, 1u);
            }
            if (g.vel_fd) {
                // robot.get_velocity of a position-only trajectory (guides.py:77-78; torch_robotics source absent, restated as
                // the central difference with zero end rows, oracle switch FD_CENTRAL)
                for (int i = tid; i < H * q; i += NTH) {
                    const int k = i % q, hrow = i / q;
                    float v = 0.f;
                    if (hrow > 0 && hrow < H - 1) v = __fdiv_rn(__fsub_rn(xu[(hrow + 1) * D + k], xu[(hrow - 1) * D + k]), __fmul_rn(2.f, g.dt));
                    xu[hrow * D + q + k] = v;
                }
                __syncthreads();
            }
        }
        if (dbg) dbg[1] = clock64();  // trajectory loaded + unnormalised

        // ---------------- (B-D) collision costs on the interpolated trajectory, one quad of lanes per row ----------------
        if (n_coll > 0) {
            int32_t* dec = a.dec ? a.dec + ((long long)it * a.B + b) * n_coll * NI * S : nullptr;
            for (int ibase = 0; ibase < NI; ibase += FK_ROWS) {
                const int il = tid >> 2, ql = tid & 3;
                const int i_raw = ibase + il;
                const bool row_on = i_raw < NI;
                const int i = row_on ? i_raw : NI - 1;  // rows past the end recompute the last row (warp-convergent), write nothing
                const float r = ratio * (float)i;
                int i0 = (int)r;
                if (i0 > H - 1) i0 = H - 1;
                const float l1 = fminf(fmaxf(r - (float)i0, 0.f), 1.f);
                const float l0 = 1.f - l1;
                const int i1 = i0 + (i0 < H - 1 ? 1 : 0);
                if (ql == 3 && row_on) { i0s[i] = i0; w1s[i] = l1; }
                float* sc = fk + il;
                float* cen = sc + 42 * FK_ROWS;
                const unsigned qbase = (unsigned)(lane & ~3);
                if (KIND == 1) {
                    // (B) lane k: sine / cosine of the interpolated joints k and k + 4, handed round the quad with shuffles
                    const int ka = ql, kb = ql + 4;
                    const float qa = __fadd_rn(__fmul_rn(l0, xu[i0 * D + ka]), __fmul_rn(l1, xu[i1 * D + ka]));
                    float s_a, c_a, s_b = 0.f, c_b = 1.f;
                    sincosf(qa, &s_a, &c_a);
                    if (kb < 7) {
                        const float qb = __fadd_rn(__fmul_rn(l0, xu[i0 * D + kb]), __fmul_rn(l1, xu[i1 * D + kb]));
                        sincosf(qb, &s_b, &c_b);
                    }
                    float sq[7], cq[7];
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        sq[k] = __shfl_sync(0xffffffffu, k < 4 ? s_a : s_b, qbase | (unsigned)(k & 3));
                        cq[k] = __shfl_sync(0xffffffffu, k < 4 ? c_a : c_b, qbase | (unsigned)(k & 3));
                    }
                    // (C) lanes 0-2: one matrix row of the kinematic chain each
                    if (ql < 3) fk_chain_row(g, ql, sq, cq, sc, cen);
                } else {
                    // point mass: every sphere sits at the interpolated position; lane k writes coordinate k
                    if (ql < 3) {
                        const float v = ql < g.ws_dim ? __fadd_rn(__fmul_rn(l0, xu[i0 * D + ql]), __fmul_rn(l1, xu[i1 * D + ql])) : 0.f;
                        for (int s = 0; s < S; ++s) cen[(s * 3 + ql) * FK_ROWS] = v;
                    }
                }
                __syncwarp();
                // (D) lane k: spheres k, k + 4, ... of every collision cost; quad-sum of the partial J^T products
                for (int f = 0; f < n_coll; ++f) {
                    const int kind = f < g.n_grid ? 0 : (f == g.n_grid && g.has_border) ? 1 : 2;
                    float dq[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    // grid field: all texel gathers of this lane are issued before any is consumed
                    long long flat[4] = {0, 0, 0, 0};
                    float tsdf[4], tg[4][3];
                    if (kind == 0) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int s = ql + 4 * u;
                            tsdf[u] = 3.4e38f; tg[u][0] = tg[u][1] = tg[u][2] = 0.f;
                            if (s < S) {
                                const float p[3] = {cen[(s * 3 + 0) * FK_ROWS], cen[(s * 3 + 1) * FK_ROWS], cen[(s * 3 + 2) * FK_ROWS]};
                                flat[u] = grid_flat_index(g, f, p);
                                grid_fetch(g, f, flat[u], tsdf[u], tg[u]);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int s = ql + 4 * u;
                        if (s >= S) continue;
                        const float p0 = cen[(s * 3 + 0) * FK_ROWS], p1 = cen[(s * 3 + 1) * FK_ROWS], p2 = cen[(s * 3 + 2) * FK_ROWS];
                        float gx = 0.f, gy = 0.f, gz = 0.f;  // d cost / d p of this sphere
                        int code = 0;
                        if (kind == 0) {
                            const float viol = __fsub_rn(__fadd_rn(g.sphere_r[s], g.margin[f]), tsdf[u]);
                            const bool act = viol > 0.f;
                            if (act) { gx = -tg[u][0]; gy = -tg[u][1]; gz = -tg[u][2]; }
                            code = (int)(flat[u] << 1) | (act ? 1 : 0);
                        } else if (kind == 1) {
                            const float p[3] = {p0, p1, p2};
                            float sdf, gr[3];
                            int bc;
                            border_lookup(g, p, sdf, gr, bc);
                            const float viol = __fsub_rn(__fadd_rn(g.sphere_r[s], g.margin[f]), sdf);
                            const bool act = viol > 0.f;
                            if (act) { gx = -gr[0]; gy = -gr[1]; gz = -gr[2]; }
                            code = (bc << 1) | (act ? 1 : 0);
                        } else {
                            // self-collision: every listed partner t with |c_s - c_t| - r_s - r_t < margin pushes s away
                            for (unsigned mask = g.self_pairs[s]; mask != 0u; mask &= mask - 1u) {
                                const int t = __ffs((int)mask) - 1;  // listed partners only
                                const float dx = __fsub_rn(p0, cen[(t * 3 + 0) * FK_ROWS]), dy = __fsub_rn(p1, cen[(t * 3 + 1) * FK_ROWS]),
                                            dz = __fsub_rn(p2, cen[(t * 3 + 2) * FK_ROWS]);
                                const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                                const float dist = __fsub_rn(__fsub_rn(n, g.sphere_r[s]), g.sphere_r[t]);
                                if (__fsub_rn(g.margin[f], dist) > 0.f) {
                                    const float inv = 1.f / fmaxf(n, 1e-12f);
                                    gx -= dx * inv; gy -= dy * inv; gz -= dz * inv;
                                    code |= 1 << t;
                                }
                            }
                        }
                        if (dec && row_on) dec[((long long)f * NI + i) * S + s] = code;
                        if (gx != 0.f || gy != 0.f || gz != 0.f) {
                            if (KIND == 1) {
                                const int nj = g.sphere_frame[s] < 7 ? g.sphere_frame[s] : 7;
#pragma unroll
                                for (int j = 0; j < 7; ++j) {
                                    if (j >= nj) break;
                                    const float rx = p0 - sc[(j * 3 + 0) * FK_ROWS], ry = p1 - sc[(j * 3 + 1) * FK_ROWS],
                                                rz = p2 - sc[(j * 3 + 2) * FK_ROWS];
                                    const float zx = sc[(21 + j * 3 + 0) * FK_ROWS], zy = sc[(21 + j * 3 + 1) * FK_ROWS],
                                                zz = sc[(21 + j * 3 + 2) * FK_ROWS];
                                    // (z x r) . g
                                    dq[j] += (zy * rz - zz * ry) * gx + (zz * rx - zx * rz) * gy + (zx * ry - zy * rx) * gz;
                                }
                            } else {
                                dq[0] += gx;
                                dq[1] += gy;
                                dq[2] += gz;
                            }
                        }
                    }
                    // fixed-order sum over the quad: (lane 0 + lane 1) + (lane 2 + lane 3)
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        dq[j] += __shfl_xor_sync(0xffffffffu, dq[j], 1);
                        dq[j] += __shfl_xor_sync(0xffffffffu, dq[j], 2);
                    }
                    if (row_on) {
                        float* dst = gq + ((long long)f * NI + i) * 8;
#pragma unroll
                        for (int j = 0; j < 7; ++j)
                            if ((j & 3) == ql) dst[j] = dq[j];
                    }
                }
                __syncwarp();  // the quad is done with this row's scratch before the next pass overwrites it
            }
        }
        __syncthreads();
        if (dbg) dbg[3] = clock64();  // interpolation + kinematic chain + lookups + J^T

        // ---------------- (E) finite-difference velocity mode only: raw velocity gradient of the GP cost, needed from the neighbours ----------------
        // (guides.py:77-79: the GP cost reaches the positions through v_h = (p_{h+1} - p_{h-1}) / 2dt as well)
        if (g.vel_fd && g.use_gp) {
            for (int h0 = 0; h0 < H; h0 += NTH / 8) {
                const int h = h0 + (tid >> 3), k = tid & 7;
                if (h < H && k < q) {
                    float gvk = 0.f;
                    if (h > 0 && h < H - 1) {
                        const float pm = xu[(h - 1) * D + k], pc = xu[h * D + k], pn = xu[(h + 1) * D + k];
                        const float vm = xu[(h - 1) * D + q + k], vc = xu[h * D + q + k], vn = xu[(h + 1) * D + q + k];
                        const float ep0 = pc - pm - g.dt * vm, ev0 = vc - vm, ep1 = pn - pc - g.dt * vc, ev1 = vn - vc;
                        const float uv0 = 2.f * (g.gp_b * ep0 + g.gp_c * ev0), uv1 = 2.f * (g.gp_b * ep1 + g.gp_c * ev1);
                        const float up1 = 2.f * (g.gp_a * ep1 + g.gp_b * ev1);
                        gvk = uv0 - uv1 - g.dt * up1;
                    }
                    gvs[h * 8 + k] = gvk;
                }
            }
            __syncthreads();
        }
        if (dbg) dbg[4] = clock64();

        if (last_it && tid == 0) pdl_launch_dependents();  // a programmatic dependent may start its prologue where SMs are free
        // ---------------- (F) per support row, 8 lanes (lane k < q owns coordinate k): for every collision cost the adjoint of the
        // interpolation (gather over the row's taps), 1/sigma^2, clip-by-norm (xor-shuffle tree over the 8 lanes), endpoint zeroing,
        // weight; then the GP prior (constant velocity) 3-tap stencil, the sum in cost order, and the update ----------------
        bool v_hi = false, v_mid = false;
        float var = 1.f;
        const bool use_var = a.model_var != nullptr || a.use_var_uniform;
        if (a.model_var != nullptr) var = a.model_var[b];
        else if (a.use_var_uniform) var = a.var_uniform;
        float* xout = a.x_out + (long long)b * H * Dio;
        for (int h0 = 0; h0 < H; h0 += NTH / 8) {
            const int h = h0 + (tid >> 3), k = tid & 7;
            const bool rowk = h < H && k < q;
            const bool inner = rowk && h > 0 && h < H - 1;  // gradient rows 0 and H-1 are zeroed by the guide manager
            float totp = 0.f, totv = 0.f;
            for (int f = 0; f < n_coll; ++f) {  // cost order, as the reference's `grad += w * g`
                float gsk = 0.f;
                if (rowk) {
                    const float* gqf = gq + (long long)f * NI * 8 + k;
                    if (use_taps) {
                        const float4 w0 = *reinterpret_cast<const float4*>(tapw + h * 8), w1 = *reinterpret_cast<const float4*>(tapw + h * 8 + 4);
                        const int4 t0 = *reinterpret_cast<const int4*>(tapi + h * 8), t1 = *reinterpret_cast<const int4*>(tapi + h * 8 + 4);
                        // a zero weight is a padding entry: skipped exactly as the on-the-fly form skips rows that do not touch h
                        gsk = w0.x != 0.f ? fmaf(w0.x, gqf[t0.x * 8], gsk) : gsk;
                        gsk = w0.y != 0.f ? fmaf(w0.y, gqf[t0.y * 8], gsk) : gsk;
                        gsk = w0.z != 0.f ? fmaf(w0.z, gqf[t0.z * 8], gsk) : gsk;
                        gsk = w0.w != 0.f ? fmaf(w0.w, gqf[t0.w * 8], gsk) : gsk;
                        gsk = w1.x != 0.f ? fmaf(w1.x, gqf[t1.x * 8], gsk) : gsk;
                        gsk = w1.y != 0.f ? fmaf(w1.y, gqf[t1.y * 8], gsk) : gsk;
                        gsk = w1.z != 0.f ? fmaf(w1.z, gqf[t1.z * 8], gsk) : gsk;
                        gsk = w1.w != 0.f ? fmaf(w1.w, gqf[t1.w * 8], gsk) : gsk;
                    } else {
                        // rows i with ratio * i in (h - 1, h + 1) touch h; floor / ceil leave one row of slack on each side for the
                        // rounding of the products (rows that do not touch h contribute with weight 0)
                        int lo_i = (int)floorf((float)(h - 1) * inv_ratio);
                        int hi_i = (int)ceilf((float)(h + 1) * inv_ratio);
                        if (lo_i < 0) lo_i = 0;
                        if (hi_i > NI - 1) hi_i = NI - 1;
                        for (int i = lo_i; i <= hi_i; ++i) {
                            const int i0 = i0s[i];
                            const int i1 = i0 + (i0 < H - 1 ? 1 : 0);
                            const float l1 = w1s[i];
                            const float cw = (i0 == h ? 1.f - l1 : 0.f) + (i1 == h ? l1 : 0.f);
                            gsk = cw != 0.f ? fmaf(cw, gqf[i * 8], gsk) : gsk;
                        }
                    }
                    gsk *= g.isig2[f];
                }
                float scale = 1.f;
                if (g.clip) {
                    const float t = rowk ? gsk + 1e-6f : 0.f;
                    float n2c = t * t;
                    n2c += __shfl_xor_sync(0xffffffffu, n2c, 1);
                    n2c += __shfl_xor_sync(0xffffffffu, n2c, 2);
                    n2c += __shfl_xor_sync(0xffffffffu, n2c, 4);
                    if (!pos_only) n2c += (float)(D - q) * (1e-6f * 1e-6f);  // the zero velocity half of the gradient, + 1e-6 each
                    scale = clip_scale(sqrtf(n2c), g.max_norm);
                }
                if (inner) totp += g.weight[f] * (scale * gsk);  // rows 0 and H-1: the guide manager zeroes them (adds +0)
            }
            float gpk = 0.f, gvk = 0.f, n2 = 0.f, n2v = 0.f;
            if (g.use_gp && inner) {
                const float pm = xu[(h - 1) * D + k], pc = xu[h * D + k], pn = xu[(h + 1) * D + k];
                const float vm = xu[(h - 1) * D + q + k], vc = xu[h * D + q + k], vn = xu[(h + 1) * D + q + k];
                const float ep0 = pc - pm - g.dt * vm, ev0 = vc - vm;  // e_{h-1}
                const float ep1 = pn - pc - g.dt * vc, ev1 = vn - vc;  // e_h
                const float up0 = 2.f * (g.gp_a * ep0 + g.gp_b * ev0), uv0 = 2.f * (g.gp_b * ep0 + g.gp_c * ev0);
                const float up1 = 2.f * (g.gp_a * ep1 + g.gp_b * ev1), uv1 = 2.f * (g.gp_b * ep1 + g.gp_c * ev1);
                gpk = up0 - up1;
                gvk = uv0 - uv1 - g.dt * up1;
                if (g.vel_fd) {
                    // + sum_h' dC/dv_h' * dv_h'/dp_h: v_{h-1} and v_{h+1} depend on p_h (rows 0 and H-1 of v are constants)
                    const float inv2dt = __fdiv_rn(1.f, __fmul_rn(2.f, g.dt));
                    gpk += (gvs[(h - 1) * 8 + k] - gvs[(h + 1) * 8 + k]) * inv2dt;
                    gvk = 0.f;
                }
                const float t0 = gpk + 1e-6f, t1 = gvk + 1e-6f;
                if (pos_only) { n2 = t0 * t0; n2v = t1 * t1; }  // position and velocity gradients are clipped separately
                else n2 = fmaf(t1, t1, t0 * t0);
            }
            n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
            n2 += __shfl_xor_sync(0xffffffffu, n2, 2);
            n2 += __shfl_xor_sync(0xffffffffu, n2, 4);
            n2v += __shfl_xor_sync(0xffffffffu, n2v, 1);
            n2v += __shfl_xor_sync(0xffffffffu, n2v, 2);
            n2v += __shfl_xor_sync(0xffffffffu, n2v, 4);
            if (!rowk) continue;
            if (g.use_gp && inner) {
                const float scale = g.clip ? clip_scale(sqrtf(n2), g.max_norm) : 1.f;
                const float scale_v = pos_only ? (g.clip ? clip_scale(sqrtf(n2v), g.max_norm) : 1.f) : scale;
                totp += g.w_gp * (scale * gpk);
                totv += g.w_gp * (scale_v * gvk);
            }
            if (pos_only) {  // gradient of the positions out, velocity trajectory updated in place (guides.py:110-112)
                xout[h * q + k] = -1.f * totp;
                if (!g.vel_fd) a.vel_io[((long long)b * H + h) * q + k] = __fsub_rn(xu[h * D + q + k], totv);
                continue;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int d = half ? q + k : k;
                const int i = h * D + d;
                float grad = -1.f * (half ? totv : totp);
                if (a.grad_only) { xout[i] = grad; continue; }
                if (use_var) grad = __fmul_rn(var, grad);
                float v = __fadd_rn(xn[i], grad);
                int hc = -1;
                for (int c = 0; c < a.n_hc; ++c)
                    if (a.hc_rows[c] == h) hc = c;
                if (hc >= 0) v = a.hc_vals[((long long)hc * a.B + b) * D + d];
                v_hi |= (v > 1.0001f) || (v < -1.0001f);
                v_mid |= (v > 1.f) || (v < -1.f);
                if (!last_it) { xn[i] = v; continue; }  // the next evaluation of this launch reads it from shared memory
                if (a.noise != nullptr && hc < 0) v = __fadd_rn(v, __fmul_rn(__fmul_rn(a.noise_sd, a.noise[(long long)b * H * D + i]), a.noise_mult));
                xout[i] = v;
                if (a.out2) a.out2[(long long)b * a.out2_bstride + i] = v;
            }
        }
        if (!last_it) {
            // clip flag of the next evaluation, resolved locally when possible (see the header comment)
            const int any_hi = __syncthreads_or(v_hi ? 1 : 0);
            const int any_mid = __syncthreads_or(v_mid ? 1 : 0);
            if (tid == 0) {
                if (any_hi) atomicOr(a.iter_flags + it + 1, 1);
                __threadfence();
                atomicAdd(a.iter_counters + it, 1u);
            }
            if (any_hi) flag = 1;
            else if (!any_mid) flag = 0;  // nothing beyond [-1, 1]: the clamp is the identity here whatever the batch decides
            else {
                if (tid == 0) {
                    const long long t0 = clock64();
                    while (*reinterpret_cast<volatile unsigned int*>(a.iter_counters + it) < gridDim.x) {
                        if (clock64() - t0 > 4000000000LL) __trap();  // a scheduling problem traps, never hangs
                    }
                    __threadfence();
                    s_flag = *reinterpret_cast<volatile int*>(a.iter_flags + it + 1);
                }
                __syncthreads();
                flag = s_flag;
            }
        } else if (a.flag_out != nullptr) {
            if (__syncthreads_or(v_hi ? 1 : 0) && tid == 0) atomicOr(a.flag_out, 1);
        }
        if (dbg) dbg[5 + (it < 2 ? it : 2)] = clock64();  // update written
    }  // evaluations
}

// ---------------------------------------------------------------------------------------------------
// Post-sampling evaluation (reference inference.py:288-326: get_trajs_collision_and_free,
// compute_collision_intensity_trajs, compute_smoothness, compute_path_length): forward-only reuse of the guide's
// interpolation + FK + field lookups. One CTA per (unnormalised) trajectory.
//   stats[b] = { #interpolated waypoints in collision, smoothness = sum_h |v_{h+1} - v_h|, path length = sum_h |p_{h+1} - p_h|,
//                minimum clearance min(sdf - radius) over waypoints, spheres and fields }
// A waypoint is in collision when any sphere has sdf - radius < margin in any object / boundary field.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FK_ROWS) eval_kernel(const __grid_constant__ GuideDev g, const float* __restrict__ x,
                                                       float* __restrict__ stats, float margin, int B, int H) {
    extern __shared__ __align__(16) float smem[];
    const int D = g.D, q = g.q_dim, NI = g.n_interp;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n_fields = g.n_grid + g.has_border;
    float* xu = smem;                      // [H][D]
    float* red = xu + H * D;               // [FK_ROWS] reduction scratch
    float* fk = red + FK_ROWS;             // [(42 + 3*n_spheres)][FK_ROWS]
    for (int i = tid; i < H * D; i += FK_ROWS) xu[i] = x[(long long)b * H * D + i];
    __syncthreads();
    int n_bad = 0;
    float clearance = 3.4e38f;
    const float ratio = NI > 1 ? (float)(H - 1) / (float)(NI - 1) : 0.f;
    for (int ibase = 0; ibase < NI; ibase += FK_ROWS) {
        const int i = ibase + tid;
        bool bad = false;
        if (i < NI) {
            float r = ratio * (float)i;
            int i0 = (int)r;
            if (i0 > H - 1) i0 = H - 1;
            const float l1 = fminf(fmaxf(r - (float)i0, 0.f), 1.f), l0 = 1.f - l1;
            const int i1 = i0 + (i0 < H - 1 ? 1 : 0);
            float qv[7];
#pragma unroll
            for (int k = 0; k < 7; ++k)
                qv[k] = k < q ? __fadd_rn(__fmul_rn(l0, xu[i0 * D + k]), __fmul_rn(l1, xu[i1 * D + k])) : 0.f;
            float* sc = fk + tid;
            float* cen = sc + 42 * FK_ROWS;
            fk_row(g, qv, sc, cen);
            for (int f = 0; f < n_fields; ++f)
                for (int s = 0; s < g.n_spheres; ++s) {
                    const float p[3] = {cen[(s * 3 + 0) * FK_ROWS], cen[(s * 3 + 1) * FK_ROWS], cen[(s * 3 + 2) * FK_ROWS]};
                    float sdf, gr[3];
                    if (f < g.n_grid) grid_fetch(g, f, grid_flat_index(g, f, p), sdf, gr);
                    else { int code; border_lookup(g, p, sdf, gr, code); }
                    const float c = __fsub_rn(sdf, g.sphere_r[s]);
                    clearance = fminf(clearance, c);
                    bad |= c < margin;
                }
        }
        n_bad += __syncthreads_count(bad ? 1 : 0);
    }
    // minimum clearance over the block
    red[tid] = clearance;
    __syncthreads();
    for (int o = FK_ROWS / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] = fminf(red[tid], red[tid + o]);
        __syncthreads();
    }
    const float min_clear = red[0];
    __syncthreads();
    // smoothness and path length on the support points, fixed-order sums
    float sm = 0.f, pl = 0.f;
    for (int h = tid; h < H - 1; h += FK_ROWS) {
        float dv = 0.f, dp = 0.f;
        for (int k = 0; k < q; ++k) {
            const float a = xu[(h + 1) * D + k] - xu[h * D + k];
            const float c = xu[(h + 1) * D + q + k] - xu[h * D + q + k];
            dp = fmaf(a, a, dp);
            dv = fmaf(c, c, dv);
        }
        sm += sqrtf(dv);
        pl += sqrtf(dp);
    }
    red[tid] = sm;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int k = 0; k < FK_ROWS; ++k) t += red[k]; stats[b * 4 + 1] = t; }
    __syncthreads();
    red[tid] = pl;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int k = 0; k < FK_ROWS; ++k) t += red[k];
        stats[b * 4 + 2] = t;
        stats[b * 4 + 0] = (float)n_bad;
        stats[b * 4 + 3] = min_clear;
    }
}

// FK unit-test kernel: one thread per configuration
__global__ void __launch_bounds__(FK_ROWS) fk_debug_kernel(const __grid_constant__ GuideDev g, const float* __restrict__ qin,
                                                           float* __restrict__ centers, int N) {
    extern __shared__ __align__(16) float smem[];
    const int n = blockIdx.x * FK_ROWS + threadIdx.x;
    float qv[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < N)
        for (int k = 0; k < g.q_dim && k < 7; ++k) qv[k] = qin[(long long)n * g.q_dim + k];
    float* sc = smem + threadIdx.x;
    float* cen = sc + 42 * FK_ROWS;
    fk_row(g, qv, sc, cen);
    if (n < N)
        for (int s = 0; s < g.n_spheres; ++s)
            for (int r3 = 0; r3 < 3; ++r3) centers[((long long)n * g.n_spheres + s) * 3 + r3] = cen[(s * 3 + r3) * FK_ROWS];
}

static size_t guide_smem_bytes(const GuideDev& g, int H) {
    size_t f = (size_t)2 * H * g.D + (size_t)g.n_coll * g.n_interp * 8 + (size_t)3 * H * 8 + 2 * (size_t)g.n_interp +
               (size_t)(42 + 3 * g.n_spheres) * FK_ROWS;
    return f * sizeof(float);
}

template <int KIND>
static int guide_occupancy(const GuideDev& g, int H, int device) {
    const size_t smem = guide_smem_bytes(g, H);
    int per_sm = 0, sms = 0;
    cudaFuncSetAttribute(guide_step_kernel<KIND, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, guide_step_kernel<KIND, 2>, GUIDE_THREADS, smem) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return per_sm * sms;
}

// CTAs of the guide kernel that can be resident at once (the selective grid wait of the multi-evaluation launch needs every
// CTA scheduled). One CTA per SM is left out of the count: the neighbouring kernels of the loop may still hold resources.
int guide_max_coresident(mpdb_guide* gd, int H) {
    if (gd->coresident_H == H) return gd->coresident;
    GuideDev g = make_dev(gd->cfg);
    const int n = g.robot_kind == 1 ? guide_occupancy<1>(g, H, gd->device) : guide_occupancy<0>(g, H, gd->device);
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, gd->device) != cudaSuccess) { cudaGetLastError(); return 0; }
    gd->coresident_H = H;
    gd->coresident = n >= 2 * sms ? n - sms : (n >= sms ? sms : 0);  // the one-per-SM instantiation (B <= #SMs) always fits when this is > 0
    return gd->coresident;
}

int guide_launch_step(mpdb_guide* gd, const GuideStepArgs& a_in, cudaStream_t stream) {
    GuideDev g = make_dev(gd->cfg);
    GuideStepArgs a = a_in;
    MPDB_REQUIRE(g.D <= MPDB_MAX_STATE_DIM && g.q_dim <= 7, "guide: state dim too large");
    MPDB_REQUIRE(g.n_interp >= 1, "guide: n_interp must be >= 1");
    const size_t smem = guide_smem_bytes(g, a.H);
    MPDB_REQUIRE(smem <= 220 * 1024, "guide: trajectory does not fit in shared memory");
    MPDB_REQUIRE(a.n_iters <= 1 || (a.iter_flags && a.iter_counters && !a.grad_only && a.B <= guide_max_coresident(gd, a.H)),
                 "guide: several evaluations per launch need flag / counter scratch and a co-resident grid");
    const bool pos_only = a.vel_io != nullptr || g.vel_fd;
    MPDB_REQUIRE(!pos_only || (a.grad_only && a.n_iters <= 1), "guide: position-only mode returns the gradient only");
    MPDB_REQUIRE(!g.vel_fd || a.vel_io == nullptr, "guide: finite-difference velocities replace the velocity trajectory");
    a.dep_count = reinterpret_cast<unsigned int*>(gd->flags + 8);
    const int n_evals = a.n_iters > 1 ? a.n_iters : 1;
    if (gd->dec_buf) {
        MPDB_REQUIRE(a.B == gd->dec_batch, "guide: decision recording was set up for another batch size");
        MPDB_REQUIRE(gd->dec_count + n_evals <= gd->dec_capacity, "guide: decision buffer is full");
        a.dec = gd->dec_buf + gd->dec_count * (long long)a.B * g.n_coll * g.n_interp * g.n_spheres;
        gd->dec_count += n_evals;
    }
#define MPDB_GUIDE_LAUNCH(K, O)                                                                                              \
    {                                                                                                                        \
        static unsigned long long configured = 0ull;                                                                         \
        if (mpdb::first_use_on_device(configured)) {                                                                         \
            MPDB_CHECK_CUDA(cudaFuncSetAttribute(guide_step_kernel<K, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
        }                                                                                                                    \
        MPDB_CHECK_CUDA(launch_kernel_pdl(guide_step_kernel<K, O>, dim3(a.B), dim3(GUIDE_THREADS), smem, stream,             \
                                          g_use_pdl || a.pdl != 0, g, a));                                                    \
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, gd->device);
    const bool one_per_sm = a.B <= sms;
    if (g.robot_kind == 1) { if (one_per_sm) MPDB_GUIDE_LAUNCH(1, 1) else MPDB_GUIDE_LAUNCH(1, 2) }
    else { if (one_per_sm) MPDB_GUIDE_LAUNCH(0, 1) else MPDB_GUIDE_LAUNCH(0, 2) }
#undef MPDB_GUIDE_LAUNCH
    MPDB_LAUNCH_CHECK();
    return 0;
}

__global__ void range_flag_kernel(const float* __restrict__ x, long long n, int* flag) {
    bool viol = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = x[i];
        viol |= (v > 1.0001f) || (v < -1.0001f);
    }
    if (__syncthreads_or(viol ? 1 : 0) && threadIdx.x == 0) atomicOr(flag, 1);
}

int guide_launch_flag(const float* x, long long n, int* flag, cudaStream_t stream) {
    MPDB_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    range_flag_kernel<<<blocks, 256, 0, stream>>>(x, n, flag);
    MPDB_LAUNCH_CHECK();
    return 0;
}

int guide_device(mpdb_guide* g) { return g->device; }
int guide_state_dim(mpdb_guide* g) { return 2 * g->cfg.q_dim; }
bool guide_recording(mpdb_guide* g) { return g->dec_buf != nullptr; }

// ---------------------------------------------------------------------------------------------------
// SDF voxel grid from analytic primitives: texel = {sdf, d sdf/dx, ...} at the node
// ---------------------------------------------------------------------------------------------------
__global__ void sdf_grid_kernel(int dim, int nx, int ny, int nz, float lox, float loy, float loz, float cell,
                                const float* __restrict__ spheres, int ns, const float* __restrict__ boxes, int nb,
                                float* __restrict__ tex) {
    const long long n = (long long)nx * ny * (dim == 3 ? nz : 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int idx[3];
        if (dim == 3) { idx[2] = (int)(i % nz); idx[1] = (int)((i / nz) % ny); idx[0] = (int)(i / ((long long)nz * ny)); }
        else { idx[1] = (int)(i % ny); idx[0] = (int)(i / ny); idx[2] = 0; }
        const float lo[3] = {lox, loy, loz};
        float p[3] = {0.f, 0.f, 0.f};
        for (int d = 0; d < dim; ++d) p[d] = __fadd_rn(lo[d], __fmul_rn((float)idx[d], cell));
        float best = INFINITY, bg[3] = {0.f, 0.f, 0.f};
        for (int s = 0; s < ns; ++s) {
            const float* sp = spheres + s * (dim + 1);
            float dv[3] = {0.f, 0.f, 0.f}, n2 = 0.f;
            for (int d = 0; d < dim; ++d) { dv[d] = __fsub_rn(p[d], sp[d]); n2 = __fadd_rn(n2, __fmul_rn(dv[d], dv[d])); }
            float nn = sqrtf(n2);
            float val = __fsub_rn(nn, sp[dim]);
            if (val < best) {
                best = val;
                float den = fmaxf(nn, 1e-12f);
                for (int d = 0; d < 3; ++d) bg[d] = d < dim ? __fdiv_rn(dv[d], den) : 0.f;
            }
        }
        for (int k = 0; k < nb; ++k) {
            const float* bx = boxes + k * 2 * dim;
            float dv[3] = {0.f, 0.f, 0.f}, qv[3], qp[3] = {0.f, 0.f, 0.f}, o2 = 0.f, qmax = -INFINITY;
            int amax = 0;
            for (int d = 0; d < dim; ++d) {
                dv[d] = __fsub_rn(p[d], bx[d]);
                qv[d] = __fsub_rn(fabsf(dv[d]), bx[dim + d]);
                qp[d] = fmaxf(qv[d], 0.f);
                o2 = __fadd_rn(o2, __fmul_rn(qp[d], qp[d]));
                if (qv[d] > qmax) { qmax = qv[d]; amax = d; }
            }
            float outside = sqrtf(o2);
            float val = __fadd_rn(outside, fminf(qmax, 0.f));
            if (val < best) {
                best = val;
                float den = fmaxf(outside, 1e-12f);
                for (int d = 0; d < 3; ++d) {
                    float sg = dv[d] >= 0.f ? 1.f : -1.f;
                    if (d >= dim) bg[d] = 0.f;
                    else if (outside > 0.f) bg[d] = __fdiv_rn(__fmul_rn(sg, qp[d]), den);
                    else bg[d] = (d == amax) ? sg : 0.f;
                }
            }
        }
        float* t = tex + i * (1 + dim);
        t[0] = best;
        for (int d = 0; d < dim; ++d) t[1 + d] = bg[d];
    }
}

}  // namespace mpdb

using namespace mpdb;

extern "C" int mpdb_guide_create(const mpdb_guide_config* cfg, int device, mpdb_guide** out) {
    MPDB_REQUIRE(cfg && out, "mpdb_guide_create: null argument");
    MPDB_REQUIRE(cfg->robot_kind == 0 || cfg->robot_kind == 1, "guide: unknown robot kind");
    MPDB_REQUIRE(cfg->q_dim >= 1 && cfg->q_dim <= 7 && 2 * cfg->q_dim <= MPDB_MAX_STATE_DIM, "guide: bad q_dim");
    MPDB_REQUIRE(cfg->robot_kind == 0 || (cfg->q_dim == 7 && cfg->ws_dim == 3), "guide: Panda needs q_dim 7, ws_dim 3");
    MPDB_REQUIRE(cfg->robot_kind == 1 || cfg->ws_dim == cfg->q_dim, "guide: point mass needs ws_dim == q_dim");
    MPDB_REQUIRE(cfg->ws_dim == 2 || cfg->ws_dim == 3, "guide: ws_dim must be 2 or 3");
    MPDB_REQUIRE(cfg->n_spheres >= 1 && cfg->n_spheres <= MPDB_MAX_SPHERES, "guide: bad sphere count");
    MPDB_REQUIRE(cfg->n_grid_fields >= 0 && cfg->n_grid_fields <= MPDB_MAX_GRID_FIELDS, "guide: bad field count");
    for (int f = 0; f < cfg->n_grid_fields; ++f) {
        MPDB_REQUIRE(cfg->grid_texels[f] != nullptr && cfg->grid_cell[f] > 0.f, "guide: grid field without texels / cell size");
        for (int d = 0; d < cfg->ws_dim; ++d) MPDB_REQUIRE(cfg->grid_shape[f][d] >= 1, "guide: bad grid shape");
        MPDB_REQUIRE(cfg->sigma_grid[f] > 0.f, "guide: sigma_coll must be positive");
    }
    MPDB_REQUIRE(!cfg->has_border || cfg->sigma_border > 0.f, "guide: sigma_coll must be positive");
    MPDB_REQUIRE(!cfg->has_self || cfg->sigma_self > 0.f, "guide: sigma_coll must be positive");
    MPDB_REQUIRE(!cfg->has_self || cfg->robot_kind == 1, "guide: the self-collision field needs an articulated robot");
    for (int s = 0; s < cfg->n_spheres && cfg->has_self; ++s)
        for (int t = 0; t < MPDB_MAX_SPHERES; ++t)
            if ((cfg->self_pairs[s] >> t) & 1u)
                MPDB_REQUIRE(t < cfg->n_spheres && t != s && ((cfg->self_pairs[t] >> s) & 1u), "guide: self_pairs must be a symmetric relation on the spheres");
    MPDB_REQUIRE(!cfg->use_gp || (cfg->dt > 0.f && cfg->sigma_gp > 0.f), "guide: dt and sigma_gp must be positive");
    MPDB_REQUIRE(!cfg->vel_from_fd || cfg->dt > 0.f, "guide: finite-difference velocities need dt > 0");
    for (int s = 0; s < cfg->n_spheres && cfg->robot_kind == 1; ++s)
        MPDB_REQUIRE(cfg->sphere_frame[s] >= 1 && cfg->sphere_frame[s] <= 8, "guide: sphere frame out of range");
    MPDB_ENTER_DEVICE(device);
    mpdb_guide* g = new mpdb_guide();
    g->cfg = *cfg;
    g->device = device;
    g->flags = nullptr;
    if (cudaMalloc(&g->flags, 64 * sizeof(int)) != cudaSuccess || cudaMemset(g->flags, 0, 64 * sizeof(int)) != cudaSuccess) {
        delete g;
        mpdb::set_error("mpdb_guide_create: cudaMalloc failed");
        return 1;
    }
    *out = g;
    return 0;
}

extern "C" void mpdb_guide_destroy(mpdb_guide* g) {
    if (!g) return;
    mpdb::DeviceGuard dg(g->device);
    cudaFree(g->flags);
    delete g;
}

extern "C" int mpdb_guide_grad(mpdb_guide* g, const float* x, float* grad, int32_t B, int32_t H, void* stream) {
    MPDB_REQUIRE(g && x && grad && B > 0 && H > 1, "mpdb_guide_grad: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(g->device);
    if (guide_launch_flag(x, (long long)B * H * 2 * g->cfg.q_dim, g->flags, st)) return 1;
    GuideStepArgs a;
    memset(&a, 0, sizeof(a));
    a.x_in = x;
    a.x_out = grad;
    a.grad_only = 1;
    a.flag_in = g->flags;
    a.B = B;
    a.H = H;
    return guide_launch_step(g, a, st);
}

extern "C" int mpdb_guide_grad_pos(mpdb_guide* g, const float* x_pos, float* velocity, float* grad, int32_t B, int32_t H,
                                   void* stream) {
    MPDB_REQUIRE(g && x_pos && grad && B > 0 && H > 1, "mpdb_guide_grad_pos: bad argument");
    MPDB_REQUIRE((velocity != nullptr) != (g->cfg.vel_from_fd != 0),
                 "mpdb_guide_grad_pos: pass the velocity trajectory, or NULL with a guide configured for finite-difference velocities");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(g->device);
    if (guide_launch_flag(x_pos, (long long)B * H * g->cfg.q_dim, g->flags, st)) return 1;  // the clip looks at the positions only
    GuideStepArgs a;
    memset(&a, 0, sizeof(a));
    a.x_in = x_pos;
    a.x_out = grad;
    a.grad_only = 1;
    a.vel_io = velocity;
    a.flag_in = g->flags;
    a.B = B;
    a.H = H;
    return guide_launch_step(g, a, st);
}

extern "C" int mpdb_guide_steps(mpdb_guide* g, float* x, int32_t n_steps, const float* model_var, int32_t n_hc,
                                const int32_t* hc_rows, const float* hc_vals, int32_t B, int32_t H, void* stream) {
    return mpdb_guide_steps_chain(g, x, n_steps, model_var, n_hc, hc_rows, hc_vals, nullptr, 0, B, H, stream);
}

extern "C" int mpdb_guide_steps_chain(mpdb_guide* g, float* x, int32_t n_steps, const float* model_var, int32_t n_hc,
                                      const int32_t* hc_rows, const float* hc_vals, float* chain_out, int64_t chain_step_stride,
                                      int32_t B, int32_t H, void* stream) {
    MPDB_REQUIRE(g && x && B > 0 && H > 1 && n_steps >= 0, "mpdb_guide_steps: bad argument");
    MPDB_REQUIRE(!chain_out || chain_step_stride >= (int64_t)B * H * 2 * g->cfg.q_dim, "mpdb_guide_steps_chain: chain step stride too small");
    MPDB_REQUIRE(n_hc >= 0 && n_hc <= MPDB_MAX_HARD_CONDS, "mpdb_guide_steps: too many hard conditions");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(g->device);
    if (n_steps == 0) return 0;
    if (guide_launch_flag(x, (long long)B * H * 2 * g->cfg.q_dim, g->flags, st)) return 1;
    if (chain_out == nullptr && n_steps >= 2 && n_steps <= 20 && g->dec_buf == nullptr && B <= guide_max_coresident(g, H)) {
        // all evaluations in ONE launch (the trajectory stays in shared memory; guide_step_kernel resolves the clip flag per CTA)
        MPDB_CHECK_CUDA(cudaMemsetAsync(g->flags + 16, 0, sizeof(int) * 44, st));
        MPDB_CHECK_CUDA(cudaMemcpyAsync(g->flags + 16, g->flags, sizeof(int), cudaMemcpyDeviceToDevice, st));
        GuideStepArgs a;
        memset(&a, 0, sizeof(a));
        a.x_in = x;
        a.x_out = x;
        a.flag_in = g->flags + 16;
        a.n_iters = n_steps;
        a.iter_flags = g->flags + 16;
        a.iter_counters = reinterpret_cast<unsigned int*>(g->flags + 40);
        a.model_var = model_var;
        a.n_hc = n_hc;
        for (int k = 0; k < n_hc; ++k) a.hc_rows[k] = hc_rows[k];
        a.hc_vals = hc_vals;
        a.B = B;
        a.H = H;
        return guide_launch_step(g, a, st);
    }
    for (int it = 0; it < n_steps; ++it) {
        GuideStepArgs a;
        memset(&a, 0, sizeof(a));
        a.x_in = x;
        a.x_out = x;  // in place: each CTA reads its whole trajectory into shared memory before writing
        a.flag_in = g->flags + (it & 1);
        a.flag_out = (it + 1 < n_steps) ? g->flags + ((it + 1) & 1) : nullptr;
        if (a.flag_out) MPDB_CHECK_CUDA(cudaMemsetAsync(a.flag_out, 0, sizeof(int), st));
        a.model_var = model_var;
        a.n_hc = n_hc;
        for (int k = 0; k < n_hc; ++k) a.hc_rows[k] = hc_rows[k];
        a.hc_vals = hc_vals;
        if (chain_out) { a.out2 = chain_out + (long long)it * chain_step_stride; a.out2_bstride = (long long)H * 2 * g->cfg.q_dim; }
        a.B = B;
        a.H = H;
        if (guide_launch_step(g, a, st)) return 1;
    }
    return 0;
}

extern "C" int mpdb_sdf_grid_build(int32_t dim, const int32_t* shape, const float* lo, float cell, const float* spheres,
                                   int32_t n_spheres, const float* boxes, int32_t n_boxes, float* texels_out,
                                   void* stream) {
    MPDB_REQUIRE(dim == 2 || dim == 3, "mpdb_sdf_grid_build: dim must be 2 or 3");
    MPDB_REQUIRE(shape && lo && texels_out, "mpdb_sdf_grid_build: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    float *dsp = nullptr, *dbx = nullptr;
    size_t ssz = sizeof(float) * (size_t)n_spheres * (dim + 1), bsz = sizeof(float) * (size_t)n_boxes * 2 * dim;
    if (ssz) { MPDB_CHECK_CUDA(cudaMalloc(&dsp, ssz)); MPDB_CHECK_CUDA(cudaMemcpyAsync(dsp, spheres, ssz, cudaMemcpyHostToDevice, st)); }
    if (bsz) { MPDB_CHECK_CUDA(cudaMalloc(&dbx, bsz)); MPDB_CHECK_CUDA(cudaMemcpyAsync(dbx, boxes, bsz, cudaMemcpyHostToDevice, st)); }
    long long n = (long long)shape[0] * shape[1] * (dim == 3 ? shape[2] : 1);
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    sdf_grid_kernel<<<blocks, 256, 0, st>>>(dim, shape[0], shape[1], dim == 3 ? shape[2] : 1, lo[0], lo[1],
                                            dim == 3 ? lo[2] : 0.f, cell, dsp, n_spheres, dbx, n_boxes, texels_out);
    mpdb::g_launch_count.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(dsp);
    cudaFree(dbx);
    MPDB_CHECK_CUDA(e);
    MPDB_CHECK_CUDA(e2);
    return 0;
}

// Average device time of one guide evaluation (CUDA events on `stream`), for bench.py's roofline object.
extern "C" int mpdb_profile_guide(mpdb_guide* g, float* x, int32_t B, int32_t H, int32_t reps, float* ms_out, void* stream) {
    MPDB_REQUIRE(g && x && ms_out && B > 0 && H > 1 && reps > 0, "mpdb_profile_guide: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(g->device);
    if (guide_launch_flag(x, (long long)B * H * 2 * g->cfg.q_dim, g->flags, st)) return 1;
    GuideStepArgs a;
    memset(&a, 0, sizeof(a));
    a.x_in = x;
    a.x_out = x;
    a.flag_in = g->flags;
    a.B = B;
    a.H = H;
    cudaEvent_t ev0, ev1;
    MPDB_CHECK_CUDA(cudaEventCreate(&ev0));
    MPDB_CHECK_CUDA(cudaEventCreate(&ev1));
    if (guide_launch_step(g, a, st)) return 1;
    MPDB_CHECK_CUDA(cudaEventRecord(ev0, st));
    for (int r = 0; r < reps; ++r)
        if (guide_launch_step(g, a, st)) return 1;
    MPDB_CHECK_CUDA(cudaEventRecord(ev1, st));
    MPDB_CHECK_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    MPDB_CHECK_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    *ms_out = ms / reps;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (getenv("MPDB_GUIDE_TIMELINE")) {  // debug: phase stamps of CTA 0
        long long* d = nullptr;
        long long h[8] = {0};
        MPDB_CHECK_CUDA(cudaMalloc(&d, sizeof(h)));
        MPDB_CHECK_CUDA(cudaMemset(d, 0, sizeof(h)));
        a.dbg = d;
        if (guide_launch_step(g, a, st)) return 1;
        MPDB_CHECK_CUDA(cudaStreamSynchronize(st));
        MPDB_CHECK_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(d);
        const int idx[5] = {0, 1, 3, 4, 5};
        const char* names[5] = {"start", "loaded", "rows (interp+fk+lookups+J^T)", "adjoint+clip", "gp+update"};
        for (int k = 1; k < 5; ++k)
            fprintf(stderr, "[guide timeline] %-30s +%.2f us (t = %.2f)\n", names[k], (h[idx[k]] - h[idx[k - 1]]) / 1965.0, (h[idx[k]] - h[0]) / 1965.0);
    }
    return 0;
}

// Average device time of ONE LAUNCH running n_evals guide evaluations in place (the loop's fused guide_gradient_steps launch).
extern "C" int mpdb_profile_guide_steps(mpdb_guide* g, float* x, int32_t n_evals, int32_t B, int32_t H, int32_t reps, float* ms_out,
                                        void* stream) {
    MPDB_REQUIRE(g && x && ms_out && B > 0 && H > 1 && reps > 0 && n_evals >= 2 && n_evals <= 20, "mpdb_profile_guide_steps: bad argument");
    MPDB_REQUIRE(B <= guide_max_coresident(g, H), "mpdb_profile_guide_steps: the batch is not co-resident (the loop launches one evaluation at a time)");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(g->device);
    cudaEvent_t ev0, ev1;
    MPDB_CHECK_CUDA(cudaEventCreate(&ev0));
    MPDB_CHECK_CUDA(cudaEventCreate(&ev1));
    for (int r = -1; r < reps; ++r) {
        if (r == 0) MPDB_CHECK_CUDA(cudaEventRecord(ev0, st));
        // what the loop does per guided step: zero the flag / counter scratch (part of one memset per loop there), one launch
        MPDB_CHECK_CUDA(cudaMemsetAsync(g->flags + 16, 0, sizeof(int) * 44, st));
        GuideStepArgs a;
        memset(&a, 0, sizeof(a));
        a.x_in = x;
        a.x_out = x;
        a.flag_in = g->flags + 16;
        a.n_iters = n_evals;
        a.iter_flags = g->flags + 16;
        a.iter_counters = reinterpret_cast<unsigned int*>(g->flags + 40);
        a.B = B;
        a.H = H;
        if (guide_launch_step(g, a, st)) return 1;
    }
    MPDB_CHECK_CUDA(cudaEventRecord(ev1, st));
    MPDB_CHECK_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    MPDB_CHECK_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    *ms_out = ms / reps;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return 0;
}

extern "C" int mpdb_eval_trajectories(mpdb_guide* gd, const float* x_unnormalized, float* stats, float margin, int32_t B,
                                      int32_t H, void* stream) {
    MPDB_REQUIRE(gd && x_unnormalized && stats && B > 0 && H > 1, "mpdb_eval_trajectories: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(gd->device);
    GuideDev g = make_dev(gd->cfg);
    const size_t smem = sizeof(float) * ((size_t)H * g.D + FK_ROWS + (size_t)(42 + 3 * g.n_spheres) * FK_ROWS);
    MPDB_REQUIRE(smem <= 200 * 1024, "mpdb_eval_trajectories: trajectory does not fit in shared memory");
    MPDB_REQUIRE(2 * gd->cfg.q_dim == g.D, "mpdb_eval_trajectories: trajectories must carry positions and velocities");
    static unsigned long long configured = 0ull;
    if (mpdb::first_use_on_device(configured)) {
        MPDB_CHECK_CUDA(cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    eval_kernel<<<B, FK_ROWS, smem, st>>>(g, x_unnormalized, stats, margin, B, H);
    MPDB_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpdb_guide_record_decisions(mpdb_guide* g, int32_t* dev_buf, int64_t capacity_evals, int32_t B) {
    MPDB_REQUIRE(g, "mpdb_guide_record_decisions: null guide");
    MPDB_REQUIRE(dev_buf == nullptr || (capacity_evals > 0 && B > 0), "mpdb_guide_record_decisions: bad capacity / batch");
    g->dec_buf = dev_buf;
    g->dec_capacity = dev_buf ? capacity_evals : 0;
    g->dec_batch = dev_buf ? B : 0;
    g->dec_count = 0;
    return 0;
}

extern "C" int64_t mpdb_guide_decisions_recorded(mpdb_guide* g) { return g ? (int64_t)g->dec_count : -1; }

extern "C" int64_t mpdb_guide_batch_dependent_clamps(mpdb_guide* g, int32_t reset) {
    if (!g) return -1;
    mpdb::DeviceGuard dg(g->device);
    unsigned int n = 0;
    if (!dg.ok || cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&n, g->flags + 8, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset) cudaMemset(g->flags + 8, 0, sizeof(n));
    return (int64_t)n;
}

extern "C" int mpdb_guide_max_coresident(mpdb_guide* g, int32_t H) {
    if (!g) return 0;
    mpdb::DeviceGuard dg(g->device);
    return dg.ok ? guide_max_coresident(g, H) : 0;
}

extern "C" int mpdb_guide_num_collision_costs(mpdb_guide* g) {
    if (!g) return -1;
    return g->cfg.n_grid_fields + (g->cfg.has_border ? 1 : 0) + ((g->cfg.has_self && g->cfg.robot_kind == 1) ? 1 : 0);
}

extern "C" int mpdb_debug_fk(mpdb_guide* gd, const float* q, float* centers, int32_t N, void* stream) {
    MPDB_REQUIRE(gd && q && centers && N > 0, "mpdb_debug_fk: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_ENTER_DEVICE(gd->device);
    GuideDev g = make_dev(gd->cfg);
    const size_t smem = sizeof(float) * (size_t)(42 + 3 * g.n_spheres) * FK_ROWS;
    static unsigned long long configured = 0ull;
    if (mpdb::first_use_on_device(configured))
        MPDB_CHECK_CUDA(cudaFuncSetAttribute(fk_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    fk_debug_kernel<<<(N + FK_ROWS - 1) / FK_ROWS, FK_ROWS, smem, st>>>(g, q, centers, N);
    MPDB_LAUNCH_CHECK();
    return 0;
}
