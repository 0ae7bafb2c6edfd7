// Cost-gradient guide: GuideManagerTrajectoriesWithVelocity.forward (reference guides.py:173-236) with the
// restated CostComposite([CostCollision(field)...], CostGPTrajectory) of SURVEY.md Appendix C, and the
// update x <- x + guide(x) of guide_gradient_steps (sample_functions.py:65-83), as ONE kernel per
// guide evaluation:
//
//   unnormalise (LimitsNormalizer.unnormalize incl. the batch-global clip flag)  ->  linear interpolation
//   H -> n_interp  ->  FK of the collision spheres  ->  nearest-texel SDF lookup {sdf, grad} per field  ->
//   hinge  ->  hand-derived adjoint (J^T through the chain, two-tap scatter of the interpolation written
//   as a gather)  ->  per-cost clip-by-norm, endpoint zeroing, weighting  ->  GP-prior 3-tap stencil
//   gradient  ->  x + grad, hard conditioning [-> + std * noise * noise_std, hard conditioning].
//
// One CTA per trajectory; the trajectory, its unnormalised copy and every per-cost gradient stay in
// shared memory. The only global traffic is x in/out (+noise) and the SDF texel gathers.
#include <math.h>

#include <vector>

#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

struct mpdb_guide {
    mpdb_guide_config cfg;
    int device;
    int* flags;  // device scratch: [2] flags for the standalone entry points
};

namespace mpdb {

constexpr int GUIDE_THREADS = 512;
constexpr int FK_ROWS = 128;  // interpolated rows per pass (one FK thread each)
constexpr int NSG = 4;        // sphere groups per row sharing the lookup + adjoint work
constexpr int SPG = (MPDB_MAX_SPHERES + NSG - 1) / NSG;  // spheres per group (upper bound)

struct GuideDev {
    int robot_kind, q_dim, ws_dim, n_spheres, D;
    int sphere_frame[MPDB_MAX_SPHERES];
    float sphere_off[MPDB_MAX_SPHERES][3];
    float sphere_r[MPDB_MAX_SPHERES];
    int frame_begin[10];                   // spheres attached to frame f (1..8) are frame_sphere[frame_begin[f-1] .. frame_begin[f])
    int frame_sphere[MPDB_MAX_SPHERES];    // sphere indices sorted by frame (stable)
    float mins[MPDB_MAX_STATE_DIM], range[MPDB_MAX_STATE_DIM];
    int n_grid;
    const float* tex[MPDB_MAX_GRID_FIELDS];
    int gshape[3];
    float glo[3];
    float cell;
    int has_border;
    float blo[3], bhi[3];
    float margin, dt;
    float w_grid[MPDB_MAX_GRID_FIELDS], w_border, w_gp;
    int use_gp, clip;
    float max_norm;
    int n_interp;
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
    for (int i = 0; i < d.D; ++i) {
        d.mins[i] = c.mins[i];
        d.range[i] = c.maxs[i] - c.mins[i];  // fp32 subtraction, as `self.maxs - self.mins`
    }
    d.n_grid = c.n_grid_fields;
    for (int i = 0; i < c.n_grid_fields; ++i) {
        d.tex[i] = c.grid_texels[i];
        d.w_grid[i] = c.weight_grid[i];
    }
    for (int k = 0; k < 3; ++k) {
        d.gshape[k] = c.grid_shape[k];
        d.glo[k] = c.grid_lo[k];
        d.blo[k] = c.border_lo[k];
        d.bhi[k] = c.border_hi[k];
    }
    d.cell = c.grid_cell;
    d.has_border = c.has_border;
    d.margin = c.cutoff_margin;
    d.dt = c.dt;
    d.w_border = c.weight_border;
    d.w_gp = c.weight_gp;
    d.use_gp = c.use_gp;
    d.clip = c.clip_grad;
    d.max_norm = c.max_grad_norm;
    d.n_interp = c.n_interp;
    const double dt = (double)c.dt, s = 1.0 / ((double)c.sigma_gp * (double)c.sigma_gp);
    d.gp_a = (float)(12.0 / (dt * dt * dt) * s);
    d.gp_b = (float)(-6.0 / (dt * dt) * s);
    d.gp_c = (float)(4.0 / dt * s);
    for (int i = 0; i < 7; ++i) {
        for (int k = 0; k < 3; ++k) d.joint_xyz[i][k] = (float)kPandaXYZ[i][k];
        d.joint_cr[i] = (float)cos(kPandaRoll[i]);
        d.joint_sr[i] = (float)sin(kPandaRoll[i]);
    }
    for (int k = 0; k < 3; ++k) d.flange[k] = (float)kPandaFlange[k];
    return d;
}

// Forward kinematics of one interpolated row: joint origins (3 x 7), joint axes (3 x 7) and collision-sphere centres
// (3 x n_spheres) into per-row scratch with stride FK_ROWS. Panda chain: T_i = T_{i-1} Trans(xyz_i) Rx(roll_i) Rz(q_i)
// (SURVEY Appendix E); point mass: centre = q.
__device__ __forceinline__ void fk_chain(const GuideDev& g, const float (&sq)[7], const float (&cq)[7], float* sc, float* cen) {
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};  // row-major
    float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
#pragma unroll
        for (int r3 = 0; r3 < 3; ++r3)
            o[r3] += R[r3 * 3 + 0] * g.joint_xyz[j][0] + R[r3 * 3 + 1] * g.joint_xyz[j][1] + R[r3 * 3 + 2] * g.joint_xyz[j][2];
        const float cr = g.joint_cr[j], sr = g.joint_sr[j];
#pragma unroll
        for (int r3 = 0; r3 < 3; ++r3) {
            float c0 = R[r3 * 3 + 0], c1 = R[r3 * 3 + 1], c2 = R[r3 * 3 + 2];
            float a1 = c1 * cr + c2 * sr;   // (R Rx) column 1
            float a2 = -c1 * sr + c2 * cr;  // (R Rx) column 2
            R[r3 * 3 + 0] = cq[j] * c0 + sq[j] * a1;
            R[r3 * 3 + 1] = -sq[j] * c0 + cq[j] * a1;
            R[r3 * 3 + 2] = a2;
        }
#pragma unroll
        for (int r3 = 0; r3 < 3; ++r3) {
            sc[(j * 3 + r3) * FK_ROWS] = o[r3];
            sc[(21 + j * 3 + r3) * FK_ROWS] = R[r3 * 3 + 2];  // joint axis = third column
        }
        for (int s = 0; s < g.n_spheres; ++s)
            if (g.sphere_frame[s] == j + 1)
                for (int r3 = 0; r3 < 3; ++r3)
                    cen[(s * 3 + r3) * FK_ROWS] = o[r3] + R[r3 * 3 + 0] * g.sphere_off[s][0] +
                                                  R[r3 * 3 + 1] * g.sphere_off[s][1] + R[r3 * 3 + 2] * g.sphere_off[s][2];
    }
#pragma unroll
    for (int r3 = 0; r3 < 3; ++r3)
        o[r3] += R[r3 * 3 + 0] * g.flange[0] + R[r3 * 3 + 1] * g.flange[1] + R[r3 * 3 + 2] * g.flange[2];
    for (int s = 0; s < g.n_spheres; ++s)
        if (g.sphere_frame[s] == 8)
            for (int r3 = 0; r3 < 3; ++r3)
                cen[(s * 3 + r3) * FK_ROWS] = o[r3] + R[r3 * 3 + 0] * g.sphere_off[s][0] + R[r3 * 3 + 1] * g.sphere_off[s][1] +
                                              R[r3 * 3 + 2] * g.sphere_off[s][2];
}

// One matrix row of the chain: row r3 of R and component r3 of the origin depend only on row r3 of the previous frame,
// so three threads per interpolated row run the chain independently (3x shorter dependency chain).
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
        // the spheres attached to this frame (host-sorted list: no scan over all spheres per joint)
        for (int t = g.frame_begin[j]; t < g.frame_begin[j + 1]; ++t) {
            const int s = g.frame_sphere[t];
            cen[(s * 3 + r3) * FK_ROWS] = o + c0 * g.sphere_off[s][0] + c1 * g.sphere_off[s][1] + c2 * g.sphere_off[s][2];
        }
    }
    o += c0 * g.flange[0] + c1 * g.flange[1] + c2 * g.flange[2];
    for (int t = g.frame_begin[7]; t < g.frame_begin[8]; ++t) {
        const int s = g.frame_sphere[t];
        cen[(s * 3 + r3) * FK_ROWS] = o + c0 * g.sphere_off[s][0] + c1 * g.sphere_off[s][1] + c2 * g.sphere_off[s][2];
    }
}

__device__ __forceinline__ void fk_row(const GuideDev& g, const float (&qv)[7], float* sc, float* cen) {
    if (g.robot_kind == 1) {
        float sq[7], cq[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) sincosf(qv[j], &sq[j], &cq[j]);
        fk_chain(g, sq, cq, sc, cen);
    } else {
        for (int s = 0; s < g.n_spheres; ++s)
            for (int r3 = 0; r3 < 3; ++r3) cen[(s * 3 + r3) * FK_ROWS] = r3 < g.ws_dim ? qv[r3] : 0.f;
    }
}

// Signed distance (and stored gradient) of field f at point p: nearest-texel lookup on the voxel grid (Appendix C.5) or
// the analytic workspace-boundary box (C.4).
__device__ __forceinline__ void field_lookup(const GuideDev& g, int f, const float (&p)[3], float& sdf, float (&gr)[3]) {
    gr[0] = gr[1] = gr[2] = 0.f;
    if (f < g.n_grid) {
        long long flat = 0;
        for (int d = 0; d < g.ws_dim; ++d) {
            float uu = rintf(__fdiv_rn(__fsub_rn(p[d], g.glo[d]), g.cell));
            uu = fminf(fmaxf(uu, 0.f), (float)(g.gshape[d] - 1));
            flat = flat * g.gshape[d] + (long long)uu;
        }
        if (g.ws_dim == 3) {
            float4 t = __ldg(reinterpret_cast<const float4*>(g.tex[f]) + flat);
            sdf = t.x; gr[0] = t.y; gr[1] = t.z; gr[2] = t.w;
        } else {
            const float* t = g.tex[f] + flat * 3;
            sdf = __ldg(t); gr[0] = __ldg(t + 1); gr[1] = __ldg(t + 2);
        }
    } else {
        int arg = 0; float sgn = 1.f, best = 3.4e38f;
        for (int d = 0; d < g.ws_dim; ++d) {
            float lo = __fsub_rn(p[d], g.blo[d]), hi = __fsub_rn(g.bhi[d], p[d]);
            float m = fminf(lo, hi);
            if (m < best) { best = m; arg = d; sgn = (lo <= hi) ? 1.f : -1.f; }
        }
        sdf = best;
        gr[arg] = sgn;
    }
}

__device__ __forceinline__ float clip_scale(float n, float max_norm) {
    // torch.clip(n, 0, max) / n
    return fminf(fmaxf(n, 0.f), max_norm) / n;
}

// KIND: 1 = Panda, 0 = point mass (compile-time copy of g.robot_kind); SPGT: spheres per sphere group actually present
// (ceil(n_spheres / NSG)). Specialising removes the other robot's code and the unrolled bodies of absent sphere slots: the
// kernel runs every instruction once per launch, so its time is largely instruction fetch (ncu: 29 % of stall samples
// "no instruction" before the split), and a smaller kernel is a faster one.
template <int KIND, int SPGT>
__global__ void __launch_bounds__(GUIDE_THREADS) guide_step_kernel(GuideDev g, GuideStepArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, D = g.D, q = g.q_dim, NI = g.n_interp, NTH = GUIDE_THREADS;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n_coll = g.n_grid + (g.has_border ? 1 : 0);

    float* xn = smem;                  // [H][D] normalised input
    float* xu = xn + H * D;            // [H][D] unnormalised
    float* tot = xu + H * D;           // [H][D] sum_c w_c * clipped grad_c
    float* gq = tot + H * D;           // [n_coll][NSG][NI][q] partial d cost_f / d q_interp per sphere group
    float* fgrad = gq + n_coll * NSG * NI * q;      // [n_coll][H][q] clipped, weighted per-field gradients
    float* w1s = fgrad + n_coll * H * q;            // [NI] interpolation weight of the upper tap
    int* i0s = reinterpret_cast<int*>(w1s + NI);    // [NI] lower tap
    float* fk = reinterpret_cast<float*>(i0s + NI); // [(42 + 3*n_spheres)][FK_ROWS] FK scratch of one pass

    long long* dbg = (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) ? a.dbg : nullptr;
    if (dbg) dbg[0] = clock64();
    pdl_launch_dependents();
    pdl_wait();  // x and the clip flag come from the previous kernel
    const int n_it = a.n_iters > 1 ? a.n_iters : 1;
    for (int it = 0; it < n_it; ++it) {
    const bool last_it = it == n_it - 1;
    int flag;
    if (it == 0) flag = a.flag_in ? *a.flag_in : 0;
    else flag = *reinterpret_cast<volatile int*>(a.iter_flags + it);  // complete: every CTA passed the grid barrier below
    const bool pos_only = a.vel_io != nullptr;  // x holds positions only, velocities come from / go to vel_io
    const int Dio = pos_only ? q : D;           // columns of x_in / x_out
    const float* xin = a.x_in + (long long)b * H * Dio;
    for (int i = tid; i < H * D; i += NTH) {
        const int d = i % D, hrow = i / D;
        tot[i] = 0.f;
        if (pos_only && d >= q) {  // velocity half of the state: already unnormalised, not part of x
            xn[i] = 0.f;
            xu[i] = a.vel_io[((long long)b * H + hrow) * q + (d - q)];
            continue;
        }
        float v = it == 0 ? xin[hrow * Dio + d] : xn[i];  // later evaluations continue from the trajectory kept in shared memory
        xn[i] = v;
        float vc = flag ? fminf(fmaxf(v, -1.f), 1.f) : v;
        // ((x + 1) / 2) * (maxs - mins) + mins, reference operation order (normalization.py:165-167)
        xu[i] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(vc, 1.f), 0.5f), g.range[d]), g.mins[d]);
    }
    __syncthreads();
    if (dbg) dbg[1] = clock64();  // trajectory loaded + unnormalised

    // ---------------- collision costs on the interpolated trajectory ----------------
    // Rows are processed in passes of FK_ROWS: (1) FK_ROWS threads interpolate + run the kinematic chain and park joint
    // frames / sphere centres in shared memory; (2) ALL threads share the (row, sphere group) lookup + adjoint items —
    // NSG sphere groups per row — writing partial dq to gq[f][sg][i][.]; the groups are summed in fixed order later.
    if (n_coll > 0) {
        const float ratio = NI > 1 ? (float)(H - 1) / (float)(NI - 1) : 0.f;  // align_corners=True
        for (int ibase = 0; ibase < NI; ibase += FK_ROWS) {
            // (1a) all threads: interpolated joint value of (row, joint) and its sine / cosine -> scratch [2][7][FK_ROWS]
            float* sc_sin = fk + (42 + 3 * g.n_spheres) * FK_ROWS;
            float* sc_cos = sc_sin + 7 * FK_ROWS;
            for (int item = tid; item < FK_ROWS * 7; item += NTH) {
                const int il = item % FK_ROWS, k = item / FK_ROWS;
                const int i = ibase + il;
                if (i >= NI) continue;
                float r = ratio * (float)i;
                int i0 = (int)r;
                if (i0 > H - 1) i0 = H - 1;
                float l1 = fminf(fmaxf(r - (float)i0, 0.f), 1.f);
                float l0 = 1.f - l1;
                int i1 = i0 + (i0 < H - 1 ? 1 : 0);
                if (k == 0) { i0s[i] = i0; w1s[i] = l1; }
                const float qk = k < q ? __fadd_rn(__fmul_rn(l0, xu[i0 * D + k]), __fmul_rn(l1, xu[i1 * D + k])) : 0.f;
                if (KIND == 1) {
                    float sv, cv;
                    sincosf(qk, &sv, &cv);
                    sc_sin[k * FK_ROWS + il] = sv;
                    sc_cos[k * FK_ROWS + il] = cv;
                } else {
                    sc_sin[k * FK_ROWS + il] = qk;  // point mass: the interpolated coordinate itself
                }
            }
            __syncthreads();
            // (1b) three threads per row (one per matrix row): the kinematic chain on the precomputed sines / cosines
            if (KIND == 1) {
                const int il = tid % FK_ROWS, r3 = tid / FK_ROWS;
                if (r3 < 3 && ibase + il < NI) {
                    float sq[7], cq[7];
#pragma unroll
                    for (int k = 0; k < 7; ++k) { sq[k] = sc_sin[k * FK_ROWS + il]; cq[k] = sc_cos[k * FK_ROWS + il]; }
                    fk_chain_row(g, r3, sq, cq, fk + il, fk + il + 42 * FK_ROWS);
                }
            } else if (tid < FK_ROWS && ibase + tid < NI) {
                float* cen = fk + tid + 42 * FK_ROWS;
                {
                    for (int sp = 0; sp < g.n_spheres; ++sp)
                        for (int r3 = 0; r3 < 3; ++r3) cen[(sp * 3 + r3) * FK_ROWS] = r3 < g.ws_dim ? sc_sin[r3 * FK_ROWS + tid] : 0.f;
                }
            }
            __syncthreads();
            if (dbg) dbg[2] = clock64();  // interpolation + kinematic chain

            // (row, sphere group) items: thread -> row = item % FK_ROWS, group = item / FK_ROWS
            for (int item = tid; item < FK_ROWS * NSG; item += NTH) {
                const int il = item % FK_ROWS, sg = item / FK_ROWS;
                const int i = ibase + il;
                if (i >= NI) continue;
                const float* sc = fk + il;
                const float* cen = sc + 42 * FK_ROWS;
                for (int f = 0; f < n_coll; ++f) {
                    float dq[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    // this group's spheres are sg, sg + NSG, ...; all their texel gathers are issued before any is
                    // consumed (latency overlap)
                    float tsdf[SPGT], tg[SPGT][3];
#pragma unroll
                    for (int u = 0; u < SPGT; ++u) {
                        const int s = sg + u * NSG;
                        tsdf[u] = 3.4e38f; tg[u][0] = tg[u][1] = tg[u][2] = 0.f;
                        if (s < g.n_spheres) {
                            const float p[3] = {cen[(s * 3 + 0) * FK_ROWS], cen[(s * 3 + 1) * FK_ROWS], cen[(s * 3 + 2) * FK_ROWS]};
                            field_lookup(g, f, p, tsdf[u], tg[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < SPGT; ++u) {
                        const int s = sg + u * NSG;
                        if (s >= g.n_spheres) continue;
                        const float viol = __fsub_rn(__fadd_rn(g.sphere_r[s], g.margin), tsdf[u]);
                        if (viol > 0.f) {
                            const float gx = -tg[u][0], gy = -tg[u][1], gz = -tg[u][2];  // d cost / d p
                            if (KIND == 1) {
                                const float p0 = cen[(s * 3 + 0) * FK_ROWS], p1 = cen[(s * 3 + 1) * FK_ROWS], p2 = cen[(s * 3 + 2) * FK_ROWS];
                                const int nj = g.sphere_frame[s] < 7 ? g.sphere_frame[s] : 7;
#pragma unroll
                                for (int j = 0; j < 7; ++j) {
                                    if (j >= nj) break;
                                    float rx = p0 - sc[(j * 3 + 0) * FK_ROWS], ry = p1 - sc[(j * 3 + 1) * FK_ROWS],
                                          rz = p2 - sc[(j * 3 + 2) * FK_ROWS];
                                    float zx = sc[(21 + j * 3 + 0) * FK_ROWS], zy = sc[(21 + j * 3 + 1) * FK_ROWS],
                                          zz = sc[(21 + j * 3 + 2) * FK_ROWS];
                                    // (z x r) . g
                                    dq[j] += (zy * rz - zz * ry) * gx + (zz * rx - zx * rz) * gy + (zx * ry - zy * rx) * gz;
                                }
                            } else {
                                dq[0] += gx;
                                if (g.ws_dim > 1) dq[1] += gy;
                                if (g.ws_dim > 2) dq[2] += gz;
                            }
                        }
                    }
                    float* dst = gq + (((long long)f * NSG + sg) * NI + i) * q;
                    for (int k = 0; k < q; ++k) dst[k] = dq[k];
                }
            }
            __syncthreads();
            if (dbg) dbg[3] = clock64();  // field lookups + hinge + J^T
        }

        // adjoint of the interpolation (gather form), then per-cost clip / endpoint zero / weight: one thread per
        // (support row, field); sphere groups are summed in fixed order (deterministic)
        const float inv_ratio = ratio > 0.f ? 1.f / ratio : 0.f;
        // 8 lanes per (support row, field): lane k < q owns coordinate k; the clip norm is an xor-shuffle tree over the 8 lanes
        for (int item0 = 0; item0 < H * n_coll; item0 += NTH / 8) {
            const int item = item0 + (tid >> 3), k = tid & 7;
            const bool on = item < H * n_coll;
            const int h = on ? item % H : 0, f = on ? item / H : 0;
            float gsk = 0.f;
            if (on && k < q) {
                // rows i with ratio * i in (h - 1, h + 1) touch h; floor / ceil leave one row of slack on each side for the
                // rounding of the products (rows that do not touch h contribute with weight 0)
                int lo_i = (int)floorf((float)(h - 1) * inv_ratio);
                int hi_i = (int)ceilf((float)(h + 1) * inv_ratio);
                if (lo_i < 0) lo_i = 0;
                if (hi_i > NI - 1) hi_i = NI - 1;
                // 32-bit indices, no data-dependent branch (rows that do not touch h contribute with weight 0): the four
                // 8-lane groups of a warp stay converged
                const float* gqf = gq + f * NSG * NI * q + k;
                const int sg_stride = NI * q;
                for (int i = lo_i; i <= hi_i; ++i) {
                    const int i0 = i0s[i];
                    const int i1 = i0 + (i0 < H - 1 ? 1 : 0);
                    const float l1 = w1s[i];
                    const float cw = (i0 == h ? 1.f - l1 : 0.f) + (i1 == h ? l1 : 0.f);
                    const float* gp = gqf + i * q;
                    float dqs = 0.f;
#pragma unroll
                    for (int sg = 0; sg < NSG; ++sg) dqs += gp[sg * sg_stride];
                    gsk = cw != 0.f ? fmaf(cw, dqs, gsk) : gsk;
                }
            }
            float scale = 1.f;
            if (g.clip) {
                float t = (on && k < q) ? gsk + 1e-6f : 0.f;
                float n2 = t * t;
                n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
                n2 += __shfl_xor_sync(0xffffffffu, n2, 2);
                n2 += __shfl_xor_sync(0xffffffffu, n2, 4);
                if (!pos_only) n2 += (float)(D - q) * (1e-6f * 1e-6f);  // the zero velocity half of the gradient, + 1e-6 each
                scale = clip_scale(sqrtf(n2), g.max_norm);
            }
            if (on && k < q) {
                const float wgt = f < g.n_grid ? g.w_grid[f] : g.w_border;
                // per-field results are parked in shared memory and added to `tot` in field order below
                fgrad[(f * H + h) * q + k] = (h != 0 && h != H - 1) ? wgt * (scale * gsk) : 0.f;
            }
        }
        __syncthreads();
        if (dbg) dbg[4] = clock64();  // interpolation adjoint + clip
        for (int idx = tid; idx < H * q; idx += NTH) {
            const int h = idx / q, k = idx - h * q;
            float t = tot[h * D + k];
            for (int f = 0; f < n_coll; ++f) t += fgrad[(f * H + h) * q + k];
            tot[h * D + k] = t;
        }
        __syncthreads();
    }

    if (dbg) dbg[5] = clock64();  // fields summed
    // ---------------- GP prior (constant-velocity) on the support points ----------------
    if (g.use_gp) {
        for (int h0 = 0; h0 < H; h0 += NTH / 8) {
            const int h = h0 + (tid >> 3), k = tid & 7;
            const bool on = h > 0 && h < H - 1 && k < q;  // gradient rows 0 and H-1 are zeroed by the guide manager
            float gpk = 0.f, gvk = 0.f, n2 = 0.f, n2v = 0.f;
            if (on) {
                const float pm = xu[(h - 1) * D + k], pc = xu[h * D + k], pn = xu[(h + 1) * D + k];
                const float vm = xu[(h - 1) * D + q + k], vc = xu[h * D + q + k], vn = xu[(h + 1) * D + q + k];
                const float ep0 = pc - pm - g.dt * vm, ev0 = vc - vm;  // e_{h-1}
                const float ep1 = pn - pc - g.dt * vc, ev1 = vn - vc;  // e_h
                const float up0 = 2.f * (g.gp_a * ep0 + g.gp_b * ev0), uv0 = 2.f * (g.gp_b * ep0 + g.gp_c * ev0);
                const float up1 = 2.f * (g.gp_a * ep1 + g.gp_b * ev1), uv1 = 2.f * (g.gp_b * ep1 + g.gp_c * ev1);
                gpk = up0 - up1;
                gvk = uv0 - uv1 - g.dt * up1;
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
            if (on) {
                const float scale = g.clip ? clip_scale(sqrtf(n2), g.max_norm) : 1.f;
                const float scale_v = pos_only ? (g.clip ? clip_scale(sqrtf(n2v), g.max_norm) : 1.f) : scale;
                tot[h * D + k] += g.w_gp * (scale * gpk);
                tot[h * D + q + k] += g.w_gp * (scale_v * gvk);
            }
        }
    }
    __syncthreads();
    if (dbg) dbg[6] = clock64();  // GP stencil

    // ---------------- output ----------------
    float* xout = a.x_out + (long long)b * H * Dio;
    bool viol = false;
    float var = 1.f;
    const bool use_var = a.model_var != nullptr || a.use_var_uniform;
    if (a.model_var != nullptr) var = a.model_var[b];
    else if (a.use_var_uniform) var = a.var_uniform;
    for (int i = tid; i < H * D; i += NTH) {
        float grad = -1.f * tot[i];
        if (pos_only) {  // gradient of the positions out, velocity trajectory updated in place (guides.py:110-112)
            const int d = i % D, hrow = i / D;
            if (d < q) xout[hrow * q + d] = grad;
            else a.vel_io[((long long)b * H + hrow) * q + (d - q)] = __fsub_rn(xu[i], tot[i]);
            continue;
        }
        if (a.grad_only) {
            xout[i] = grad;
            continue;
        }
        if (use_var) grad = __fmul_rn(var, grad);
        float v = __fadd_rn(xn[i], grad);
        const int h = i / D, d = i - h * D;
        int hc = -1;
        for (int k = 0; k < a.n_hc; ++k)
            if (a.hc_rows[k] == h) hc = k;
        if (hc >= 0) v = a.hc_vals[((long long)hc * a.B + b) * D + d];
        viol |= (v > 1.0001f) || (v < -1.0001f);
        if (!last_it) {  // next evaluation of this launch reads it from shared memory
            xn[i] = v;
            continue;
        }
        if (a.noise != nullptr && hc < 0) v = __fadd_rn(v, __fmul_rn(__fmul_rn(a.noise_sd, a.noise[(long long)b * H * D + i]), a.noise_mult));
        xout[i] = v;
        if (a.out2) a.out2[(long long)b * a.out2_bstride + i] = v;
    }
    if (!last_it) {
        // batch-global clip flag of the next evaluation (LimitsNormalizer.unnormalize looks at the whole batch), then a grid
        // barrier: every CTA has contributed before anyone reads it. Bounded spin: a scheduling problem traps, never hangs.
        const int any = __syncthreads_or(viol ? 1 : 0);
        if (tid == 0) {
            if (any) atomicOr(a.iter_flags + it + 1, 1);
            __threadfence();
            atomicAdd(a.iter_counters + it, 1u);
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned int*>(a.iter_counters + it) < gridDim.x) {
                if (clock64() - t0 > 4000000000LL) __trap();
            }
            __threadfence();
        }
        __syncthreads();
    } else if (a.flag_out != nullptr) {
        if (__syncthreads_or(viol ? 1 : 0) && tid == 0) atomicOr(a.flag_out, 1);
    }
    }  // evaluations
    if (dbg) dbg[7] = clock64();  // update written
}

// ---------------------------------------------------------------------------------------------------
// Post-sampling evaluation (reference inference.py:288-326: get_trajs_collision_and_free,
// compute_collision_intensity_trajs, compute_smoothness, compute_path_length): forward-only reuse of the guide's
// interpolation + FK + field lookups. One CTA per (unnormalised) trajectory.
//   stats[b] = { #interpolated waypoints in collision, smoothness = sum_h |v_{h+1} - v_h|, path length = sum_h |p_{h+1} - p_h|,
//                minimum clearance min(sdf - radius) over waypoints, spheres and fields }
// A waypoint is in collision when any sphere has sdf - radius < margin in any field.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FK_ROWS) eval_kernel(GuideDev g, const float* __restrict__ x, float* __restrict__ stats,
                                                       float margin, int B, int H) {
    extern __shared__ __align__(16) float smem[];
    const int D = g.D, q = g.q_dim, NI = g.n_interp;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n_coll = g.n_grid + (g.has_border ? 1 : 0);
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
            for (int f = 0; f < n_coll; ++f)
                for (int s = 0; s < g.n_spheres; ++s) {
                    const float p[3] = {cen[(s * 3 + 0) * FK_ROWS], cen[(s * 3 + 1) * FK_ROWS], cen[(s * 3 + 2) * FK_ROWS]};
                    float sdf, gr[3];
                    field_lookup(g, f, p, sdf, gr);
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

static size_t guide_smem_bytes(const GuideDev& g, int H) {
    const int n_coll = g.n_grid + (g.has_border ? 1 : 0);
    size_t f = (size_t)3 * H * g.D + (size_t)n_coll * NSG * g.n_interp * g.q_dim + (size_t)n_coll * H * g.q_dim +
               2 * (size_t)g.n_interp + (size_t)(42 + 3 * g.n_spheres + 14) * FK_ROWS;  // + sine / cosine scratch [2][7][FK_ROWS]
    return f * sizeof(float);
}

int guide_max_coresident(mpdb_guide* gd, int H) {
    GuideDev g = make_dev(gd->cfg);
    const size_t smem = guide_smem_bytes(g, H);
    int per_sm = 0, sms = 0;
    cudaFuncSetAttribute(guide_step_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, guide_step_kernel<1, 2>, GUIDE_THREADS, smem) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, gd->device) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (per_sm > 0 ? 1 : 0) * sms;  // one CTA per SM counted: the other kernels of the loop may still hold shared memory
}

int guide_launch_step(mpdb_guide* gd, const GuideStepArgs& a, cudaStream_t stream) {
    GuideDev g = make_dev(gd->cfg);
    MPDB_REQUIRE(g.D <= MPDB_MAX_STATE_DIM && g.q_dim <= 7, "guide: state dim too large");
    MPDB_REQUIRE(g.n_interp >= 1, "guide: n_interp must be >= 1");
    const size_t smem = guide_smem_bytes(g, a.H);
    MPDB_REQUIRE(smem <= 220 * 1024, "guide: trajectory does not fit in shared memory");
    MPDB_REQUIRE(a.n_iters <= 1 || (a.iter_flags && a.iter_counters && !a.grad_only && a.B <= guide_max_coresident(gd, a.H)),
                 "guide: several evaluations per launch need flag / counter scratch and a co-resident grid");
    MPDB_REQUIRE(a.vel_io == nullptr || (a.grad_only && a.n_iters <= 1 && g.use_gp >= 0), "guide: position-only mode returns the gradient only");
    const int spg = (g.n_spheres + NSG - 1) / NSG;
    MPDB_REQUIRE(spg >= 1 && spg <= SPG, "guide: bad sphere count");
#define MPDB_GUIDE_LAUNCH(K, S)                                                                                              \
    {                                                                                                                        \
        static bool configured = false;                                                                                      \
        if (!configured) {                                                                                                   \
            MPDB_CHECK_CUDA(cudaFuncSetAttribute(guide_step_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
            configured = true;                                                                                               \
        }                                                                                                                    \
        MPDB_CHECK_CUDA(launch_kernel(guide_step_kernel<K, S>, dim3(a.B), dim3(GUIDE_THREADS), smem, stream, g, a));         \
    }
    if (g.robot_kind == 1) {
        if (spg <= 2) MPDB_GUIDE_LAUNCH(1, 2) else MPDB_GUIDE_LAUNCH(1, SPG)
    } else {
        if (spg <= 1) MPDB_GUIDE_LAUNCH(0, 1) else MPDB_GUIDE_LAUNCH(0, SPG)
    }
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
    for (int s = 0; s < cfg->n_spheres && cfg->robot_kind == 1; ++s)
        MPDB_REQUIRE(cfg->sphere_frame[s] >= 1 && cfg->sphere_frame[s] <= 8, "guide: sphere frame out of range");
    MPDB_CHECK_CUDA(cudaSetDevice(device));
    mpdb_guide* g = new mpdb_guide();
    g->cfg = *cfg;
    g->device = device;
    g->flags = nullptr;
    if (cudaMalloc(&g->flags, 16 * sizeof(int)) != cudaSuccess) {
        delete g;
        mpdb::set_error("mpdb_guide_create: cudaMalloc failed");
        return 1;
    }
    *out = g;
    return 0;
}

extern "C" void mpdb_guide_destroy(mpdb_guide* g) {
    if (!g) return;
    cudaFree(g->flags);
    delete g;
}

extern "C" int mpdb_guide_grad(mpdb_guide* g, const float* x, float* grad, int32_t B, int32_t H, void* stream) {
    MPDB_REQUIRE(g && x && grad && B > 0 && H > 1, "mpdb_guide_grad: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_CHECK_CUDA(cudaSetDevice(g->device));
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
    MPDB_REQUIRE(g && x_pos && velocity && grad && B > 0 && H > 1, "mpdb_guide_grad_pos: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_CHECK_CUDA(cudaSetDevice(g->device));
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
    MPDB_REQUIRE(g && x && B > 0 && H > 1 && n_steps >= 0, "mpdb_guide_steps: bad argument");
    MPDB_REQUIRE(n_hc >= 0 && n_hc <= MPDB_MAX_HARD_CONDS, "mpdb_guide_steps: too many hard conditions");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_CHECK_CUDA(cudaSetDevice(g->device));
    if (n_steps == 0) return 0;
    if (guide_launch_flag(x, (long long)B * H * 2 * g->cfg.q_dim, g->flags, st)) return 1;
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
    MPDB_CHECK_CUDA(cudaSetDevice(g->device));
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
        const char* names[8] = {"start", "loaded", "fk", "lookups", "adjoint+clip", "fields summed", "gp", "written"};
        for (int k = 1; k < 8; ++k) fprintf(stderr, "[guide timeline] %-14s +%.2f us (t = %.2f)\n", names[k], (h[k] - h[k - 1]) / 1965.0, (h[k] - h[0]) / 1965.0);
    }
    return 0;
}

extern "C" int mpdb_eval_trajectories(mpdb_guide* gd, const float* x_unnormalized, float* stats, float margin, int32_t B,
                                      int32_t H, void* stream) {
    MPDB_REQUIRE(gd && x_unnormalized && stats && B > 0 && H > 1, "mpdb_eval_trajectories: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPDB_CHECK_CUDA(cudaSetDevice(gd->device));
    GuideDev g = make_dev(gd->cfg);
    const size_t smem = sizeof(float) * ((size_t)H * g.D + FK_ROWS + (size_t)(42 + 3 * g.n_spheres) * FK_ROWS);
    MPDB_REQUIRE(smem <= 200 * 1024, "mpdb_eval_trajectories: trajectory does not fit in shared memory");
    static bool configured = false;
    if (!configured) {
        MPDB_CHECK_CUDA(cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    eval_kernel<<<B, FK_ROWS, smem, st>>>(g, x_unnormalized, stats, margin, B, H);
    MPDB_LAUNCH_CHECK();
    return 0;
}
