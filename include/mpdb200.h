/*
 * mpdb200 — C ABI of the B200-native guided-diffusion trajectory sampler.
 *
 * The reference (jacarvalho/mpd-public) is pure Python/PyTorch and has no FFI layer; its boundary for
 * this path is the Python surface that scripts/inference/inference.py touches. The
 * entry points below are what a ctypes binding inside that Python surface would call — each one
 * cites the reference interface it replaces. Plain pointers and sizes only: device pointers are raw
 * CUDA addresses (`tensor.data_ptr()`), `stream` is a `cudaStream_t` cast to `void*`
 * (`torch.cuda.current_stream().cuda_stream`), all tensors are contiguous fp32 unless noted.
 *
 * Every function returns 0 on success or a non-zero status; `mpdb_last_error()` returns the message
 * of the last failure on the calling thread (the Python binding raises RuntimeError with it —
 * the reference signals errors with Python exceptions, diffusion_model_base.py:72,275).
 *
 * Concurrency: an engine or guide handle owns device scratch (activation workspace, clip flags, captured graphs) and serves ONE
 * stream at a time, like the reference's single Python thread on the current stream. Calls on different handles
 * may run concurrently; calls on the same handle must be ordered by the caller (same stream, or events between streams).
 * Every entry point saves and restores the calling thread's current device.
 */
#ifndef MPDB200_H
#define MPDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPDB_MAX_LEVELS 8
#define MPDB_MAX_STATE_DIM 32
#define MPDB_MAX_SPHERES 16
#define MPDB_MAX_GRID_FIELDS 4
#define MPDB_MAX_HARD_CONDS 8

typedef struct mpdb_engine mpdb_engine; /* TemporalUnet + GaussianDiffusionModel state on one device */
typedef struct mpdb_guide mpdb_guide;   /* GuideManagerTrajectoriesWithVelocity + CostComposite state */

/* TemporalUnet(n_support_points, state_dim, unet_input_dim, dim_mults) — temporal_unet.py:22-35;
 * GaussianDiffusionModel(n_diffusion_steps, predict_epsilon, clip_denoised) — diffusion_model_base.py:48-56 */
typedef struct {
    int32_t state_dim;
    int32_t horizon;
    int32_t unet_input_dim;
    int32_t n_levels;
    int32_t dim_mults[MPDB_MAX_LEVELS];
    int32_t n_diffusion_steps;
    int32_t predict_epsilon;
    int32_t clip_denoised;
    int32_t max_batch; /* initial workspace size; grows on demand */
} mpdb_engine_config;

/* What CostComposite([CostCollision(field)..., CostGPTrajectory]) and the guide manager hold
 * (inference.py:195-236, guides.py:149-171). Arithmetic of the absent cost/robot/field classes follows
 * SURVEY.md Appendix C/E. One CostCollision per field of task.get_collision_fields() (inference.py:193-204): grid-backed
 * object fields (each with its own lattice), the workspace-boundary box, the robot's self-collision field; every one
 * carries its own cutoff margin, sigma_coll and gradient weight. */
typedef struct {
    int32_t robot_kind; /* 0 = point mass (FK identity), 1 = Panda 7-DoF chain */
    int32_t q_dim;
    int32_t ws_dim;
    int32_t n_spheres;
    int32_t sphere_frame[MPDB_MAX_SPHERES]; /* Panda: 1..8 (link 1..7, flange) */
    float sphere_offset[MPDB_MAX_SPHERES][3];
    float sphere_radius[MPDB_MAX_SPHERES];
    float mins[MPDB_MAX_STATE_DIM]; /* LimitsNormalizer limits, normalization.py:149-167 */
    float maxs[MPDB_MAX_STATE_DIM];
    int32_t n_grid_fields;
    const float* grid_texels[MPDB_MAX_GRID_FIELDS]; /* device; [cells][1+ws_dim] = {sdf, grad} */
    int32_t grid_shape[MPDB_MAX_GRID_FIELDS][3];
    float grid_lo[MPDB_MAX_GRID_FIELDS][3];
    float grid_cell[MPDB_MAX_GRID_FIELDS];
    float margin_grid[MPDB_MAX_GRID_FIELDS];     /* cutoff margin of the field's CostCollision */
    float sigma_grid[MPDB_MAX_GRID_FIELDS];      /* sigma_coll: cost = sum hinge / sigma^2 */
    float weight_grid[MPDB_MAX_GRID_FIELDS];
    int32_t has_border; /* workspace-boundary field */
    float border_lo[3];
    float border_hi[3];
    float margin_border, sigma_border, weight_border;
    /* robot self-collision field: cost = sum over interpolated rows and listed sphere pairs (a, b) of
     * relu(margin - (|c_a - c_b| - r_a - r_b)) / sigma^2; bit b of self_pairs[a] lists the pair (a, b), symmetric */
    int32_t has_self;
    uint32_t self_pairs[MPDB_MAX_SPHERES];
    float margin_self, sigma_self, weight_self;
    float dt;
    float sigma_gp;
    float weight_gp;
    int32_t use_gp;
    int32_t clip_grad;
    float max_grad_norm;
    int32_t n_interp; /* num_interpolated_points_for_collision (guides.py:153); == horizon when off */
    int32_t vel_from_fd; /* position-only manager with use_velocity_from_finite_difference (guides.py:77-79): the velocity
                          * half of the state is the central difference of the positions (zero at both ends) */
} mpdb_guide_config;

/* p_sample_loop(..., sample_fn=ddpm_sample_fn, n_diffusion_steps_without_noise, **sample_kwargs)
 * — diffusion_model_base.py:158-182, sample_functions.py:18-62 */
typedef struct {
    int32_t n_steps_without_noise;
    int32_t t_start_guide; /* INT32_MAX for +inf */
    int32_t n_guide_steps;
    int32_t scale_grad_by_std;
    const float* noise_std; /* host, one value per loop iteration (noise_std_extra_schedule_fn(t)) */
    int32_t n_hard_conds;
    int32_t hard_cond_rows[MPDB_MAX_HARD_CONDS];
    const float* hard_cond_vals; /* device [n_hard_conds][B][D] */
    int32_t use_cuda_graph;
    int32_t horizon;   /* H and D of the noise / x_out / chain tensors: must equal the engine's (checked; a mismatch would */
    int32_t state_dim; /* read and write past the caller's buffers) */
} mpdb_loop_params;

/* ddim_sample(shape, hard_conds, t_start_guide, guide, **sample_kwargs) — diffusion_model_base.py:184-259 (eta = 0).
 * The host computes the time pairs and the two per-step coefficients with the reference's own torch expressions
 * (:203-209, :232-236) so that they are bit-identical; the device runs, per pair, UNet -> x_start / pred_noise ->
 * x_start * sqrt(alpha_next) + c * pred_noise [-> guide_gradient_steps when time_next < t_start_guide] -> hard conditions. */
typedef struct {
    int32_t n_steps;              /* number of (time, time_next) pairs */
    const int32_t* times;         /* host [n_steps] */
    const int32_t* times_next;    /* host [n_steps]; -1 ends the loop with x = x_start (:223-230) */
    const float* sqrt_alpha_next; /* host [n_steps]: alpha_next.sqrt() */
    const float* coef_noise;      /* host [n_steps]: (1 - alpha_next - sigma ** 2).sqrt() */
    int32_t t_start_guide;        /* INT32_MAX for +inf */
    int32_t n_guide_steps;        /* what guide_gradient_steps receives through **sample_kwargs (its default is 1) */
    int32_t n_hard_conds;
    int32_t hard_cond_rows[MPDB_MAX_HARD_CONDS];
    const float* hard_cond_vals;  /* device [n_hard_conds][B][D] */
    int32_t horizon;
    int32_t state_dim;
} mpdb_ddim_params;

const char* mpdb_last_error(void);
int mpdb_version(void);

/* ---- engine: nn.Module construction / load_state_dict (inference.py:138-149) ---- */
int mpdb_engine_create(const mpdb_engine_config* cfg, int device, mpdb_engine** out);
void mpdb_engine_destroy(mpdb_engine* e);
/* one TemporalUnet state-dict entry (reference layout), copied from device memory */
int mpdb_engine_set_param(mpdb_engine* e, const char* name, const float* dev_ptr, int64_t numel, void* stream);
/* schedule buffers of GaussianDiffusionModel (diffusion_model_base.py:82-104), host arrays of [T] */
int mpdb_engine_set_schedule(mpdb_engine* e, const float* sqrt_recip_alphas_cumprod,
                             const float* sqrt_recipm1_alphas_cumprod, const float* posterior_mean_coef1,
                             const float* posterior_mean_coef2, const float* posterior_log_variance_clipped,
                             const float* posterior_std, const float* posterior_var);
/* options: "tc_mode" = 0 exact fp32 FMA path only | 1 auto (default: tcgen05 path, 22-bit scaled-fp16 operand split, at
 * every loop step whose sqrt(1/abar_t - 1) <= tc_amp_limit, exact path otherwise; the per-call entry points use the tensor
 * cores when no finite limit is set) | 2
 * force tensor cores everywhere; "tc_amp_limit" (default: no limit — the split matches the fp32 path even at t = T-1); "mega" = 1 (default: the UNet runs as ONE launch of the whole-forward cluster
 * kernel, unet_mega.cu, whenever the batch fits one wave of 8-CTA clusters) | 0 per-layer kernels | 2 cluster kernel for
 * any batch; "fuse_final" = 1 (default: final_conv.1 + DDPM update in the cluster kernel's last epilogue inside the loop);
 * "fuse_guide" = 1 (default: the n_guide_steps evaluations of a loop step in one launch when the batch is co-resident — the
 * trajectory stays in shared memory, the batch-global clip flag is resolved per CTA, bit-identical to one launch per
 * evaluation; 0: one launch per evaluation); "prec1_amp_limit" (default 0.21; 0 disables): loop steps whose eps-to-mean
 * amplification posterior_mean_coef1[t] * sqrt(1/abar_t - 1) is at most this issue one fp16 product per MMA step instead of
 * the three of the 22-bit split ("tc_mode" = 2 always uses the split);
 * "fuse_rtb" = 1 (default: the per-layer path runs a residual block with C_out <= 128 as one cluster-fused launch while its
 * CTAs fit one wave; larger batches run every layer as a persistent kernel with double-buffered accumulators) | 2 always | 0 never;
 * "alias_buffers" = 1 (default) shares activation storage between layers with disjoint lifetimes, 0 keeps one buffer
 * per layer (needed by mpdb_engine_read_buffer; disables the cluster kernel); "timeline" / "mega_timeline" = 1 enable the
 * clock64 stamps read by mpdb_engine_read_timeline / mpdb_engine_read_mega_timeline */
int mpdb_engine_set_option(mpdb_engine* e, const char* name, double value);
/* repack weights, precompute the time-conditioning tables; errors if a parameter is missing */
int mpdb_engine_finalize(mpdb_engine* e, void* stream);

/* TemporalUnet.forward(x, time, context=None) — temporal_unet.py:118-171. x,eps: [B,H,D]; t: int64 [B] */
int mpdb_unet_forward(mpdb_engine* e, const float* x, const int64_t* t, float* eps, int32_t B, void* stream);
/* the same forward at one uniform timestep, exactly as mpdb_sample_loop runs it (tensor-core policy of the loop; the
 * whole-forward cluster kernel of unet_mega.cu when option "mega" = 1 (default) and the configuration supports it).
 * make_timesteps fills t with one value (diffusion_model_base.py:25-27), so this is the loop's only forward. */
int mpdb_unet_forward_uniform(mpdb_engine* e, const float* x, int32_t t, float* eps, int32_t B, void* stream);
/* 1 if batch B runs the UNet as one cluster-kernel launch; G = trajectories per cluster; why = reason when not */
int mpdb_engine_mega_info(mpdb_engine* e, int32_t B, int32_t* G, int32_t* n_layers, int32_t* a_bytes, int32_t* smem_bytes,
                          char* why, int why_cap);
/* GaussianDiffusionModel.p_mean_variance -> model_mean — diffusion_model_base.py:143-155 */
int mpdb_p_mean(mpdb_engine* e, const float* x, const int64_t* t, float* mean, int32_t B, void* stream);
/* x + model_std * noise * noise_std with noise[t == 0] = 0 — sample_functions.py:50-62 (in place on x) */
int mpdb_add_noise(mpdb_engine* e, float* x, const int64_t* t, const float* noise, float noise_std, int32_t B,
                   void* stream);
/* the whole reverse loop, fused. noise: [n_iters+1][B][H][D] (row 0 = initial x). chain_out may be NULL;
 * chain strides are in floats (so [B,S,H,D] and [S,B,H,D] are both expressible). */
int mpdb_sample_loop(mpdb_engine* e, mpdb_guide* g, const mpdb_loop_params* p, const float* noise, float* x_out,
                     float* chain_out, int64_t chain_step_stride, int64_t chain_batch_stride, int32_t B, void* stream);
/* the DDIM loop, fused: x_init [B][H][D] (= randn with hard conditions not yet applied), x_out [B][H][D]; chain_out (may be
 * NULL) receives n_steps + 1 entries (x_T with hard conditions, then every step), strides in floats as above. */
int mpdb_ddim_loop(mpdb_engine* e, mpdb_guide* g, const mpdb_ddim_params* p, const float* x_init, float* x_out,
                   float* chain_out, int64_t chain_step_stride, int64_t chain_batch_stride, int32_t B, void* stream);
/* The loop's noise — x = torch.randn(shape) (diffusion_model_base.py:165) and noise = torch.randn_like(x) per step
 * (sample_functions.py:51) — drawn in ONE launch, bit-identical to n_draws consecutive `normal_()` calls of torch's CUDA
 * generator on contiguous fp32 tensors of `numel` elements: out device [n_draws][numel]; state_dev device int64[2] = {seed,
 * Philox offset of the first draw} (read when the kernel runs, so the launch can sit in a CUDA graph). The caller advances
 * the torch generator by n_draws * mpdb_normal_offset_increment(numel, device). */
int mpdb_normal_fill(float* out, int64_t numel, int32_t n_draws, const int64_t* state_dev, int device, void* stream);
int64_t mpdb_normal_offset_increment(int64_t numel, int device);
/* LimitsNormalizer.normalize (mpd/datasets/normalization.py:150-155) in one launch: rows x[n_rows][d_in] (device, fp32,
 * contiguous), zero-extended to d_out >= d_in columns — `get_hard_conditions` normalises cat(position, zeros),
 * mpd/datasets/trajectories.py:214-237 —, out[n_rows][d_out] = 2 * ((v - mins[d]) / range[d]) - 1 with range = maxs - mins; the
 * reference's operation order, every operation rounded on its own as in the eager torch ops (bit-identical, tested). */
int mpdb_limits_normalize(const float* x, int64_t n_rows, int32_t d_in, const float* mins, const float* range, float* out,
                          int32_t d_out, int device, void* stream);
/* counter bumped whenever the engine reallocates device buffers, reloads parameters or changes an option: a caller that
 * captured mpdb_sample_loop (use_cuda_graph = 0) into its own CUDA graph must re-capture when it changes */
int64_t mpdb_engine_generation(mpdb_engine* e);
/* kernels launched by this engine/guide pair since creation (bench.py's gpu_launches) */
int64_t mpdb_launch_count(void);

/* measurement (bench.py's roofline object): per-layer device time of one UNet forward at uniform t, CUDA
 * events on `stream`; arrays hold mpdb_engine_num_ops(e) entries (mode 0..3 = conv5/conv1/down/up, 4 = fused
 * final projection + posterior mean). */
int mpdb_engine_num_ops(mpdb_engine* e);
/* device time (ms) of the UNet body per forward as the loop runs it, its algorithmic FLOPs and kernels per forward */
int mpdb_profile_unet_body(mpdb_engine* e, const float* x, int32_t t, int32_t B, int32_t reps, float* ms_out,
                           double* flops_out, int32_t* launches_out, void* stream);
int mpdb_profile_forward(mpdb_engine* e, const float* x, int32_t t, int32_t B, int32_t reps, float* ms_out,
                         double* flops_out, int32_t* mode_out, void* stream);
/* average device time of one guide evaluation on x (in place), CUDA events on `stream` */
int mpdb_profile_guide(mpdb_guide* g, float* x, int32_t B, int32_t H, int32_t reps, float* ms_out, void* stream);

/* average device time of ONE launch that runs n_evals guide evaluations on x in place, as mpdb_sample_loop launches them when
 * the batch is co-resident (option "fuse_guide") */
int mpdb_profile_guide_steps(mpdb_guide* g, float* x, int32_t n_evals, int32_t B, int32_t H, int32_t reps, float* ms_out,
                             void* stream);
/* 1 / 3: MMA products per step the loop issues for a forward at timestep t (option "prec1_amp_limit") */
int mpdb_engine_step_precision(mpdb_engine* e, int32_t t);
/* trajectories the fused guide launch can hold at once (0: never fused) */
int mpdb_guide_max_coresident(mpdb_guide* g, int32_t H);

/* unit-test hook for the tcgen05 implicit-GEMM core: raw fp32 accumulators of a k=5 convolution (no bias).
 * x_cm device [B][CI][L+4] with zero halo, w device [CO][CI][5], raw device [ceil(B/SPT)][CO/32][128][32] with
 * SPT = 132/(L+4); row r of a tile holds sample (b % SPT), position l at r = (b % SPT)*(L+4) + l. */
int mpdb_debug_tc_conv5(const float* x_cm, const float* w, float* raw, int32_t B, int32_t CI, int32_t CO, int32_t L,
                        void* stream);

/* debugging: with option "timeline" = 1, clock64 stamps (16 per layer) of CTA (0,0) of every tensor-core conv of the
 * last forward; host_out holds 16 * max_ops int64 */
int mpdb_engine_read_timeline(mpdb_engine* e, int64_t* host_out, int32_t max_ops);

/* debugging: with option "mega_timeline" = 1, clock64 stamps of the whole-forward cluster kernel: per layer and CTA rank
 * of cluster 0 {inputs landed, accumulators complete, epilogue arithmetic done, outputs delivered}; desc_out gets
 * {type, L, C_out, active CTAs} per layer. Returns the number of layers written (negative on error is not used: 1/2 =
 * error codes are returned only when nothing was written). */
int mpdb_engine_read_mega_timeline(mpdb_engine* e, int64_t* host_out, int32_t* desc_out, int32_t max_layers);

/* debugging / parity: intermediate activations of the last mpdb_unet_forward */
int mpdb_engine_num_buffers(mpdb_engine* e);
int mpdb_engine_buffer_info(mpdb_engine* e, int idx, char* name, int name_cap, int32_t* channels, int32_t* length);
int mpdb_engine_read_buffer(mpdb_engine* e, int idx, float* dev_out /* [B][C][L] */, int32_t B, void* stream);

/* ---- guide: GuideManagerTrajectoriesWithVelocity (guides.py:149-236) ---- */
int mpdb_guide_create(const mpdb_guide_config* cfg, int device, mpdb_guide** out);
void mpdb_guide_destroy(mpdb_guide* g);
/* guide(x_normalized) -> grad, [B,H,D] — guides.py:173-211 */
int mpdb_guide_grad(mpdb_guide* g, const float* x, float* grad, int32_t B, int32_t H, void* stream);
/* GuideManagerTrajectories.forward (position-only state, guides.py:60-118): x_pos [B,H,q] normalised positions,
 * velocity [B,H,q] the manager's unnormalised velocity trajectory — read as the velocity half of the state and updated in
 * place (velocity -= sum_c w_c * clip(d cost_c / d velocity)); grad [B,H,q] = -sum_c w_c * zero_ends(clip(d cost_c / d pos)).
 * Position and velocity gradients of a cost are clipped separately. The guide config's mins/maxs cover the q positions.
 * A guide configured with vel_from_fd (use_velocity_from_finite_difference, guides.py:77-79) takes velocity = NULL: the
 * velocity half of the state is the central difference of the positions and only the position gradient exists. */
int mpdb_guide_grad_pos(mpdb_guide* g, const float* x_pos, float* velocity, float* grad, int32_t B, int32_t H, void* stream);
/* guide_gradient_steps(x, hard_conds, guide, n_guide_steps, scale_grad_by_std, model_var) in place —
 * sample_functions.py:65-83. model_var: device [B] or NULL. hard-cond arrays as in mpdb_loop_params. */
int mpdb_guide_steps(mpdb_guide* g, float* x, int32_t n_steps, const float* model_var, int32_t n_hard_conds,
                     const int32_t* hard_cond_rows, const float* hard_cond_vals, int32_t B, int32_t H, void* stream);
/* the diffusion_prior_then_guide post-loop (inference.py:263-282): n_steps x { x <- x + guide(x); hard conditioning }, every
 * iterate also written to chain_out [n_steps][B][H][D] (step stride in floats; may be NULL). x is updated in place. */
int mpdb_guide_steps_chain(mpdb_guide* g, float* x, int32_t n_steps, const float* model_var, int32_t n_hard_conds,
                           const int32_t* hard_cond_rows, const float* hard_cond_vals, float* chain_out,
                           int64_t chain_step_stride, int32_t B, int32_t H, void* stream);
/* Parity instrumentation: every guide evaluation launched for `g` after this call records the discrete decisions of its
 * collision costs into dev_buf (int32 [capacity_evals][B][n_costs][n_interp][n_spheres], cost order = grid fields, border,
 * self): grid field -> (flat texel index << 1) | hinge active; border -> (axis << 2) | (low wall ? 2 : 0) | hinge active;
 * self -> bit mask of the partner spheres whose hinge is active. Evaluations take consecutive slots from 0; the call
 * fails when the capacity is exceeded. dev_buf = NULL switches recording off. Loops run without CUDA graphs while recording. */
int mpdb_guide_record_decisions(mpdb_guide* g, int32_t* dev_buf, int64_t capacity_evals, int32_t B);
/* evaluations recorded since the last mpdb_guide_record_decisions call */
int64_t mpdb_guide_decisions_recorded(mpdb_guide* g);
/* How many (trajectory, evaluation) pairs had their LimitsNormalizer clamp (normalization.py:160-162, a batch-global
 * branch) decided by OTHER trajectories of the batch since the guide was created / last reset: the batch's flag was set while
 * the trajectory itself had elements in (1, 1 + 1e-4] and none beyond. 0 means every result so far is independent of how
 * the batch was composed or sharded (what the multi-GPU equivalence tests check). Synchronises the device. */
int64_t mpdb_guide_batch_dependent_clamps(mpdb_guide* g, int32_t reset);
/* collision costs of this guide (grid fields + border + self) */
int mpdb_guide_num_collision_costs(mpdb_guide* g);
/* forward kinematics of the guide's robot: q device [N][q_dim] -> sphere centres device [N][n_spheres][3] (tests: known
 * answers of the Panda chain) */
int mpdb_debug_fk(mpdb_guide* g, const float* q, float* centers, int32_t N, void* stream);
/* Post-sampling evaluation (inference.py:288-326): per UNNORMALISED trajectory x [B,H,D] ->
 * stats device [B][4] = {#interpolated waypoints in collision (sdf - radius < margin in any field), smoothness
 * sum_h |v_{h+1} - v_h|, path length sum_h |p_{h+1} - p_h|, minimum clearance}. Uses the guide's robot / fields. */
int mpdb_eval_trajectories(mpdb_guide* g, const float* x_unnormalized, float* stats, float margin, int32_t B, int32_t H,
                           void* stream);
/* samples analytic primitives (spheres [ns][dim+1], boxes [nb][2*dim], host arrays) onto the voxel grid:
 * texels_out device [prod(shape)][1+dim] */
int mpdb_sdf_grid_build(int32_t dim, const int32_t* shape, const float* lo, float cell, const float* spheres,
                        int32_t n_spheres, const float* boxes, int32_t n_boxes, float* texels_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPDB200_H */
