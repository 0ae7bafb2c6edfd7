"""The body of the reference's `scripts/inference/inference.py` (lines 76-282: seed -> dataset -> model -> start / goal ->
costs -> guide -> `run_inference` -> optional prior-then-guide post-loop -> metrics), written against the reference's OWN
module paths. `mpd_public_b200.compat.install()` makes those paths resolve to this framework; nothing else differs from how
a user of the reference writes it. Checkpoints are downloads the sandbox does not have, so `load_state_dict` takes seeded
weights in the reference's state-dict layout instead of `ema_model_current_state_dict.pth` (pass --model-dir to load a real
run with `mpd_public_b200.ingest`).

    python examples/inference_like_reference.py [--model-id EnvSpheres3D-RobotPanda] [--planner-alg mpd] [--n-samples 50]
"""
import argparse
import os
import sys
from math import ceil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import einops
import torch

import mpd_public_b200.compat as compat

compat.install()

# ---- the reference's import lines (inference.py:13-26), unchanged ----
from experiment_launcher import single_experiment_yaml, run_experiment  # noqa: E402
from mp_baselines.planners.costs.cost_functions import CostCollision, CostComposite, CostGPTrajectory  # noqa: E402
from mpd.models import TemporalUnet, UNET_DIM_MULTS  # noqa: E402
from mpd.models.diffusion_models.guides import GuideManagerTrajectoriesWithVelocity  # noqa: E402
from mpd.models.diffusion_models.sample_functions import guide_gradient_steps, ddpm_sample_fn  # noqa: E402
from mpd.trainer import get_dataset, get_model  # noqa: E402
from torch_robotics.torch_utils.seed import fix_random_seed  # noqa: E402
from torch_robotics.torch_utils.torch_timer import TimerCUDA  # noqa: E402
from torch_robotics.torch_utils.torch_utils import get_torch_device, freeze_torch_model_params  # noqa: E402
from torch_robotics.trajectory.metrics import compute_smoothness, compute_path_length, compute_variance_waypoints  # noqa: E402


@single_experiment_yaml
def experiment(model_id='EnvSpheres3D-RobotPanda', planner_alg='mpd', use_guide_on_extra_objects_only=False, n_samples=50,
               start_guide_steps_fraction=0.25, n_guide_steps=5, n_diffusion_steps_without_noise=5,
               weight_grad_cost_collision=1e-2, weight_grad_cost_smoothness=1e-7,
               factor_num_interpolated_points_for_collision=1.5, trajectory_duration=5.0, device='cuda', seed=30,
               model_dir=None, compile_model=True, verbose=True, **kwargs):
    fix_random_seed(seed)
    device = get_torch_device(device)
    tensor_args = {'device': device, 'dtype': torch.float32}
    run_prior_only = planner_alg == 'diffusion_prior'
    run_prior_then_guidance = planner_alg == 'diffusion_prior_then_guide'
    if planner_alg not in ('mpd', 'diffusion_prior', 'diffusion_prior_then_guide'):
        raise NotImplementedError

    # training arguments: args.yaml of the run (inference.py:103); the shipped runs use these values (train.py:19-44)
    args = dict(dataset_subdir=model_id, include_velocity=True, variance_schedule='exponential', n_diffusion_steps=25,
                predict_epsilon=True, unet_input_dim=32, unet_dim_mults_option=1, diffusion_model_class='GaussianDiffusionModel',
                use_ema=True)
    if model_dir is not None:
        from mpd.utils.loading import load_params_from_yaml
        args = load_params_from_yaml(os.path.join(model_dir, "args.yaml"))

    # dataset with env, robot, task (inference.py:107-123)
    train_subset, train_dataloader, val_subset, val_dataloader = get_dataset(
        dataset_class='TrajectoryDataset', use_extra_objects=True, obstacle_cutoff_margin=0.05,
        **{k: v for k, v in args.items() if k in ('dataset_subdir', 'include_velocity')}, tensor_args=tensor_args)
    dataset = train_subset.dataset
    n_support_points = dataset.n_support_points
    robot, task = dataset.robot, dataset.task
    dt = trajectory_duration / n_support_points
    robot.dt = dt

    # prior model (inference.py:127-154)
    diffusion_configs = dict(variance_schedule=args['variance_schedule'], n_diffusion_steps=args['n_diffusion_steps'],
                             predict_epsilon=args['predict_epsilon'])
    unet_configs = dict(state_dim=dataset.state_dim, n_support_points=dataset.n_support_points,
                        unet_input_dim=args['unet_input_dim'], dim_mults=UNET_DIM_MULTS[args['unet_dim_mults_option']])
    diffusion_model = get_model(model_class=args['diffusion_model_class'], model=TemporalUnet(**unet_configs),
                                tensor_args=tensor_args, **diffusion_configs, **unet_configs)
    if model_dir is not None:
        name = 'ema_model_current_state_dict.pth' if args['use_ema'] else 'model_current_state_dict.pth'
        diffusion_model.load_state_dict(torch.load(os.path.join(model_dir, 'checkpoints', name), map_location=tensor_args['device']))
    else:
        from mpd_public_b200 import synthetic as S
        sd = S.make_unet_state_dict(0, dataset.state_dim, args['unet_input_dim'], UNET_DIM_MULTS[args['unet_dim_mults_option']])
        diffusion_model.load_state_dict({'model.' + k: torch.as_tensor(v) for k, v in sd.items()}, strict=False)
    diffusion_model.eval()
    model = diffusion_model
    freeze_torch_model_params(model)
    if compile_model:
        model = torch.compile(model)
    model.warmup(horizon=n_support_points, device=device)

    # random initial and final positions (inference.py:156-175)
    start_state_pos, goal_state_pos = None, None
    for _ in range(100):
        q_free = task.random_coll_free_q(n_samples=2)
        start_state_pos, goal_state_pos = q_free[0], q_free[1]
        if torch.linalg.norm(start_state_pos - goal_state_pos) > dataset.threshold_start_goal_pos:
            break
    if start_state_pos is None or goal_state_pos is None:
        raise ValueError("No collision free configuration was found")

    # hard conditions, costs, guide (inference.py:181-245)
    hard_conds = dataset.get_hard_conditions(torch.vstack((start_state_pos, goal_state_pos)), normalize=True)
    context = None
    cost_collision_l, weights_grad_cost_l = [], []
    collision_fields = task.get_collision_fields_extra_objects() if use_guide_on_extra_objects_only else task.get_collision_fields()
    for collision_field in collision_fields:
        cost_collision_l.append(CostCollision(robot, n_support_points, field=collision_field, sigma_coll=1.0, tensor_args=tensor_args))
        weights_grad_cost_l.append(weight_grad_cost_collision)
    cost_smoothness_l = [CostGPTrajectory(robot, n_support_points, dt, sigma_gp=1.0, tensor_args=tensor_args)]
    weights_grad_cost_l.append(weight_grad_cost_smoothness)
    cost_func_list = [*cost_collision_l, *cost_smoothness_l]
    cost_composite = CostComposite(robot, n_support_points, cost_func_list, weights_cost_l=weights_grad_cost_l, tensor_args=tensor_args)
    guide = GuideManagerTrajectoriesWithVelocity(
        dataset, cost_composite, clip_grad=True, interpolate_trajectories_for_collision=True,
        num_interpolated_points=ceil(n_support_points * factor_num_interpolated_points_for_collision), tensor_args=tensor_args)
    t_start_guide = ceil(start_guide_steps_fraction * model.n_diffusion_steps)
    sample_fn_kwargs = dict(guide=None if run_prior_then_guidance or run_prior_only else guide, n_guide_steps=n_guide_steps,
                            t_start_guide=t_start_guide, noise_std_extra_schedule_fn=lambda x: 0.5)

    # sample (inference.py:248-258)
    with TimerCUDA() as timer_model_sampling:
        trajs_normalized_iters = model.run_inference(
            context, hard_conds, n_samples=n_samples, horizon=n_support_points, return_chain=True, sample_fn=ddpm_sample_fn,
            **sample_fn_kwargs, n_diffusion_steps_without_noise=n_diffusion_steps_without_noise)
    t_total = timer_model_sampling.elapsed

    # extra guiding steps without diffusion (inference.py:263-282)
    if run_prior_then_guidance:
        n_post_diffusion_guide_steps = (t_start_guide + n_diffusion_steps_without_noise) * n_guide_steps
        with TimerCUDA() as timer_post_model_sample_guide:
            trajs = trajs_normalized_iters[-1]
            trajs_post_diff_l = []
            for i in range(n_post_diffusion_guide_steps):
                trajs = guide_gradient_steps(trajs, hard_conds=hard_conds, guide=guide, n_guide_steps=1, unnormalize_data=False)
                trajs_post_diff_l.append(trajs)
            chain = torch.stack(trajs_post_diff_l, dim=1)
            chain = einops.rearrange(chain, 'b post_diff_guide_steps h d -> post_diff_guide_steps b h d')
            trajs_normalized_iters = torch.cat((trajs_normalized_iters, chain))
        t_total = timer_model_sampling.elapsed + timer_post_model_sample_guide.elapsed

    # metrics (inference.py:285-326)
    trajs_iters = dataset.unnormalize_trajectories(trajs_normalized_iters)
    trajs_final = trajs_iters[-1]
    trajs_final_coll, trajs_final_coll_idxs, trajs_final_free, trajs_final_free_idxs, _ = \
        task.get_trajs_collision_and_free(trajs_final, return_indices=True)
    results = dict(t_total=t_total, success=task.compute_success_free_trajs(trajs_final),
                   fraction_free=task.compute_fraction_free_trajs(trajs_final),
                   collision_intensity=task.compute_collision_intensity_trajs(trajs_final),
                   trajs_iters_shape=tuple(trajs_iters.shape), hard_conds=hard_conds, trajs_final_normalized=trajs_normalized_iters[-1])
    if trajs_final_free is not None:
        cost_smoothness = compute_smoothness(trajs_final_free, robot)
        cost_path_length = compute_path_length(trajs_final_free, robot)
        cost_all = cost_path_length + cost_smoothness
        results.update(cost_smoothness=float(cost_smoothness.mean()), cost_path_length=float(cost_path_length.mean()),
                       idx_best_traj=int(torch.argmin(cost_all)), cost_best=float(torch.min(cost_all)),
                       variance_waypoints=compute_variance_waypoints(trajs_final_free, robot))
    if verbose:
        print({k: v for k, v in results.items() if k not in ('hard_conds', 'trajs_final_normalized')})
    return results


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-id", default="EnvSpheres3D-RobotPanda")
    ap.add_argument("--planner-alg", default="mpd")
    ap.add_argument("--n-samples", type=int, default=50)
    ap.add_argument("--model-dir", default=None)
    a = ap.parse_args()
    run_experiment(experiment, model_id=a.model_id, planner_alg=a.planner_alg, n_samples=a.n_samples, model_dir=a.model_dir)
