"""B200-native batched drop-in for the `env.step()` path of pybullet_multigoal_gym.

`make_env` keeps the reference's signature (/root/reference/pybullet_multigoal_gym/__init__.py:4-11)
and adds `batch` (number of environments stepped in lockstep; None = one unbatched env with the
reference's numpy shapes), `device`, `seed`, and the throughput options `device_sampling` (resets drawn on the
device from Philox streams instead of the reference's host MT19937 streams) / `auto_reset` (environments reset
themselves on the device when their episode ends).  Tasks on the accelerated path: reach, push,
pick_and_place, slide, block_stack, block_rearrange with the parallel-jaw gripper and state observations,
including the `joint_control`, (block_stack) `grip_informed_goal` / `task_decomposition` and (block_stack, block_rearrange)
`use_curriculum` variants.
"""
from .envs import (ActionError, KukaBlockRearrangeEnv, KukaBlockStackEnv, KukaBulletMGEnv,  # noqa: F401
                   KukaPickAndPlaceEnv, KukaPushEnv, KukaReachEnv, KukaSlideEnv)

__all__ = ["make_env", "KukaBulletMGEnv", "KukaReachEnv", "KukaPushEnv", "KukaPickAndPlaceEnv", "KukaSlideEnv",
           "KukaBlockStackEnv", "KukaBlockRearrangeEnv", "ActionError"]

_TASKS = ['push', 'reach', 'slide', 'pick_and_place',
          'block_stack', 'block_rearrange', 'chest_pick_and_place', 'chest_push',
          'primitive_push_assemble', 'primitive_push_reach', 'insertion']
_TAGS = {'reach': 'Reach', 'push': 'Push', 'pick_and_place': 'PickAndPlace', 'slide': 'Slide', 'block_stack': 'BlockStack',
         'block_rearrange': 'BlockRearrangeEnv'}  # __init__.py:21-40 (the rearrange tag really ends in 'Env')
_ENTRY = {'reach': KukaReachEnv, 'push': KukaPushEnv, 'pick_and_place': KukaPickAndPlaceEnv, 'slide': KukaSlideEnv,
          'block_stack': KukaBlockStackEnv, 'block_rearrange': KukaBlockRearrangeEnv}


def make_env(task='reach', gripper='parallel_jaw', num_block=5, render=False, binary_reward=True,
             grip_informed_goal=False, task_decomposition=False,
             joint_control=False, max_episode_steps=50, distance_threshold=0.05,
             primitive=None,
             image_observation=False, depth_image=False, goal_image=False, point_cloud=False, state_noise=False,
             visualize_target=True,
             camera_setup=None, observation_cam_id=None, goal_cam_id=0,
             use_curriculum=False, num_goals_to_generate=1e6,
             batch=None, device=0, seed=0, check_actions=True, device_sampling=False, auto_reset=False,
             env_index_base=0):
    grippers = ['robotiq85', 'parallel_jaw']
    assert gripper in grippers, 'invalid gripper: {}, only support: {}'.format(gripper, grippers)
    if task not in _TASKS:
        raise ValueError('invalid task name: {}, only support: {}'.format(task, _TASKS))
    unsupported = []
    if task not in _TAGS:
        unsupported.append("task=%r" % task)
    if gripper != 'parallel_jaw':
        unsupported.append("gripper=%r" % gripper)
    if grip_informed_goal and task != 'block_stack':
        if task == 'block_rearrange':  # kuka_multi_step_envs.py:158
            raise AssertionError("Block rearranging task does not support gripper informed goal representation.")
        grip_informed_goal = False  # the single-step tasks do not take the kwarg (__init__.py:88-106)
    if task_decomposition and task != 'block_stack':
        if task == 'block_rearrange':  # kuka_multi_step_envs.py:159
            raise AssertionError("Block rearranging task does not support task decomposition.")
        if task in _TAGS:
            task_decomposition = False  # the single-step tasks do not take the kwarg (__init__.py:88-106)
    if use_curriculum and task in ('reach', 'push', 'pick_and_place', 'slide'):
        use_curriculum = False  # the single-step tasks do not take the kwarg (__init__.py:88-106)
    for name, val in (('render', render),
                      ('image_observation', image_observation), ('depth_image', depth_image),
                      ('goal_image', goal_image), ('point_cloud', point_cloud), ('state_noise', state_noise),
                      ('use_curriculum', use_curriculum and task not in ('block_stack', 'block_rearrange'))):
        if val:
            unsupported.append("%s=True" % name)
    if primitive is not None:
        unsupported.append("primitive=%r" % primitive)
    if unsupported:
        raise NotImplementedError(
            "outside the accelerated step path (reach/push/pick_and_place/slide/block_stack/block_rearrange, "
            "parallel_jaw, state observations): " + ", ".join(unsupported))
    if task in ('block_stack', 'block_rearrange'):
        assert num_block <= 5, "only support up to 5 blocks"
    env_id = ('Kuka' + _TAGS[task] + 'ParallelGrip' + ('SparseReward' if binary_reward else 'DenseReward') +
              ('JointCtrl' if joint_control else '') + '-v0')
    print('Task id: %s' % env_id)  # __init__.py:56-84
    return _ENTRY[task](batch=batch, device=device, binary_reward=binary_reward,
                        distance_threshold=distance_threshold, max_episode_steps=max_episode_steps,
                        num_block=num_block, seed=seed, check_actions=check_actions,
                        grip_informed_goal=grip_informed_goal, joint_control=joint_control,
                        task_decomposition=task_decomposition,
                        use_curriculum=use_curriculum, num_goals_to_generate=num_goals_to_generate,
                        device_sampling=device_sampling, auto_reset=auto_reset, env_index_base=env_index_base)
