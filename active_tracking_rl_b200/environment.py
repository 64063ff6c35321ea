"""create_env (environment.py:11-32) for the 2D ids: the batched env already delivers what frame_stack
(environment.py:128-156) produces for stack_frames = 1 -- float32 observations with a unit stack axis."""
from .envs import Track2DVecEnv


def create_env(env_id, args, num_envs=None, device=None, seed=None, rng="philox"):
    if '2D' not in env_id:
        raise NotImplementedError("only the Track2D ids are built (UnrealCV envs are out of scope)")
    if int(getattr(args, 'stack_frames', 1)) != 1:
        raise NotImplementedError("stack_frames != 1 is not used by any 2D command of the reference")
    if getattr(args, 'rescale', False) or getattr(args, 'single', False):
        raise NotImplementedError("--rescale / --single belong to the 3D image path")
    E = int(num_envs if num_envs is not None else getattr(args, 'num_envs', 1))
    dev = device if device is not None else getattr(args, 'device', 'cuda:0')
    sd = seed if seed is not None else getattr(args, 'seed', 1)
    return Track2DVecEnv(env_id, num_envs=E, device=dev, seed=sd, rng=rng, auto_reset=True, plan_ahead=bool(getattr(args, 'plan_ahead', False)))
