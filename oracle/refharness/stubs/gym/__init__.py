"""Minimal stand-in for gym==0.12.5 -- TEST INFRASTRUCTURE ONLY.

The reference (zfw1226/active_tracking_rl) pins gym==0.12.5 (requirements.txt:1), which is not
installed in this image.  This stub restates exactly the pieces the reference's 2D path touches
(track_1v1.py:4-6, gym_track2d/__init__.py:1, environment.py:2,6): Env, Wrapper, ObservationWrapper,
spaces.Discrete/Box, utils.seeding.np_random, envs.registration.register, make + TimeLimit.
It exists so that oracle/refharness can import the UNMODIFIED reference from /root/reference in the
build container and record golden vectors.  Nothing in the product imports it.
"""
from gym.core import Env, Wrapper, ObservationWrapper
from gym import spaces, utils, envs
from gym.envs.registration import make, register

__version__ = "0.12.5-stub"
