from gym.spaces.box import Box
from gym.spaces.discrete import Discrete
