"""sgqn-carla_b200: B200-native (sm_100a) SGSAC / SAC / SVEA update path, drop-in for the agent API of
gferraro2019/SGQN-CARLA (`make_agent` -> `update` / `select_action` / `sample_action`)."""
from . import _lib  # noqa: F401
from .agents import SAC, RAD, DrQ, SVEA, SGSAC, CURL, PAD, SODA, make_agent, algorithm  # noqa: F401
from .replay import ReplayBuffer, LazyFrames  # noqa: F401
from .arguments import default_args  # noqa: F401
