"""Names-only stand-in for `stable_baselines3` (oracle process only).

TEST INFRASTRUCTURE.  stable-baselines3 (pinned ~=2.2.1 by the reference's setup.py:206)
is not installed here.  The reference imports it for base-class names, the VecEnv
wrapper protocol, the SB3 Logger, and two trivially specified preprocessing helpers
(SURVEY.md Appendix B).  The PPO arithmetic is NOT here; it is restated in
`oracle/ppo_port.py` and declared "parity unpinned".
"""
from .ppo import PPO  # noqa: F401
