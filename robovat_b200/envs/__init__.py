"""Environment layer: same entry points as robovat.envs (robovat/envs/__init__.py:1-2)."""
from robovat_b200.envs.push_env import PushEnv  # noqa: F401
