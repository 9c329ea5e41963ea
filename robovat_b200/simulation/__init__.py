from robovat_b200.simulation.simulator import Simulator  # noqa: F401
