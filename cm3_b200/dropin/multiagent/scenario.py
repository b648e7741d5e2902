"""Scenario plugin protocol (the reference's multiagent/scenario.py:4-10): a scenario builds a
world and re-initialises it; reward / observation / done are optional per-agent callbacks."""
import abc


class BaseScenario(abc.ABC):
    @abc.abstractmethod
    def make_world(self, *args, **kwargs):
        """Create the World and its entities."""

    @abc.abstractmethod
    def reset_world(self, world):
        """Draw the initial conditions of an episode."""
