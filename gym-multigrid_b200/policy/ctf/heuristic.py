"""Scripted CtF opponents with the reference's interface and behaviour (policy/ctf/heuristic.py:18-463):
`RwPolicy`, `DestinationPolicy` and its four targets `FightPolicy`, `CapturePolicy`, `PatrolPolicy`, `PatrolFightPolicy`.

Drop-in for `enemy_policies=` of `CtFMvNEnv` (ctf.py:666): same constructor arguments, same attributes (`name`,
`field_map`, `action_set`, `random_generator`, `randomness`, `ego_agent`, `border`, `obstacle`), `act(observation_dict,
curr_pos)` returns the action the reference returns AND draws from `random_generator` with the same numpy calls in the
same order, so a generator shared with other consumers (the reference env shares its `np_random` with every policy,
ctf.py:821-826) stays in step.  Checked against the unmodified reference classes on recorded decisions
(`tests/golden/ctf_policies.npz`, `tests/test_policies.py`).

These run on the host, like in the reference; the red actions they produce reach the CUDA step through
`mg_set_red_actions`.  What differs is cost: the route towards a target is memoised per (cell, target) - it does not
depend on anything else (see `utils.a_star`) - and the patrol candidates are built once, not per call."""
from __future__ import annotations

import numpy as np

from ...actions import CtfActions
from ...world import CtfWorld
from ..base import BaseAgentPolicy
from .utils import NEIGHBOUR_ORDER, a_star, closest_area_pos


class CtfPolicy(BaseAgentPolicy):
    """heuristic.py:18-37"""

    def act(self, observation, curr_pos) -> int:
        raise NotImplementedError


class RwPolicy(CtfPolicy):
    """heuristic.py:40-72: uniform random action.  As an `enemy_policies` entry of the adaptors it selects the built-in
    opponent, whose draws come from the env's Philox stream on the device; `act` is for callers that drive it by hand."""

    def __init__(self, action_set=CtfActions, random_generator=None):
        super().__init__(action_set, random_generator)
        self.name = "rw"

    def act(self, observation=None, curr_pos=None) -> int:
        return self.random_generator.integers(0, len(self.action_set))


def _team_keys(ego_agent):
    """(ego territory, opponent territory, opponent agents, opponent flag) observation keys for `ego_agent`."""
    ego, opp = ("red", "blue") if ego_agent == "red" else ("blue", "red")   # anything but "red" is blue, as heuristic.py:217-219
    return f"{ego}_territory", f"{opp}_territory", f"{opp}_agent", f"{opp}_flag"


class DestinationPolicy(CtfPolicy):
    """heuristic.py:75-177: with probability `randomness` the first move of the route to `get_target(...)`, otherwise a
    uniform random action."""

    def __init__(self, field_map=None, action_set=CtfActions, random_generator=None, randomness: float = 0.75):
        super().__init__(action_set, random_generator)
        self.name = "destination"
        self.field_map = field_map
        self.randomness = randomness
        self._routes, self._routes_map = {}, None

    def get_target(self, observation, curr_pos):
        return None     # heuristic.py:108-123 leaves it to the subclasses

    def _next_cell(self, start, target):
        """Second cell of `a_star(start, target, field_map)`, or `target` when the route has no second cell
        (start == target, or no route: heuristic.py:144-147)."""
        if self._routes_map is not self.field_map:      # the env installs the map after construction (ctf.py:796-799)
            self._routes, self._routes_map = {}, self.field_map
        key = (start, target)
        nxt = self._routes.get(key)
        if nxt is None:
            path = a_star(start, target, self.field_map)
            nxt = self._routes[key] = path[1] if len(path) > 1 else target
        return nxt

    def act(self, observation, curr_pos) -> int:
        start = (int(curr_pos[0]), int(curr_pos[1]))
        goal = self.get_target(observation, curr_pos)
        target = (int(goal[0]), int(goal[1]))
        nxt = self._next_cell(start, target)
        # heuristic.py:150-152: this draw happens whether or not the optimal move is taken
        follow_route = self.random_generator.choice([True, False], p=[self.randomness, 1 - self.randomness])
        if not follow_route:
            return self.random_generator.integers(0, len(self.action_set))
        step = (nxt[0] - start[0], nxt[1] - start[1])
        acts = self.action_set
        if step == (0, 0):
            return acts.stay
        if step == (0, -1):
            return acts.left
        if step == (-1, 0):
            return acts.down
        if step == (0, 1):
            return acts.right
        if step == (1, 0):
            return acts.up
        raise ValueError("Invalid direction")       # heuristic.py:172


class FightPolicy(DestinationPolicy):
    """heuristic.py:180-226: head for the closest opponent agent (terminated ones included - the observation still
    lists them)."""

    def __init__(self, field_map=None, action_set=CtfActions, random_generator=None, randomness: float = 0.75,
                 ego_agent="red"):
        super().__init__(field_map, action_set, random_generator, randomness)
        self.name = "fight"
        self.ego_agent = ego_agent

    def _opponents(self, observation):
        return [tuple(p) for p in np.asarray(observation[_team_keys(self.ego_agent)[2]]).reshape(-1, 2)]

    def get_target(self, observation, curr_pos):
        return closest_area_pos(curr_pos, self._opponents(observation))


class CapturePolicy(DestinationPolicy):
    """heuristic.py:229-272: head for the opponent's flag."""

    def __init__(self, field_map=None, action_set=CtfActions, random_generator=None, randomness: float = 0.75,
                 ego_agent="red"):
        super().__init__(field_map, action_set, random_generator, randomness)
        self.name = "capture"
        self.ego_agent = ego_agent

    def get_target(self, observation, curr_pos):
        if self.ego_agent not in ("red", "blue"):
            return None                               # the reference's `match` falls through (heuristic.py:266-272)
        key = _team_keys(self.ego_agent)[3]
        assert key in observation
        return observation[key]


class PatrolPolicy(DestinationPolicy):
    """heuristic.py:275-391: walk to the border and then along it.

    `border` is what the reference computes (heuristic.py:340-391), quirks included: for every own-territory cell, in
    `np.where` order, the FIRST neighbour (order (0,+1), (0,-1), (+1,0), (-1,0)) that is opponent territory or an
    obstacle is appended - so the border lies on the far side, may list a cell more than once, and is computed once,
    at construction (a policy built without `field_map` keeps an empty border and raises on its first decision, as
    the reference does).  On the border the next target is drawn uniformly from the border cells that have a border
    neighbour - all of them, wherever the agent stands (heuristic.py:322-335)."""

    def __init__(self, field_map=None, action_set=CtfActions, random_generator=None, randomness: float = 0.75,
                 ego_agent="red", world=CtfWorld):
        super().__init__(field_map, action_set, random_generator, randomness)
        self.name = "patrol"
        self.ego_agent = ego_agent
        self.world = world
        self.directions = list(NEIGHBOUR_ORDER)
        self.border, self.obstacle = self.locate_border(world, self.directions)
        self._border_cells = {(int(p[0]), int(p[1])) for p in self.border}
        # one candidate per (border cell, direction) pair whose neighbour is a border cell; duplicates weigh the draw
        self._along_border = np.array([(p[0] + d[0], p[1] + d[1]) for p in self.border for d in self.directions
                                       if (int(p[0] + d[0]), int(p[1] + d[1])) in self._border_cells], dtype=np.int64)

    def locate_border(self, world, directions):
        assert self.world is not None
        idx = world.OBJECT_TO_IDX
        own = "red_territory" if self.ego_agent == "red" else "blue_territory"
        other = "red_territory" if self.ego_agent == "blue" else "blue_territory"
        if self.field_map is None:
            return [], []
        fm = np.asarray(self.field_map)
        cells = lambda name: list(zip(*np.where(fm == idx[name])))   # noqa: E731
        obstacle = cells("obstacle")
        far_side = {(int(x), int(y)) for x, y in cells(other) + obstacle}
        border = []
        for x, y in cells(own):
            for dx, dy in directions:
                if (int(x + dx), int(y + dy)) in far_side:
                    border.append((x + dx, y + dy))
                    break
        return border, obstacle

    def get_target(self, observation, curr_pos):
        if (int(curr_pos[0]), int(curr_pos[1])) in self._border_cells:
            # heuristic.py:333: Generator.choice over the candidate list (rows of an int64 array) - same draw
            if len(self._along_border) == 0:
                return self.random_generator.choice([])
            return self.random_generator.choice(self._along_border)
        return closest_area_pos(curr_pos, self.border)


class PatrolFightPolicy(PatrolPolicy):
    """heuristic.py:394-463: patrol; while any opponent stands on the ego territory, chase the closest opponent."""

    def __init__(self, field_map=None, action_set=CtfActions, random_generator=None, randomness: float = 0.75,
                 ego_agent="red", world=CtfWorld):
        super().__init__(field_map, action_set, random_generator, randomness, ego_agent, world)
        self.name = "patrol_fight"

    def get_target(self, observation, curr_pos):
        ego_territory, _, opp_agents, _ = _team_keys(self.ego_agent)
        opp = np.asarray(observation[opp_agents]).reshape(-1, 2)
        home = np.asarray(observation[ego_territory]).reshape(-1, 2)
        if (opp[:, None, :] == home[None, :, :]).all(axis=2).any():      # heuristic.py:448-455, all pairs at once
            return closest_area_pos(curr_pos, [tuple(p) for p in opp])
        return super().get_target(observation, curr_pos)
