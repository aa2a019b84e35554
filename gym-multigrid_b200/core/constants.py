"""`gym_multigrid.core.constants` data tables (core/constants.py:5-74) for host-side callers: colour palettes per world, colour /
state index maps, direction vectors.  The kernels and the render atlas bake the same numbers in (csrc/render_kernels.cu,
csrc/mg_device.cuh); tests/test_abi_and_host.py checks this module against the oracle's rasteriser and, where it exists, the
reference's module."""
import numpy as np

TILE_PIXELS = 32                                                              # constants.py:5

_BASE = dict(red=(228, 3, 3), orange=(255, 140, 0), yellow=(255, 237, 0), green=(0, 128, 38), blue=(0, 77, 255),
             purple=(117, 7, 135), brown=(120, 79, 23), grey=(100, 100, 100))
_PALE = dict(light_red=(255, 228, 225), light_blue=(240, 248, 255), white=(255, 250, 250))


def _palette(**named):
    return {k: np.array(v) for k, v in named.items()}


COLORS = _palette(**_BASE, light_red=(234, 153, 153), light_blue=(90, 170, 223))                       # constants.py:8-19
MAZE_COLORS = _palette(**_BASE, **_PALE)                                                               # constants.py:37-49
CTF_COLORS = _palette(**_BASE, **_PALE, red_grey=(170, 152, 169), blue_grey=(140, 146, 172))           # constants.py:21-35

COLOR_NAMES = sorted(COLORS)
COLOR_TO_IDX = {name: i for i, name in enumerate(COLORS)}                     # red 0 ... grey 7, light_red 8, light_blue 9
IDX_TO_COLOR = {i: name for name, i in COLOR_TO_IDX.items()}
STATE_TO_IDX = dict(open=0, closed=1, locked=2)                               # door states (object.py:238-259)
DIR_TO_VEC = [np.array(v) for v in ((1, 0), (0, 1), (-1, 0), (0, -1))]       # right (+x), down (+y), left, up (constants.py:65-74)
