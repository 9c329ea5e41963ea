"""Tile layouts of the PushEnv tasks.

Same public surface as the reference module robovat/envs/push/layouts.py
(`PushLayout`, `TASK_NAME_TO_LAYOUTS[task][layout_id]`, :9-245), but stored as
compact "row col" digit strings and expanded at import time; the expanded tables
are checked tile for tile against tests/golden/layouts.json, which was dumped
from the reference module.

Tile (r, c) is centred at offset + (r, c) * size on the 0.76 m x 1.22 m table
centred at (0.6, 0) (reference comment at layouts.py:30).
"""
import collections

PushLayout = collections.namedtuple(
    'PushLayout', 'size offset region goal target obstacle region_rgba goal_rgba')

_SIZE = 0.15
_OFFSET = (0.295, -0.485)
_BLUE = [0.4667, 0.7098, 0.9961, 1]
_RED = [1, .4235, .4235, 1]
_SAND = [0.867, 0.776, 0.678, 0]
_GREY = [0.8, 0.8, 0.8, 1]
_YELLOW = [1, 0.9412, 0.4235, 1]


def _tiles(spec):
    """'02 13' -> [[0, 2], [1, 3]]; None stays None."""
    if spec is None:
        return None
    return [[int(tok[0]), int(tok[1])] for tok in spec.split()]


def _grid(rows, cols):
    return ' '.join('%d%d' % (r, c) for r in rows for c in cols)


def _layout(region, goal, target, obstacle, region_rgba, goal_rgba):
    return PushLayout(size=_SIZE, offset=list(_OFFSET), region=_tiles(region), goal=_tiles(goal),
                      target=_tiles(target), obstacle=_tiles(obstacle),
                      region_rgba=list(region_rgba), goal_rgba=None if goal_rgba is None else list(goal_rgba))


_G36 = _grid((1, 2, 3), range(1, 7))       # the 18-tile obstacle grid of the crossing task
_INS_REGION = '00 01 10 11 20 30 31 40 41'

TASK_NAME_TO_LAYOUTS = {
    'clearing': [
        _layout(_grid((0, 1, 2), (2, 3, 4, 5)), None, '13 14 23 24', '13 14 23 24', _BLUE, None),
        _layout(_grid((1, 2), (2, 3, 4, 5)), None, _grid((1, 2), (2, 3, 4, 5)), _grid((1, 2), (2, 3, 4, 5)),
                _BLUE, None),
        _layout(_grid((0, 1), (2, 3, 4, 5)) + ' 23 24', None, _grid((0, 1), (2, 3, 4, 5)) + ' 23 24',
                _grid((0, 1), (2, 3, 4, 5)) + ' 23 24', _BLUE, None),
    ],
    'insertion': [
        _layout(_INS_REGION, '21', '23 24', _grid((1, 2, 3), (3, 4)), _RED, _SAND),
        _layout(_INS_REGION, '21', '23 24', _grid((1, 2, 3), (3, 4)), _RED, _SAND),
        _layout('31 32 35 36 41 42 43 44 45 46', '21', '12 13 14 15', _grid((1, 2), (2, 3, 4, 5)), _RED, _SAND),
    ],
    'crossing': [
        _layout('00 02 05 10 11 12 15 22 23 24 25 32', '12', '25', _G36, _GREY, _YELLOW),
        _layout('00 01 02 05 10 12 13 14 15 20 22 25 32 35', '32', '14 15', _G36, _GREY, _YELLOW),
        _layout('02 05 12 15 16 21 22 23 24 25 31 32 35', '16', '12 21 22', _G36, _GREY, _YELLOW),
    ],
}
