"""3-D soccer-pitch model used by the camera solve.

Follows the reference's pitch tables (all values in metres, origin at the centre
mark, x towards the right goal, y towards the bottom touch line, z NEGATIVE up):

* ``baseline/soccerpitch.py:109-263``  - FIFA-rule marks (corners, boxes, posts, arcs)
* ``src/datatools/ellipse.py:16-92``   - circle tangent / diagonal / axis points
* ``src/datatools/ellipse.py:99-157``  - keypoint id (0..56) -> point name
* ``src/datatools/ellipse.py:182-185`` - left/right point sets
* ``src/models/hrnet/prediction.py:15-26`` - plane sets, keep_points
* ``src/datatools/intersections.py:13-44`` - line pair -> keypoint id
* ``src/datatools/line.py:35-57``      - line-model channel -> line class name

The table is rebuilt here from the pitch dimensions (not copied); the golden test
``tests/test_pitch.py`` pins every coordinate against a dump of the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

PITCH_LENGTH = 105.0
PITCH_WIDTH = 68.0
GOAL_LINE_TO_PENALTY_MARK = 11.0
PENALTY_AREA_WIDTH = 40.32
PENALTY_AREA_LENGTH = 16.5
GOAL_AREA_WIDTH = 18.32
GOAL_AREA_LENGTH = 5.5
CENTER_CIRCLE_RADIUS = 9.15
GOAL_HEIGHT = 2.44
GOAL_LENGTH = 7.32

NUM_KEYPOINTS = 57
IMG_SIZE = (960, 540)


def _tangents(center, radius, ext):
    """Two tangent points on a circle seen from an external point (counter-clockwise
    first).  Same construction as ellipse.py:20-34 (acos of r/hyp around atan2)."""
    hyp = math.sqrt((ext[0] - center[0]) ** 2 + (ext[1] - center[1]) ** 2)
    th = np.arccos(radius / hyp)
    d = np.arctan2(ext[1] - center[1], ext[0] - center[0])
    out = []
    for a in (d + th, d - th):
        out.append((center[0] + radius * np.cos(a), center[1] + radius * np.sin(a), 0.0))
    return out


def _build() -> Dict[str, np.ndarray]:
    hl, hw = PITCH_LENGTH / 2.0, PITCH_WIDTH / 2.0
    r = CENTER_CIRCLE_RADIUS
    p: Dict[str, Tuple[float, float, float]] = {}
    p["CENTER_MARK"] = (0.0, 0.0, 0.0)
    p["T_TOUCH_AND_HALFWAY_LINES_INTERSECTION"] = (0.0, -hw, 0.0)
    p["B_TOUCH_AND_HALFWAY_LINES_INTERSECTION"] = (0.0, hw, 0.0)
    p["T_HALFWAY_LINE_AND_CENTER_CIRCLE_INTERSECTION"] = (0.0, -r, 0.0)
    p["B_HALFWAY_LINE_AND_CENTER_CIRCLE_INTERSECTION"] = (0.0, r, 0.0)
    arc_dx = PENALTY_AREA_LENGTH - GOAL_LINE_TO_PENALTY_MARK
    arc_y = math.sqrt(r * r - arc_dx * arc_dx)
    # per side: sx = -1 left half, +1 right half
    for side, sx in (("L", -1.0), ("R", 1.0)):
        gl = sx * hl                                   # goal line x
        p[f"{side}_PENALTY_MARK"] = (sx * (hl - GOAL_LINE_TO_PENALTY_MARK), 0.0, 0.0)
        for box, length, width in (("PENALTY", PENALTY_AREA_LENGTH, PENALTY_AREA_WIDTH),
                                   ("GOAL", GOAL_AREA_LENGTH, GOAL_AREA_WIDTH)):
            inner = sx * (hl - length)
            # "L"/"R" corner suffix is image-left / image-right of the box
            xl, xr = (gl, inner) if sx < 0 else (inner, gl)
            p[f"{side}_{box}_AREA_TL_CORNER"] = (xl, -width / 2.0, 0.0)
            p[f"{side}_{box}_AREA_TR_CORNER"] = (xr, -width / 2.0, 0.0)
            p[f"{side}_{box}_AREA_BL_CORNER"] = (xl, width / 2.0, 0.0)
            p[f"{side}_{box}_AREA_BR_CORNER"] = (xr, width / 2.0, 0.0)
        # posts: "left/right" as seen by the goal keeper looking at the pitch
        yl = GOAL_LENGTH / 2.0 if sx < 0 else -GOAL_LENGTH / 2.0
        p[f"{side}_GOAL_TL_POST"] = (gl, yl, -GOAL_HEIGHT)
        p[f"{side}_GOAL_TR_POST"] = (gl, -yl, -GOAL_HEIGHT)
        p[f"{side}_GOAL_BL_POST"] = (gl, yl, 0.0)
        p[f"{side}_GOAL_BR_POST"] = (gl, -yl, 0.0)
        x16 = sx * (hl - PENALTY_AREA_LENGTH)
        p[f"T{side}_16M_LINE_AND_PENALTY_ARC_INTERSECTION"] = (x16, -arc_y, 0.0)
        p[f"B{side}_16M_LINE_AND_PENALTY_ARC_INTERSECTION"] = (x16, arc_y, 0.0)
    p["TL_PITCH_CORNER"] = (-hl, -hw, 0.0)
    p["BL_PITCH_CORNER"] = (-hl, hw, 0.0)
    p["TR_PITCH_CORNER"] = (hl, -hw, 0.0)
    p["BR_PITCH_CORNER"] = (hl, hw, 0.0)

    # centre-circle tangents from the halfway-line / touch-line intersections
    top = _tangents((0.0, 0.0), r, p["T_TOUCH_AND_HALFWAY_LINES_INTERSECTION"][:2])
    bot = _tangents((0.0, 0.0), r, p["B_TOUCH_AND_HALFWAY_LINES_INTERSECTION"][:2])
    p["CENTER_CIRCLE_TANGENT_TR"], p["CENTER_CIRCLE_TANGENT_TL"] = top[0], top[1]
    p["CENTER_CIRCLE_TANGENT_BR"], p["CENTER_CIRCLE_TANGENT_BL"] = bot[1], bot[0]
    q = math.sqrt(2.0) * r / 2
    p["CENTER_CIRCLE_TR"] = (q, -q, 0.0)
    p["CENTER_CIRCLE_TL"] = (-q, -q, 0.0)
    p["CENTER_CIRCLE_BR"] = (q, q, 0.0)
    p["CENTER_CIRCLE_BL"] = (-q, q, 0.0)
    p["CENTER_CIRCLE_R"] = (r, 0.0, 0.0)
    p["CENTER_CIRCLE_L"] = (-r, 0.0, 0.0)
    lm, rm = p["L_PENALTY_MARK"], p["R_PENALTY_MARK"]
    p["LEFT_CIRCLE_R"] = (lm[0] + r, 0.0, 0.0)
    p["RIGHT_CIRCLE_L"] = (rm[0] - r, 0.0, 0.0)
    p["LEFT_CIRCLE_TANGENT_T"] = _tangents(lm[:2], r, p["L_PENALTY_AREA_TR_CORNER"][:2])[0]
    p["LEFT_CIRCLE_TANGENT_B"] = _tangents(lm[:2], r, p["L_PENALTY_AREA_BR_CORNER"][:2])[1]
    p["RIGHT_CIRCLE_TANGENT_T"] = _tangents(rm[:2], r, p["R_PENALTY_AREA_TL_CORNER"][:2])[1]
    p["RIGHT_CIRCLE_TANGENT_B"] = _tangents(rm[:2], r, p["R_PENALTY_AREA_BL_CORNER"][:2])[0]
    p["L_MIDDLE_PENALTY"] = (p["L_PENALTY_AREA_BR_CORNER"][0], 0.0, 0.0)
    p["R_MIDDLE_PENALTY"] = (p["R_PENALTY_AREA_BL_CORNER"][0], 0.0, 0.0)
    return {k: np.array(v, dtype=float) for k, v in p.items()}


PITCH_POINTS: Dict[str, np.ndarray] = _build()

# keypoint id -> name (the 57 heat-map channels, ellipse.py:99-157)
KEYPOINT_NAMES: List[str] = [
    "L_GOAL_TL_POST", "L_GOAL_TR_POST", "L_GOAL_BL_POST", "L_GOAL_BR_POST",
    "L_GOAL_AREA_BR_CORNER", "L_GOAL_AREA_TR_CORNER", "L_GOAL_AREA_BL_CORNER", "L_GOAL_AREA_TL_CORNER",
    "L_PENALTY_AREA_BR_CORNER", "L_PENALTY_AREA_TR_CORNER", "L_PENALTY_AREA_BL_CORNER",
    "L_PENALTY_AREA_TL_CORNER", "BL_PITCH_CORNER", "TL_PITCH_CORNER",
    "B_TOUCH_AND_HALFWAY_LINES_INTERSECTION", "T_TOUCH_AND_HALFWAY_LINES_INTERSECTION",
    "R_PENALTY_AREA_BL_CORNER", "R_PENALTY_AREA_TL_CORNER", "R_PENALTY_AREA_BR_CORNER",
    "R_PENALTY_AREA_TR_CORNER", "R_GOAL_AREA_BL_CORNER", "R_GOAL_AREA_TL_CORNER",
    "R_GOAL_AREA_BR_CORNER", "R_GOAL_AREA_TR_CORNER", "R_GOAL_TL_POST", "R_GOAL_TR_POST",
    "R_GOAL_BL_POST", "R_GOAL_BR_POST", "BR_PITCH_CORNER", "TR_PITCH_CORNER",
    "CENTER_CIRCLE_TANGENT_TR", "CENTER_CIRCLE_TANGENT_TL", "CENTER_CIRCLE_TANGENT_BR",
    "CENTER_CIRCLE_TANGENT_BL", "CENTER_CIRCLE_TR", "CENTER_CIRCLE_TL", "CENTER_CIRCLE_BR",
    "CENTER_CIRCLE_BL", "CENTER_CIRCLE_R", "CENTER_CIRCLE_L",
    "T_HALFWAY_LINE_AND_CENTER_CIRCLE_INTERSECTION", "B_HALFWAY_LINE_AND_CENTER_CIRCLE_INTERSECTION",
    "CENTER_MARK", "LEFT_CIRCLE_R", "BL_16M_LINE_AND_PENALTY_ARC_INTERSECTION",
    "TL_16M_LINE_AND_PENALTY_ARC_INTERSECTION", "LEFT_CIRCLE_TANGENT_T", "LEFT_CIRCLE_TANGENT_B",
    "L_PENALTY_MARK", "L_MIDDLE_PENALTY", "RIGHT_CIRCLE_L", "BR_16M_LINE_AND_PENALTY_ARC_INTERSECTION",
    "TR_16M_LINE_AND_PENALTY_ARC_INTERSECTION", "RIGHT_CIRCLE_TANGENT_T", "RIGHT_CIRCLE_TANGENT_B",
    "R_PENALTY_MARK", "R_MIDDLE_PENALTY",
]
assert len(KEYPOINT_NAMES) == NUM_KEYPOINTS
INTERSECTON_TO_PITCH_POINTS: Dict[int, str] = dict(enumerate(KEYPOINT_NAMES))  # reference spelling


def get_pitch() -> Dict[str, np.ndarray]:
    """Hydra target ``src.datatools.ellipse.get_pitch`` (ellipse.py:95-96)."""
    return PITCH_POINTS


def keypoint_world_table(pitch: Dict[str, np.ndarray] | None = None) -> np.ndarray:
    """(57, 3) float64 world coordinates in keypoint-id order."""
    pitch = PITCH_POINTS if pitch is None else pitch
    return np.stack([np.asarray(pitch[n], dtype=np.float64) for n in KEYPOINT_NAMES], axis=0)


# plane sets (prediction.py:15-26). top_gates = crossbar ends (z = -2.44).
TOP_GATES: List[int] = [0, 1, 24, 25]
POINT_SETS: Dict[str, List[int]] = {
    "groundplane": [i for i in range(58) if i not in TOP_GATES],
    "goal_left": [0, 1, 2, 3, 6, 7, 10, 11, 12, 13],
    "goal_right": [18, 19, 22, 23, 24, 25, 26, 27, 28, 29],
}
KEEP_POINTS: List[int] = list(range(29)) + [40, 41, 42, 44, 45, 48, 51, 52, 55]
POINTS_LEFT: List[int] = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 31, 33,
                          35, 37, 39, 43, 44, 45, 46, 47, 48, 49]
POINTS_RIGHT: List[int] = [16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28,
                           29, 30, 32, 34, 36, 38, 50, 51, 52, 53, 54, 55, 56]

# line-model channel -> class name (line.py:35-57); note the trailing blank in
# 'Goal left post left ' is the dataset's own spelling.
LINE_CLS: Dict[int, str] = dict(enumerate([
    "Goal left post left ", "Goal right post right", "Middle line", "Small rect. right top",
    "Side line bottom", "Goal right post left", "Big rect. right main", "Goal left crossbar",
    "Small rect. left bottom", "Side line left", "Big rect. right top", "Small rect. left top",
    "Side line right", "Big rect. left top", "Goal left post right", "Small rect. right bottom",
    "Side line top", "Goal right crossbar", "Small rect. left main", "Big rect. left main",
    "Big rect. right bottom", "Small rect. right main", "Big rect. left bottom",
]))


def _line_pairs() -> Dict[int, Tuple[str, str]]:
    """keypoint id -> the two line classes whose intersection it is
    (intersections.py:13-44): ids 0..29 are exactly the line/line crossings."""
    pairs: Dict[int, Tuple[str, str]] = {}
    for side, base in (("left", 0), ("right", 16)):
        # left: ids 0..13 ; right mirrored ids 16..29
        post_l = "Goal left post left " if side == "left" else "Goal right post left"
        post_r = f"Goal {side} post right"
        cross, sl = f"Goal {side} crossbar", f"Side line {side}"
        small, big = f"Small rect. {side}", f"Big rect. {side}"
        if side == "left":
            pairs.update({0: (cross, post_l), 1: (cross, post_r), 2: (sl, post_l), 3: (sl, post_r),
                          4: (f"{small} main", f"{small} bottom"), 5: (f"{small} main", f"{small} top"),
                          6: (sl, f"{small} bottom"), 7: (sl, f"{small} top"),
                          8: (f"{big} main", f"{big} bottom"), 9: (f"{big} main", f"{big} top"),
                          10: (sl, f"{big} bottom"), 11: (sl, f"{big} top"),
                          12: (sl, "Side line bottom"), 13: (sl, "Side line top")})
        else:
            pairs.update({16: (f"{big} main", f"{big} bottom"), 17: (f"{big} main", f"{big} top"),
                          18: (sl, f"{big} bottom"), 19: (sl, f"{big} top"),
                          20: (f"{small} main", f"{small} bottom"), 21: (f"{small} main", f"{small} top"),
                          22: (sl, f"{small} bottom"), 23: (sl, f"{small} top"),
                          24: (cross, post_l), 25: (cross, post_r), 26: (sl, post_l), 27: (sl, post_r),
                          28: (sl, "Side line bottom"), 29: (sl, "Side line top")})
    pairs[14] = ("Middle line", "Side line bottom")
    pairs[15] = ("Middle line", "Side line top")
    return dict(sorted(pairs.items()))


LINE_INTERSECTIONS: Dict[int, Tuple[str, str]] = _line_pairs()


# ------------------------------------------------------------------ evaluation-side tables
# the 28 line classes of the dataset (baseline/soccerpitch.py:15-44), alphabetical as there
LINES_CLASSES: List[str] = sorted(
    [f"{box} rect. {side} {part}" for box in ("Big", "Small") for side in ("left", "right") for part in ("bottom", "main", "top")]
    + ["Circle central", "Circle left", "Circle right", "Goal left crossbar", "Goal left post left ", "Goal left post right",
       "Goal right crossbar", "Goal right post left", "Goal right post right", "Goal unknown", "Line unknown", "Middle line",
       "Side line bottom", "Side line left", "Side line right", "Side line top"])


def symmetric_class(name: str) -> str:
    """The class a line maps to under the point reflection through the pitch centre
    (SoccerPitch.symetric_classes, soccerpitch.py:46-75): left <-> right everywhere, top <-> bottom for the
    rectangle and side lines; the goal posts keep their own 'left' / 'right'."""
    flip = {"top": "bottom", "bottom": "top", "left": "right", "right": "left", "main": "main"}
    if name in ("Middle line", "Circle central", "Goal unknown", "Line unknown"):
        return name
    if name.startswith("Goal "):
        side, rest = name.rstrip()[5:].split(" ", 1)
        out = f"Goal {flip[side]} {rest}"
        return out + " " if out == "Goal left post left" else out      # the dataset's own spelling
    if name.startswith("Circle "):
        return "Circle " + flip[name[7:]]
    if name.startswith("Side line "):
        return "Side line " + flip[name[10:]]
    box, _, side, part = name.split(" ")           # '<Big|Small> rect. <side> <part>'
    return f"{box} rect. {flip[side]} {flip[part]}"


def line_extremities() -> Dict[str, Tuple[str, str]]:
    """class -> names of its two end points, in the reference's insertion order
    (SoccerPitch.line_extremities_keys, soccerpitch.py:318-375); the order fixes the sampling direction."""
    ext: Dict[str, Tuple[str, str]] = {}
    for box, tag in (("Big", "PENALTY_AREA"), ("Small", "GOAL_AREA")):
        for side, s in (("left", "L"), ("right", "R")):
            ext[f"{box} rect. {side} bottom"] = (f"{s}_{tag}_BL_CORNER", f"{s}_{tag}_BR_CORNER")
            ext[f"{box} rect. {side} top"] = (f"{s}_{tag}_TL_CORNER", f"{s}_{tag}_TR_CORNER")
            ext[f"{box} rect. {side} main"] = ((f"{s}_{tag}_TR_CORNER", f"{s}_{tag}_BR_CORNER") if side == "left"
                                               else (f"{s}_{tag}_TL_CORNER", f"{s}_{tag}_BL_CORNER"))
    ext["Side line top"] = ("TL_PITCH_CORNER", "TR_PITCH_CORNER")
    ext["Side line bottom"] = ("BL_PITCH_CORNER", "BR_PITCH_CORNER")
    ext["Side line left"] = ("TL_PITCH_CORNER", "BL_PITCH_CORNER")
    ext["Side line right"] = ("TR_PITCH_CORNER", "BR_PITCH_CORNER")
    ext["Middle line"] = ("T_TOUCH_AND_HALFWAY_LINES_INTERSECTION", "B_TOUCH_AND_HALFWAY_LINES_INTERSECTION")
    ext["Goal left crossbar"] = ("L_GOAL_TR_POST", "L_GOAL_TL_POST")
    ext["Goal left post left "] = ("L_GOAL_TL_POST", "L_GOAL_BL_POST")
    ext["Goal left post right"] = ("L_GOAL_TR_POST", "L_GOAL_BR_POST")
    ext["Goal right crossbar"] = ("R_GOAL_TL_POST", "R_GOAL_TR_POST")
    ext["Goal right post left"] = ("R_GOAL_TL_POST", "R_GOAL_BL_POST")
    ext["Goal right post right"] = ("R_GOAL_TR_POST", "R_GOAL_BR_POST")
    ext["Circle right"] = ("TR_16M_LINE_AND_PENALTY_ARC_INTERSECTION", "BR_16M_LINE_AND_PENALTY_ARC_INTERSECTION")
    ext["Circle left"] = ("TL_16M_LINE_AND_PENALTY_ARC_INTERSECTION", "BL_16M_LINE_AND_PENALTY_ARC_INTERSECTION")
    return ext


def sample_field_points(dist: float = 0.1, dist_circles: float = 0.2) -> Dict[str, np.ndarray]:
    """class -> (n, 3) points sampled along the pitch element (SoccerPitch.sample_field_points,
    soccerpitch.py:420-510): the centre circle every ``dist_circles`` metres of arc from angle 0, the two
    penalty arcs between their 16 m-line marks (bottom -> top on the right, top -> bottom on the left, end
    point appended), straight lines every ``dist`` metres from the first extremity (running sum, end
    point appended)."""
    P = get_pitch()
    r = CENTER_CIRCLE_RADIUS
    out: Dict[str, np.ndarray] = {}

    def arc(center, a0, a1, closed):
        if a1 < a0:
            a1 += 2 * np.pi
        n = int(r * (a1 - a0) / dist_circles)
        da = dist_circles / r
        pts = [np.array((center[0] + np.cos(a0) * r, center[1] + np.sin(a0) * r, 0.0))]
        for i in range(1, n if closed else n + 1):
            a = a0 + i * da
            pts.append(np.array((center[0] + np.cos(a) * r, center[1] + np.sin(a) * r, 0)))
        if not closed:
            pts.append(np.array((center[0] + np.cos(a1) * r, center[1] + np.sin(a1) * r, 0.0)))
        return np.array(pts, dtype=np.float64)

    out["Circle central"] = arc(P["CENTER_MARK"], 0.0, 2 * np.pi, True)
    for key, (na, nb) in line_extremities().items():
        a, b = np.asarray(P[na], dtype=np.float64), np.asarray(P[nb], dtype=np.float64)
        if key.startswith("Circle"):
            c = np.asarray(P["R_PENALTY_MARK" if key == "Circle right" else "L_PENALTY_MARK"], dtype=np.float64)
            ang = lambda q: np.arctan2(q[1] - c[1], q[0] - c[0]) + 2 * np.pi
            a0, a1 = (ang(b), ang(a)) if key == "Circle right" else (ang(a), ang(b))
            out[key] = arc(c, a0, a1, False)
        else:
            total = np.sqrt(np.sum(np.square(a - b)))
            n = int(total / dist - 1)
            v = b - a
            v = v / np.linalg.norm(v)
            pts, prev = [a], a
            for _ in range(n):
                prev = prev + dist * v
                pts.append(prev)
            pts.append(b)
            out[key] = np.array(pts, dtype=np.float64)
    return out
