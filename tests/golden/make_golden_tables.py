"""Dump the reference's constant tables (pitch points, plane sets, line tables)
to tests/golden/tables.json.  Run in the build container only:
    python tests/golden/make_golden_tables.py
Sources: src/datatools/ellipse.py:16-185, src/models/hrnet/prediction.py:15-26,
src/datatools/intersections.py:13-44, src/datatools/line.py:35-57."""
import json
import os

import refimport

refimport.setup()
from src.datatools.ellipse import (INTERSECTON_TO_PITCH_POINTS, PITCH_POINTS,  # noqa: E402
                                   POINTS_LEFT, POINTS_RIGHT)
from src.datatools.intersections import LINE_INTERSECTIONS  # noqa: E402
from src.datatools.line import LINE_CLS  # noqa: E402
from src.models.hrnet import prediction as P  # noqa: E402

out = {
    "pitch_points": {k: [float(x) for x in v] for k, v in PITCH_POINTS.items()},
    "id_to_name": {str(k): v for k, v in INTERSECTON_TO_PITCH_POINTS.items()},
    "points_left": POINTS_LEFT, "points_right": POINTS_RIGHT,
    "top_gates": P.top_gates, "point_sets": P.point_sets, "keep_points": P.keep_points,
    "img_size": list(P.IMG_SIZE),
    "line_intersections": {str(k): list(v) for k, v in LINE_INTERSECTIONS.items()},
    "line_cls": {str(k): v for k, v in LINE_CLS.items()},
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", path)
