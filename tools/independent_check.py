#!/usr/bin/env python
"""10^4 skew rays per fixture: oracle (orc_trace3d) against the independent 50-digit vector-geometry tracer
(oracle/independent_tracer.py).  Writes profiles/r02_oracle_vs_independent_tracer.json.  CPU only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ort_b200 as ort  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_oracle_independent import FIXTURES, compare  # noqa: E402

orc.build()
out = {"_what": "max |oracle - 50-digit independent tracer| over 10^4 random skew rays per fixture: hit coordinates at every "
                "surface relative to the largest coordinate of the ray, final direction cosines absolute",
       "fixtures": {name: compare(orc, ort, name, 10000, seed=1) for name in FIXTURES}}
s = json.dumps(out, indent=1)
open(os.path.join(ROOT, "profiles", "r02_oracle_vs_independent_tracer.json"), "w").write(s)
print(s)
