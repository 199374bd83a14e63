#!/usr/bin/env python3
"""Transcribe the golden vectors of TimeUnit.Next for the fixed-duration units (hours and single days) from the
reference's own test table (pkg/util/timeunit_test.go: nextTestCases, :162-228) into tests/golden/timeunit_vectors.json.
Only the numbers travel: unit name, step in seconds, input and expected output as Unix seconds.  Multi-day, week, month,
quarter and year units truncate on the CALENDAR (timeunit.go:200-231); their window edges are computed by the Go side
(gpu.WindowEdges) and are not restated by the oracle, so they are not transcribed.

Usage: python tests/golden/extract_timeunit_vectors.py [/root/reference]
"""
import calendar
import json
import os
import re
import sys
import time

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "pkg/util/timeunit_test.go")).read()
block = src[src.index("var nextTestCases"):src.index("func TestNext")]
STEP = {"h": 3600, "2h": 7200, "3h": 10800, "d": 86400}


def unix(s):
    return calendar.timegm(time.strptime(s, "%Y-%m-%dT%H:%M:%SZ"))


out = []
for unit, body in re.findall(r'"(\w+)":\s*\{(.*?)\n\t\},', block, re.S):
    if unit not in STEP:
        continue
    for a, b in re.findall(r'\{tm\("([^"]+)"\),\s*tm\("([^"]+)"\)\}', body):
        out.append({"unit": unit, "step_s": STEP[unit], "in": unix(a), "next": unix(b), "in_text": a, "next_text": b})
assert len(out) >= 15, len(out)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "timeunit_vectors.json")
json.dump({"source": "pkg/util/timeunit_test.go:162-228 (nextTestCases)", "cases": out}, open(dst, "w"), indent=1)
print(len(out), "cases ->", dst)
