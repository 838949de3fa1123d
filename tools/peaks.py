"""Print the in-tree measured peaks (gdk_measure_peaks) as one JSON line: python tools/peaks.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from getdist_b200 import _abi  # noqa: E402

print(json.dumps(_abi.Context(0).measure_peaks()))
