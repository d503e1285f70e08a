"""GPU tier: the multi-config tool (tools/configs.py) runs at reduced scale in one process."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configs_tool_small_scale():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "configs.py"), "c3", "c4", "c5", "--scale", "0.004", "--reps", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert [l["config"] for l in lines] == ["c3", "c4", "c5"]
    assert lines[0]["inliers"] > 0 and lines[1]["pairs"] == 16 and lines[2]["tri_points_per_s"] > 0
