"""Own bounds check in place of compute-sanitizer (closed on this GPU pool): with CFFM_GUARD=1 every
device allocation of the library has a 4 KB pattern band on either side.  tests/guard_worker.py drives every arithmetic and
shape class in a fresh process (the switch is read at the first allocation) and reports damaged bands."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.timeout(580)
def test_no_kernel_writes_outside_its_buffers(tmp_path):
    out = tmp_path / "guard.json"
    env = dict(os.environ, CFFM_GUARD="1")
    r = subprocess.run([sys.executable, os.path.join(HERE, "guard_worker.py"), str(out)], capture_output=True, text=True,
                       timeout=560, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.load(open(out))
    assert len(rep) >= 15
    bad = [x for x in rep if x["damaged"] != 0]
    assert not bad, bad
    # keep the evidence next to the other run outputs
    os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(os.path.dirname(HERE), "gpurun_out", "guard_report.json"), "w"), indent=1)


def test_guard_detects_an_overrun():
    """The checker itself: a write one element before / after a guarded allocation is reported (and only then)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cffm_b200 import _lib\n"
            "print('selftest', _lib.load().cffm_debug_guard_selftest())\n") % (os.path.dirname(HERE),)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=200, env=dict(os.environ, CFFM_GUARD="1"))
    assert r.returncode == 0 and "selftest 0" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=200,
                       env={k: v for k, v in os.environ.items() if k != "CFFM_GUARD"})
    assert "selftest -1" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]       # off by default: no cost in production
