"""GPU: the tcgen05 forward must give the same bits on the first launch of a fresh process as on the tenth, and on
every launch that starts from poisoned tensor memory / shared memory / output buffers with a different weight set
than its predecessor (round 1 saw a one-off first-launch deviation; DESIGN.md section 6 has the analysis).
One fresh interpreter per tile configuration: <16,2>, <32,2>, <16,4>."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = {
    "ns16_nt2_save": (dict(FGRNN_TC_NS="16", FGRNN_TC_NT="2"), 64, 1, 1),
    "ns32_nt2": (dict(FGRNN_TC_NS="32", FGRNN_TC_NT="2"), 1024, 0, 2),
    "ns16_nt4": (dict(FGRNN_TC_NT="4"), 1024, 0, 3),
    "ns32_nt2_save": (dict(FGRNN_TC_NS="32", FGRNN_TC_NT="2"), 200, 1, 4),
    "ns32_nt2_split_epilogue": (dict(FGRNN_TC_NS="32", FGRNN_TC_NT="2", FGRNN_TC_ALT="0"), 1024, 0, 5),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_first_launch_and_poisoned_launches_are_bit_identical(name):
    env_extra, B, save, seed = CONFIGS[name]
    env = dict(os.environ)
    for k in ("FGRNN_TC_NS", "FGRNN_TC_NT", "FGRNN_TC_ALT"):
        env.pop(k, None)
    env.update(env_extra)
    res = subprocess.run([sys.executable, os.path.join(HERE, "first_launch_child.py"), str(B), str(save), str(seed), "24"],
                         env=env, capture_output=True, text=True, timeout=600)
    tail = (res.stdout + res.stderr)[-2000:]
    assert res.returncode == 0 and "RESULT ok" in res.stdout, tail
