"""The library's opt-in / opt-out code paths (read from CFFT_B200_* variables once per process) give the same bits: each
setting runs tests/env_variant_check.py in a fresh process against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

SETTINGS = [
    {"CFFT_B200_FAST_PREFETCH": "0", "CFFT_B200_COLUMN_PREFETCH": "0", "CFFT_B200_TWOPASS_PREFETCH": "0"},
    {"CFFT_B200_FAST_PREFETCH": "2", "CFFT_B200_COLUMN_PREFETCH": "2"},
    {"CFFT_B200_FAST_PREFETCH": "1", "CFFT_B200_FUSED_MUL_PREFETCH": "0"},
    {"CFFT_B200_COLPIPE": "1"},
    {"CFFT_B200_TMEM_COLUMNS": "1"},
    {"CFFT_B200_ROWS_STD_ONE_EXCHANGE": "1"},
    {},
]


@pytest.mark.parametrize("setting", SETTINGS, ids=lambda s: ",".join("%s=%s" % (k[10:], v) for k, v in s.items()) or "defaults")
def test_env_variant_bit_exact(setting):
    env = dict(os.environ, **setting)
    r = subprocess.run([sys.executable, os.path.join(HERE, "env_variant_check.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
