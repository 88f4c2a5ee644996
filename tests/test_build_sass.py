"""The shipped library really is the Blackwell path: its SASS holds tcgen05 MMAs (UTCHMMA, incl. the cta_group::2 form),
TMEM loads / stores (LDTM / STTM), TMA tensor loads (UTMALDG) and bulk copies (UBLKCP) -- the mnemonics
/opt/skills/guides/B200_PROFILING.md names as the evidence.  Needs cuobjdump (part of the CUDA toolkit), no GPU."""
import collections
import re
import shutil
import subprocess

import pytest


def test_library_sass_contains_tcgen05_tmem_and_tma_instructions():
    from codes_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300)
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    if out.returncode != 0:
        pytest.skip("cuobjdump failed: " + out.stderr[:200])
    counts = collections.Counter(re.findall(r"\b(UTCHMMA(?:\.2CTA)?|LDTM|STTM|UTMALDG|UBLKCP|UTCBAR)\b", out.stdout))
    assert "sm_100a" in out.stdout
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR"):
        assert counts[mnemonic] > 0, (mnemonic, dict(counts))
