"""Static evidence without a GPU: what nvcc put into voicesplit_b200/libvoicesplit_sm100.so.  Every embedded ELF is sm_100a, and the
hot kernels hold the SASS that proves the B200-native path (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,
UTMALDG = TMA tensor load, SYNCS = mbarrier) - i.e. the library that the GPU tests load is not a recompiled mma.sync / CUDA-core
fallback.  The dominant conv kernel must not spill (no stack frame)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "voicesplit_b200", "libvoicesplit_sm100.so")
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")


def _kernel(name):
    m = re.search(r"_ZN2vs\d+([A-Za-z0-9_]+?)[IE]", name)
    return m.group(1) if m else name


def test_every_embedded_elf_is_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    elfs = re.findall(r"ELF file\s+\d+: (\S+)", out)
    assert len(elfs) >= 10 and all(e.endswith(".sm_100a.cubin") for e in elfs), elfs
    ptx = subprocess.run(["cuobjdump", "-lptx", LIB], capture_output=True, text=True).stdout
    assert "sm_90" not in ptx and "sm_80" not in ptx                      # no other architecture rides along


def test_hot_kernels_carry_tcgen05_and_tma_sass():
    usage, cur = {}, None
    for line in subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL)\s*:\s*(\d+)", line)}
            cur = None
    pats = {"mma": r"\bUTC[A-Z]*MMA", "f16mma": r"\bUTCHMMA", "f8mma": r"\bUTCQMMA", "tmem_ld": r"\bLDTM", "tma": r"\b(UTMALDG|UBLKCP)", "mbar": r"\bSYNCS"}
    counts, cur = collections.defaultdict(collections.Counter), None
    for line in subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
        elif cur:
            for k, p in pats.items():
                if re.search(p, line):
                    counts[cur][k] += 1
    by_kernel = collections.defaultdict(list)
    for fn in usage:
        by_kernel[_kernel(fn)].append(fn)
    for k in ("k_conv_tc", "k_gemm_tc", "k_lstm_tc", "k_lstm_uni_tc", "k_point8_mma", "k_wgrad_tc"):      # every tensor-core kernel of the path
        assert by_kernel[k], k
        for fn in by_kernel[k]:
            c = counts[fn]
            assert c["mma"] > 0 and c["tmem_ld"] > 0 and c["tma"] > 0 and c["mbar"] > 0, (k, dict(c))
    conv = by_kernel["k_conv_tc"]
    assert all(usage[fn].get("STACK", 0) == 0 and usage[fn]["REG"] <= 128 for fn in conv)               # no spills in the dominant kernel
    # the fixed-schedule 5x5 variants: 15 steps fully unrolled - 180 MMAs (fp16x3 / bf16x3: 12 per step) or 120 (fp16_f8c: 4 f16 + 4 e4m3)
    n_mma = sorted({counts[fn]["mma"] for fn in conv})
    assert 180 in n_mma and 120 in n_mma, n_mma
    f8c = [fn for fn in conv if counts[fn]["mma"] == 120]
    assert all(counts[fn]["f16mma"] == 60 and counts[fn]["f8mma"] == 60 for fn in f8c)
