"""Static checks on the SASS of the built library (CPU-only: cuobjdump reads the sm_100a cubin without a GPU).

Guards two properties the measurements depend on (profiles/r01_gemm_mainloop_timeline.md):
  * the hot kernels really are tcgen05 / TMA kernels (UTCHMMA, UTMALDG in their SASS), and
  * the producer / MMA-issuer loops stay free of the ELECT + R2UR.BROADCAST + BRA.U.ANY "waterfall" the compiler puts
    in front of every TMA / MMA instruction when it cannot prove the operands warp-uniform (it cost the GEMM 40 % of
    its main-loop time before the whole-warp / elected-lane rewrite).
"""
import collections
import shutil
import subprocess
from pathlib import Path

import pytest

LIB = Path(__file__).resolve().parents[1] / "audioldm_with_lora_b200" / "libb200ldm.so"


@pytest.fixture(scope="module")
def sass_counts():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    if not LIB.exists():
        pytest.skip("library not built (python -m audioldm_with_lora_b200.build)")
    out = subprocess.run([cuobjdump, "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    name = None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            continue
        if name is None:
            continue
        for op in ("UTCHMMA", "UTMALDG", "UTMASTG", "BRA.U.ANY", "R2UR.BROADCAST"):
            if op in line:
                counts[name][op] += 1
    return counts


def _kernels(counts, needle):
    return {k: v for k, v in counts.items() if needle in k}


def test_hot_kernels_use_tcgen05_and_tma(sass_counts):
    for needle in ("conv_gemm_kernel", "attention_kernel", "attention_bwd_kernel", "lora_wgrad_tc_kernel"):
        ks = _kernels(sass_counts, needle)
        assert ks, f"no {needle} in the library"
        for name, c in ks.items():
            assert c["UTCHMMA"] > 0, f"{name}: no tcgen05.mma (UTCHMMA) in SASS"
            assert c["UTMALDG"] > 0, f"{name}: no TMA load (UTMALDG) in SASS"
    for name, c in _kernels(sass_counts, "conv_gemm_kernel").items():
        assert c["UTMASTG"] > 0, f"{name}: the epilogue should store through TMA"


def test_issue_loops_have_no_uniformity_waterfalls(sass_counts):
    # one waterfall loop is tolerated per kernel (the epilogue's once-per-chunk TMA store); the main-loop roles have none
    for needle in ("conv_gemm_kernel", "attention_kernel", "attention_bwd_kernel"):
        for name, c in _kernels(sass_counts, needle).items():
            assert c["BRA.U.ANY"] <= 2, f"{name}: {c['BRA.U.ANY']} waterfall loops (ELECT / R2UR.BROADCAST / BRA.U.ANY) in SASS"
