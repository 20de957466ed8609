"""CPU: how far the REFERENCE is from ITSELF.  north_star asks for <= 1e-3 max-abs on logits against the reference CPU backend.
The reference's own AVX2 (x86-64-v3) and AVX-512 (x86-64-v4) builds of ggml-quants.c / ggml.c sum the same integer dot products in
different fp32 orders; one ulp of difference in an activation flips `round(x * 127 / amax)` of a q8_K code now and then, the flip is
a ~3e-4 relative kick to one mat-vec output and it stays in the KV cache.  This test measures that divergence on the model the GPU
parity tests use (tests/test_gpu_engine.py::test_engine_vs_port_longer_decode), so that their multi-token bar (first token to fp32
order, NMSE <= 2e-3, max-abs < 0.25, >= 90 % greedy agreement) is evidence-based: it is the bar the reference meets against itself."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O

ROOT = Path(__file__).resolve().parent.parent


def decode(variant, arch, n_tok, tmp_path, branch=0.1):
    out = tmp_path / f"{variant}_{arch}.npy"
    p = subprocess.run([sys.executable, str(ROOT / "tests" / "ref_variant_decode.py"), arch, str(n_tok), str(out), str(branch)],
                       env=dict(os.environ, PB200_REF_VARIANT=variant), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip().splitlines()[-1] == variant
    return np.load(out)


@pytest.mark.parametrize("arch", ["llama", "qwen2"])
def test_avx2_and_avx512_builds_of_the_reference_diverge_beyond_1e_3(arch, tmp_path):
    if not ((O.ORACLE / "_ref" / "v3" / "libgraph_ref.so").exists() and (O.ORACLE / "_ref" / "v4" / "libgraph_ref.so").exists()):
        pytest.skip("oracle/_ref v3 + v4 not built")
    if not {"avx512f", "avx512bw", "avx512vl"} <= O._cpu_flags():
        pytest.skip("this CPU has no AVX-512: the v4 build cannot run")
    a, b = decode("v3", arch, 40, tmp_path), decode("v4", arch, 40, tmp_path)
    e = np.max(np.abs(a - b), axis=1)
    nmse = float(np.sum((a - b) ** 2) / np.sum(a ** 2))
    print(f"{arch}: reference AVX2 vs AVX-512, per-token max-abs {np.array2string(e, precision=4)}; NMSE {nmse:.3e}")
    assert e[0] < 1e-4                      # same integers, only the fp32 order differs
    assert e.max() > 1e-3                   # ... yet over a few dozen tokens the reference leaves its own 1e-3 neighbourhood
    # and it stays inside the statistical bar the GPU tests use
    assert nmse < 2e-3 and e.max() < 0.25 and np.mean(a.argmax(1) == b.argmax(1)) >= 0.9
