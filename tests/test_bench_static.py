"""Static checks of bench.py that need no GPU: the multi-rank control flow must not issue a collective after the non-zero ranks
have returned (that deadlocks rank 0 until the NCCL timeout), and the JSON contract keys must all be produced."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_collective_after_nonzero_ranks_leave():
    s = open(os.path.join(ROOT, "bench.py")).read()
    marker = "        dist.destroy_process_group()\n        if rank != 0:\n            ctx.close()\n            return"
    assert marker in s
    tail = s[s.index(marker) + len(marker):]
    tail = tail[:tail.index("\ndef ")]  # the rest of main(); later top-level functions have their own single-rank helpers
    assert not re.findall(r"\btotal\(|all_reduce|\bbarrier\(|\btimed\(", tail)


def test_contract_keys_present():
    s = open(os.path.join(ROOT, "bench.py")).read()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "h2d_bytes_per_step", "d2h_bytes_per_step",
                "impl", "bound", "achieved", "peak", "frac", "traffic", "cores", "kind", "sample"):
        assert '"%s"' % key in s, key
