#!/usr/bin/env python
"""Long parity runs that do not fit the test suite: the full horizons of cfg2, three cfg3 members and both tasks of cfg4
(0.01 and 0.08 strain, the real static ramps of 0.04 s and 0.32 s) against the C++ oracle at tight tolerances.  Writes one
JSON line per case to profiles/r02_parity_long.jsonl (minutes of host CPU time; the oracle jobs run on host threads).

  python tools/parity_long.py [--quick] [--cfg5]

--cfg5 runs one more case alone: the 100 x 100 lattice with active contact (BASELINE.json cfg5) over two drive periods, generic
kernels spread over all SMs, against the single-threaded C++ oracle (ten minutes of host time) -> r02_parity_cfg5.jsonl."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_tight as T  # noqa: E402


def cases(quick):
    from difflexmm_b200.problems import KagomeFocusing, QuadsFocusing, QuadsStaticTuning
    out = []
    P = KagomeFocusing() if not quick else KagomeFocusing(simulation_time=1 / 30.0, n_timepoints=34)
    out.append(("cfg2 kagome 20x12, full horizon (3 periods, n_t 200)", P, P.initial_design(), 1e-10, 1e-10))
    P = QuadsFocusing() if not quick else QuadsFocusing(simulation_time=1 / 30.0, n_timepoints=50)
    hs, vs = P.random_ensemble(3, noise=0.15, seed0=0)
    for m in range(3):
        out.append((f"cfg3 member {m}, full horizon (2 periods, n_t 200)", P, (hs[m], vs[m]), 1e-10, 1e-10))
    for strain in (0.01, 0.08):
        P = QuadsStaticTuning(compressive_strain=strain) if not quick else QuadsStaticTuning(
            compressive_strain=strain, simulation_time_dynamic=1 / 30.0, n_timepoints=20)
        out.append((f"cfg4 strain {strain}, real static ramp + full dynamic window (n_t 201)", P, P.initial_design(), 1e-8, 1e-8))
    return out


def cfg5_case():
    import math
    from difflexmm_b200.problems import QuadsFocusing
    P = QuadsFocusing(n1_blocks=100, n2_blocks=100, simulation_time=2 / 30.0, n_timepoints=9, target_shift=(2, 2),
                      min_angle=15 * math.pi / 180, cutoff_angle=45 * math.pi / 180)
    return [("cfg5 quads 100x100, contact active, 2 drive periods (n_t 9)", P, P.initial_design(), 1e-8, 1e-8)]


def main():
    quick = "--quick" in sys.argv
    cs = cfg5_case() if "--cfg5" in sys.argv else cases(quick)
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=len(cs)) as pool:
        refs = list(pool.map(lambda c: T.oracle_job(c[1], c[2], c[3], c[4]), cs))
    t_oracle = time.time() - t0
    path = os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "profiles", "r02_parity_cfg5.jsonl" if "--cfg5" in sys.argv else "r02_parity_long.jsonl")
    with open(path, "w") as f:
        for (name, P, design, rtol, atol), ref in zip(cs, refs):
            try:
                res = T.compare(name, P, ref, rtol, atol)
                res["ok"] = True
            except AssertionError as e:
                res = {"ok": False, "error": str(e)[:2000]}
            res.update(case=name, rtol=rtol, atol=atol, oracle_wall_s_all_cases=round(t_oracle, 1))
            line = json.dumps(res)
            print(line, flush=True)
            f.write(line + "\n")


if __name__ == "__main__":
    main()
