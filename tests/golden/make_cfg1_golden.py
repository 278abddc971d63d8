"""Golden vectors of config 1 (BASELINE configs[0]: RegionalTemporalGCN, dataloading_type=2, real TPIMS graph, 12-step
window, synthetic occupancy; SURVEY 8d: B=1, N=104, R=5, H=256, O=6, seed 101), produced by the fp64 oracle:
    python tests/golden/make_cfg1_golden.py            # writes tests/golden/cfg1_oracle_fp64.npz
Inputs are regenerated from the seed by regt_b200.workloads (graph from tests/golden/tpims_links.npz), so the file holds only
the oracle's outputs: out, out_hidden, loss, and for every live parameter gradient its max-abs, its sum and its first 32
entries.  tests/test_oracle.py checks that the oracle still reproduces it (drift guard); tests/test_gpu_model_parity.py checks
the CUDA path against it (1e-5)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from parity_util import W, is_dead, oracle_step  # noqa: E402


def main():
    w = W.make_workload(1)
    ref = oracle_step(w, 1)
    out = {"out": ref["out"].numpy(), "hid": ref["hid"].numpy(), "loss": np.float64(ref["loss"])}
    for k, g in ref["grads"].items():
        if g is None or is_dead(w.model, k):
            continue
        g = g.double().reshape(-1).numpy()
        out["gmax:" + k] = np.float64(np.abs(g).max())
        out["gsum:" + k] = np.float64(g.sum())
        out["ghead:" + k] = g[:32].copy()
    np.savez_compressed(os.path.join(HERE, "cfg1_oracle_fp64.npz"), **out)
    print("wrote", len(out), "arrays; loss", float(ref["loss"]))


if __name__ == "__main__":
    main()
