"""builds experiment variants of libregt_b200.so (extra -D switches) under regt-gcn_b200/lib/variants/."""
import concurrent.futures as cf
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("regt_build", os.path.join(ROOT, "regt-gcn_b200", "csrc", "build.py"))
B = importlib.util.module_from_spec(spec)
spec.loader.exec_module(B)

NOPIPE = ["REGT_FWD_PIPE=0", "REGT_BWD_H_UNDER_M2=0", "REGT_BWD_EARLY_E0=0"]
VARIANTS = {
    "base": [],
    "nopipe": NOPIPE,
    "nopipe_thrarrive": NOPIPE + ["REGT_WARP_ARRIVE=0"],
    "fwdpipe_only": ["REGT_BWD_H_UNDER_M2=0", "REGT_BWD_EARLY_E0=0"],
    "bwd_early_only": ["REGT_FWD_PIPE=0", "REGT_BWD_H_UNDER_M2=0"],
    "cw16": ["REGT_CW=16"],
    "cw16_nopipe": ["REGT_CW=16"] + NOPIPE,
    # round 2, fused 3xTF32 cell (cell_f.cu)
    "fpre0": ["REGT_F_PRE=0"],     # h of the next step computed after the candidate MMAs (default: under them)
    "xp_noh16": ["REGT_XP_SKIP_H16=1"],              # timing experiments (wrong results)
    "xp_nostore": ["REGT_XP_SKIP_E0_STORES=1"],
    "tn2": ["REGT_TN_GROUP=2"],      # row contraction: drain the TMEM accumulator every 2 chunks (24 MMAs) instead of 8
    "tn4": ["REGT_TN_GROUP=4"],
    "prefetch1": ["REGT_F_PREFETCH=1"],    # backward: L2 prefetch of the next step's planes a whole step ahead (round-2 default until call X)
    "prefetch0": ["REGT_F_PREFETCH=0"],    # backward: no L2 prefetch
    "fns3": ["REGT_F_NS_CAP=3"],           # forward: weight ring of 3 stages instead of 5 (what an extra 48 KB of operands would leave)
    "fns4": ["REGT_F_NS_CAP=4"],
    "stcs0": ["REGT_F_STCS=0"],            # backward: plain plane stores
    "ldcs0": ["REGT_F_LDCS=0"],            # backward: Z / H~ with ld.global.nc instead of the streaming operator
    "cache_plain": ["REGT_F_STCS=0", "REGT_F_LDCS=0"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    B.build()
    with cf.ThreadPoolExecutor(4) as ex:
        for n, p in zip(names, ex.map(lambda n: B.build_variant(n, VARIANTS[n]), names)):
            print(n, p)
