"""First-light diagnostics on a B200: every conv case reported (not asserted) with the location pattern of
mismatches, so one gpurun call is enough to localise descriptor / swizzle / pipeline bugs."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import gpu_util as U                    # noqa: E402
from tests.test_gpu_conv import CASES, make_case, run_case   # noqa: E402


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    only = sys.argv[1:] 
    for name, B, H, W, cin, cout, k, s, opts in CASES:
        if only and not any(o in name for o in only):
            continue
        for simt in (False, True):
            try:
                case = make_case(B, H, W, cin, cout, k, s, opts)
                got, want, _ = run_case(case, simt=simt)
                f32 = bool(opts.get("f32"))
                rep = U.error_report(got, want, name, 2e-3 if f32 else 2e-2, 2e-3 if f32 else 2e-2)
                print(("SIMT " if simt else "TC   ") + ("OK   " if rep["bad_frac"] == 0 and rep["nan"] == 0 else "FAIL ") + str(rep), flush=True)
            except Exception as e:  # keep going: later cases still tell us something unless the context died
                print(("SIMT " if simt else "TC   ") + f"EXC  {name}: {e}", flush=True)
                traceback.print_exc()
                try:
                    torch.cuda.synchronize()
                except Exception as e2:
                    print("context is dead:", e2, flush=True)
                    return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
