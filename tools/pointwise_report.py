"""How often, and where, does a device BSDF sample differ from the reference's at the same inputs and the same LCG seed?
(tests/test_gpu_pointwise.py asserts <= 0.2 % of the probes; this prints the actual shares and the worst probes per BSDF.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import test_gpu_pointwise as tp  # noqa: E402

pkg = ge.load_package()
n = 200000
for scene_name in tp.SCENES:
    try:
        scene, ours, ref = tp.setup(pkg, scene_name)
    except BaseException as e:  # pytest.skip
        print(scene_name, "skipped:", e)
        continue
    _, types = tp.counts(scene)
    rng = np.random.RandomState(5)
    for index, kind in enumerate(types):
        if kind not in tp.BSDF_NAMES:
            continue
        inp = tp.bsdf_inputs(n, rng, both_sides=kind in (5, 6))
        a, b = ours.debug_eval(pkg.EVAL_BSDF_SAMPLE, index, inp), ref.eval(pkg.EVAL_BSDF_SAMPLE, index, inp)
        state = a[:, 15].view(np.uint32) != b[:, 15].view(np.uint32)
        flag = a[:, 0] != b[:, 0]
        both = (~flag) & (b[:, 0] != 0)
        scale = np.maximum(np.abs(b[:, 1:5]).max(axis=1), 1e-3)
        err = np.abs(a[:, 1:5].astype(np.float64) - b[:, 1:5]).max(axis=1) / scale
        derr = np.abs(a[:, 5:8].astype(np.float64) - b[:, 5:8]).max(axis=1)
        val = both & (err > 2e-4)
        dirs = both & (derr > 2e-4)
        print(f"{scene_name:45s} #{index} {tp.BSDF_NAMES[kind]:15s} probes {n}: draws differ {state.sum():5d}  valid flag differs {flag.sum():5d}  "
              f"value off>2e-4 {val.sum():5d}  direction off>2e-4 {dirs.sum():5d}  (valid {int((b[:, 0] != 0).sum())})", flush=True)
        for k in np.nonzero(state | flag | dirs)[0][:3]:
            print("    probe wi", inp[k, 0:3], "wo", inp[k, 3:6], "n", inp[k, 6:9], "inside", inp[k, 17], "seed %08x" % inp[k, 18:19].view(np.uint32)[0])
            print("      ours valid %g pdf %.7g att %s wi %s state %08x" % (a[k, 0], a[k, 1], a[k, 2:5], a[k, 5:8], a[k, 15:16].view(np.uint32)[0]))
            print("      ref  valid %g pdf %.7g att %s wi %s state %08x" % (b[k, 0], b[k, 1], b[k, 2:5], b[k, 5:8], b[k, 15:16].view(np.uint32)[0]))
