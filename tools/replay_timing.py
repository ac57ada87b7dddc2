"""Wall time of b200pt_debug_render_replay (one thread per pixel, the reference's loop shape around the product's device functions)
on the C2 workload — the megakernel datapoint quoted in DESIGN.md §8."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
r = pkg.Renderer(pkg.Scene(os.path.join(ROOT, "scenes", "dragon.b200scene")), device=0, max_paths_in_flight=1 << 20)
for k in range(3):
    t0 = time.time()
    r.render_replay(1024, 1024, 256)
    dt = time.time() - t0
    print(f"replay #{k}: {dt * 1e3:.1f} ms  {1024 * 1024 * 256 / dt / 1e6:.0f} Msamples/s (includes the 12.6 MB copy to the host)", flush=True)
