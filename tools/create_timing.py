"""b200pt_create timing on the GPU box: three creations of the Dragon scene in one process, phases printed (B200PT_VERBOSE_CREATE)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200PT_VERBOSE_CREATE"] = "1"
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
import torch  # noqa: E402

torch.zeros(1, device="cuda")
torch.cuda.synchronize()
name = sys.argv[1] if len(sys.argv) > 1 else "dragon"
t0 = time.time()
scene = pkg.Scene(os.path.join(ROOT, "scenes", name + ".b200scene"))
print(f"pack load {time.time() - t0:.3f} s", flush=True)
for k in range(3):
    t0 = time.time()
    r = pkg.Renderer(scene, device=0)
    print(f"== create #{k}: {time.time() - t0:.3f} s", flush=True)
    r.close()
print("nproc", os.cpu_count(), "loadavg", os.getloadavg())
