"""Exact-mode debugging: where does a pixel of b200pt_debug_render_replay leave the reference's path?

  python tools/replay_trace.py SCENE W H SPP [N]      the N worst pixels: vertex-by-vertex traces of both sides, first divergence
  python tools/replay_trace.py SCENE W H SPP --pixel I,J --side gpu|cpu      one trace (used by the call above)

SCENE: a name under scenes/ or tests/golden/synthetic_*.  The CPU side is the C restatement (oracle/pt_oracle.c, bit-equal to
the reference build on these scenes) with ORACLE_TRACE_PIXEL; the GPU side is B200PT_REPLAY_TRACE (debug_eval.cu)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def scene_path(name):
    for p in (os.path.join(ROOT, "scenes", name + ".b200scene"), os.path.join(ROOT, "tests", "golden", name + ".b200scene")):
        if os.path.exists(p):
            return p
    raise SystemExit(f"no pack for {name}")


def gpu_frame(path, w, h, spp):
    import __graft_entry__ as ge
    pkg = ge.load_package()
    r = pkg.Renderer(pkg.Scene(path), device=0, max_paths_in_flight=1 << 20)
    f = r.render_replay(w, h, spp)
    r.close()
    return f


def cpu_frame(path, w, h, spp):
    import refcheck
    return refcheck.OracleLib().render_pack(path, w, h, spp)


def main():
    name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    path = scene_path(name)
    if "--pixel" in sys.argv:
        pixel = sys.argv[sys.argv.index("--pixel") + 1]
        side = sys.argv[sys.argv.index("--side") + 1]
        if side == "gpu":
            os.environ["B200PT_REPLAY_TRACE"] = pixel
            gpu_frame(path, w, h, spp)
        else:
            os.environ["ORACLE_TRACE_PIXEL"] = pixel
            cpu_frame(path, w, h, spp)
        return
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    a, b = gpu_frame(path, w, h, spp), cpu_frame(path, w, h, spp)
    d = np.abs(a.astype(np.float64) - b).max(axis=2) / np.maximum(np.abs(b).max(axis=2), 1e-3)
    print(f"## {name} {w}x{h}x{spp}: px<=1e-4 {np.mean(d <= 1e-4):.4f} <=1e-3 {np.mean(d <= 1e-3):.4f} <=1e-2 {np.mean(d <= 1e-2):.4f}; "
          f"median rel diff {np.median(d):.3e}")
    # the SMALLEST clear mismatches are the most telling (a single flipped decision), then the largest
    order = np.argsort(d.ravel())
    bad = [k for k in order if d.ravel()[k] > 1e-3]
    picks = bad[:n] + bad[-1:] if bad else []
    for k in picks:
        j, i = divmod(int(k), w)
        print(f"# pixel ({i},{j}) gpu {a[j, i]} cpu {b[j, i]} rel {d[j, i]:.3e}")
        out = {}
        for side in ("gpu", "cpu"):
            r = subprocess.run([sys.executable, __file__, name, str(w), str(h), str(spp), "--pixel", f"{i},{j}", "--side", side],
                               capture_output=True, text=True, timeout=600)
            out[side] = [l for l in (r.stdout + r.stderr).splitlines() if l.startswith("[trace]")]
        # first REAL divergence: a different LCG state, or a number off by more than 1e-4 relative (the last printed digits differ anyway)
        def parse(line):
            f = line.split()
            if len(f) < 4 or not f[1].startswith("v"):
                return None
            return f[1], f[3], [float(x) for x in f[5:6] + f[7:10] + f[11:14]]
        kind = "none within the printed vertices"
        for k2 in range(min(len(out["gpu"]), len(out["cpu"]))):
            g, c = parse(out["gpu"][k2]), parse(out["cpu"][k2])
            if g is None and c is None:
                continue
            if g is None or c is None or g[0] != c[0]:
                kind = "one side has a vertex the other has not (a path ended / escaped earlier)"
            elif g[1] != c[1]:
                kind = "LCG STATE differs (for volpath the GPU prints before, the CPU after the medium draw of the vertex)"
            else:
                rel = [abs(x - y) / max(abs(y), 1e-6) for x, y in zip(g[2], c[2])]
                if max(rel) <= 1e-4:
                    continue
                kind = ("hit distance" if rel[0] > 1e-4 else "radiance so far" if max(rel[1:4]) > 1e-4 else "throughput") + f" differs by {max(rel):.1e} at equal LCG state"
            print(f"  first divergence (line {k2}): {kind}")
            for k3 in range(max(0, k2 - 2), min(k2 + 2, len(out["gpu"]), len(out["cpu"]))):
                print("    gpu", out["gpu"][k3])
                print("    cpu", out["cpu"][k3])
            break
        else:
            print("  first divergence:", kind)

if __name__ == "__main__":
    main()
