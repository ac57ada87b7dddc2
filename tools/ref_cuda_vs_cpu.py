import ctypes, json, os, sys
import numpy as np
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'oracle'))
import __graft_entry__ as ge, refcheck
pkg=ge.load_package()
pack=os.path.join(ROOT,'scenes','cornell-box.b200scene'); w=h=64; spp=16
L=ctypes.CDLL(os.path.join(ROOT,'oracle','_ref','libcsrt_ref_cuda.so'))
L.ref_create_cuda.restype=ctypes.c_void_p
L.ref_create_cuda.argtypes=[ctypes.c_void_p,ctypes.c_int,ctypes.c_int,ctypes.c_int,ctypes.POINTER(ctypes.c_double)]
L.ref_draw_cuda.restype=ctypes.c_double; L.ref_draw_cuda.argtypes=[ctypes.c_void_p,ctypes.c_void_p]
scene=pkg.Scene(pack)
hnd=L.ref_create_cuda(scene.desc,w,h,spp,None)
g=np.zeros((h,w,3),np.float32); L.ref_draw_cuda(hnd,g.ctypes.data)
c,_,_=refcheck.ref_lib('woop').render_pack(pack,w,h,spp)
rel=lambda a,b: float(np.linalg.norm(a.astype(np.float64)-b)/np.linalg.norm(b.astype(np.float64)))
print(json.dumps({"still_ref_cuda_vs_ref_cpu":{"identical_pixels":float((g==c).all(axis=2).mean()),"rel_l2":rel(g,c),"max_abs":float(np.abs(g-c).max()),"mean_ratio":float(g.mean()/c.mean())}}))
