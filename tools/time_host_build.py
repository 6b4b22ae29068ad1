"""Times BKE_pbvh_build_mesh of the host library on a 4.2 M-vertex grid (DUNE_PBVH_TIMING=1 prints the two passes).
Usage: OMP_NUM_THREADS=N python tools/time_host_build.py"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dune_sculpt_b200 import capi, meshgen
from dune_sculpt_b200.capi import *
H = capi.host_lib()
mesh = meshgen.grid(2048)
mvert = np.zeros(mesh.totvert, dtype=MVERT); mvert["co"]=mesh.co
mpoly = np.zeros(mesh.totpoly, dtype=MPOLY); mpoly["loopstart"]=mesh.poly_start; mpoly["totloop"]=mesh.poly_len
mloop = np.zeros(mesh.totloop, dtype=MLOOP); mloop["v"]=mesh.loop_v.astype(np.uint32)
tottri=int(H.BKE_mesh_poly_to_tri_count(mesh.totpoly, mesh.totloop))
looptri=np.zeros(tottri, dtype=MLOOPTRI)
H.BKE_mesh_recalc_looptri(mloop.ctypes.data, mpoly.ctypes.data, mvert.ctypes.data, mesh.totloop, mesh.totpoly, looptri.ctypes.data)
t=time.time()
pb=H.BKE_pbvh_new(); H.DUNE_pbvh_mesh_sizes_set(pb, mesh.totpoly, mesh.totloop)
H.BKE_pbvh_build_mesh(pb, None, mpoly.ctypes.data, mloop.ctypes.data, mvert.ctypes.data, mesh.totvert, None,None,None, looptri.ctypes.data, tottri)
print("build_mesh %.2f"%(time.time()-t))
