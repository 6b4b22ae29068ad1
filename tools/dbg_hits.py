import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from dune_sculpt_b200 import meshgen, stroke, capi
m = meshgen.grid(int(sys.argv[1]))
diag = m.bbox_diag()
dabs = stroke.c3_radius_sweep(diag, dabs_per_radius=2)
ses = capi.SculptSession(m, device=0)
ses.stroke_begin()
prev=0
for d in dabs:
    ses.dab(d)
    st=ses.stats()
    print(round(d.radius/diag*100,1), len(ses.hits()), st["vertex_dabs"]-prev, st)
    prev=st["vertex_dabs"]
ses.stroke_end()
