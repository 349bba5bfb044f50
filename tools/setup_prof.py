import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import edgefem_b200
from edgefem_b200 import meshgen, cabi
pe = edgefem_b200.load_pyedgefem()
n=int(sys.argv[1])
T=time.perf_counter
t=T(); xyz,tets,tp,tris,trp = meshgen.cube_cavity(n, jitter=0.1); t1=T(); print("meshgen", round(t1-t,3))
hm = pe.mesh_from_arrays(xyz,tets,tp,tris,trp); t2=T(); print("mesh_from_arrays+build_edges", round(t2-t1,3))
bc = pe.build_edge_pec(hm,1); t3=T(); print("build_edge_pec", round(t3-t2,3))
a = [hm.xyz_array(), hm.tet_nodes_array(), hm.tet_edges_array(), hm.tet_orient_array(), hm.tet_phys_array(), hm.edge_nodes_array()]; t4=T(); print("arrays", round(t4-t3,3))
flags = np.zeros(hm.num_edges(), dtype=np.uint8); flags[np.asarray(bc.dirichlet_edges, dtype=np.int64)] = 1
ctx = cabi.Ctx(0); t5=T()
dm = cabi.DeviceMesh(ctx, *a); t6=T(); print("DeviceMesh", round(t6-t5,3))
pe_idx = np.nonzero(flags)[0].astype(np.int32)
s = cabi.DeviceSystem.from_mesh(dm, pe_idx, pe_idx); t7=T(); print("system create", round(t7-t6,3))
s.set_dirichlet(flags); t8=T(); print("set_dirichlet", round(t8-t7,3))
print(hm.num_tets(), hm.num_edges(), s.nnz)
