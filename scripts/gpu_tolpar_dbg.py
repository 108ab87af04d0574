import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nxblink_b200 as nb
from nxblink_b200 import p53, packing
m = p53.p53_model(); pm = packing.PackedModel(m)
ls, tt, tf = p53.synthetic_alignment(pm, 2048, seed=0, reference_leaf=pm.node_index[m.name_to_leaf['Has']])
d = packing.pack_alignment(pm, ls, tt, tf)
s = nb.BlinkSampler(pm, validate=True)
s.set_data(d); s.init_chains(chains=16, seed=7)
U = int(sys.argv[1]) if len(sys.argv) > 1 else 11176
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 7
s.sweep(NS, accumulate=False)
a = s.export_unit(U)
try:
    s.sweep(1, accumulate=False)
    print('no error')
except Exception as ex:
    print('ERR', ex)
b = s.export_unit(U)
cls = np.array(pm.state_class) if hasattr(pm, 'state_class') else None
print('attrs', [x for x in dir(pm) if not x.startswith('_')])
np.savez('gpurun_out/dbg_unit.npz', a_node_pri=a['node_pri'], a_node_tol=a['node_tol'], a_pt=a['pri'][0], a_pe=a['pri'][1], a_psa=a['pri'][2], a_psb=a['pri'][3],
    a_bt=a['blk'][0], a_be=a['blk'][1], a_bk=a['blk'][2], a_bs=a['blk'][3],
    b_node_pri=b['node_pri'], b_node_tol=b['node_tol'], b_pt=b['pri'][0], b_pe=b['pri'][1], b_psa=b['pri'][2], b_psb=b['pri'][3],
    b_bt=b['blk'][0], b_be=b['blk'][1], b_bk=b['blk'][2], b_bs=b['blk'][3])
