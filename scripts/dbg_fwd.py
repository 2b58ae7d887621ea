import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from cases import load_golden, rel_l2, golden_names
from difflexmm_b200 import _abi, _lib
for name in golden_names():
    c = load_golden(name)
    topo = _lib.Topology(c.spec, 0)
    leaves = {k: torch.as_tensor(np.asarray(v), dtype=torch.float64, device="cuda").contiguous() for k, v in c.leaves.items()}
    ps = _abi.ParamSet(c.spec, 1, leaves, c.per_bond, c.damping_per_dof)
    ys, st = _lib.forward(topo, ps, torch.as_tensor(c.y0, device="cuda"), torch.as_tensor(c.ts, device="cuda"), c.rtol, c.atol, _abi.DfxOptions(0,0,0))
    s = st.numpy()[0]
    y = ys[0].cpu().numpy()
    print(name, 'relL2', rel_l2(y, c.ref['ys']), 'steps', s['steps'], s['accepted'], 'ref', c.ref['fwd_steps'], c.ref['fwd_accepted'], 'status', s['status'], 'nan', np.isnan(y).sum())
