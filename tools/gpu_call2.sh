mkdir -p gpurun_out
python -m pytest tests/test_gpu_bounds.py -q -m gpu --tb=line 2>&1 | grep -v "^$" | cut -c1-400 > gpurun_out/t_bounds2.log
python - > gpurun_out/diag_bounds.log 2>&1 <<'PY'
import sys; sys.path[:0]=['.','tests']
import numpy as np, problems as P
from test_gpu_parity import make_pair, rel
from p2de_b200 import *
from p2de_b200.api import rhs
np.set_printoptions(linewidth=200, precision=6)
for name,b,N in [("cell",PositivityAndCellEntropyBound(),1),("cell",PositivityAndCellEntropyBound(),3),("relcell",PositivityAndRelaxedCellEntropyBound(beta=0.5),2),("tvd",TVDBound(),2)]:
    param, solver, st, orc, U0 = make_pair(P.wave2d(N=N, limiter=SubcellLimiter(bound=b)))
    tp=param.timestepping_param
    orc.rhs(tp.t0, tp.CFL*tp.dt0, 1); rhs(st, solver, None, TimeParam(t=tp.t0, dt=tp.CFL*tp.dt0, nstage=1))
    Lg, Lo = st.preallocation.L_local[0], orc.field("L_local")[0]
    N1D=N+1
    d = np.abs(Lg-Lo)
    print(name, N, "max dL", d.max(), "n>1e-10", (d>1e-10).sum(), "of", d.size, "rhsU rel", rel(st.preallocation.rhsU, orc.field("rhsU")))
    bad = np.argwhere(d>1e-10)[:12]
    for k,dd,i in bad: print("   k",k,"d",dd,"idx",i,"gpu",Lg[k,dd,i],"orc",Lo[k,dd,i])
    k0 = bad[0][0] if len(bad) else 0
    print(" gpu x", Lg[k0,0].reshape(N1D,N1D+1)); print(" orc x", Lo[k0,0].reshape(N1D,N1D+1))
    print(" gpu y", Lg[k0,1].reshape(N1D+1,N1D)); print(" orc y", Lo[k0,1].reshape(N1D+1,N1D))
PY
