import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, dataclasses, math
import ebn_b200
from ebn_b200 import expr, _lib as L
from oracle import cpu as ocpu
import formula_fuzz
INF = math.inf
ebn_b200.init()
SEED, TRIAL = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(77001 + SEED)
names = ["x0", "x1", "x2", "x3"]
for trial in range(TRIAL + 1):
    models, perf, consts = formula_fuzz.chain(rng, names)
prog = expr.compile_program(names, models, perf, consts)
rows = [[(L.IN_NORMAL, 0.3, 1.5, 0, 0, -INF, INF), (L.IN_NORMAL, -0.2, 1.0, 0, 0, -INF, INF),
         (L.IN_UNIFORM, -1.0, 2.0, 0, 0, -INF, INF), (L.IN_NORMAL, 1.0, 0.7, 0, 0, -INF, INF)]]
n = 20
job = L.Job(L.LS_EXPRESSION, prog.consts, L.make_inputs(rows), n, seed=0x51C0 + TRIAL, strict=False)
X = ebn_b200.sample_matrix(job, 0, 0, n)
cnt, g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, X, want_g=True)
print("oracle g", np.round(g, 4))
Xs, gs = ebn_b200.sample_matrix(dataclasses.replace(job, strict=True), 0, 0, n, want_g=True)
Xf, gf = ebn_b200.sample_matrix(job, 0, 0, n, want_g=True)
print("gpu strict g", np.round(gs, 4)); print("gpu fast g", np.round(gf, 4))
per = [int(ebn_b200.pf_batch(dataclasses.replace(job, n=1, sample_offset=i))[0][0]) for i in range(n)]
pers = [int(ebn_b200.pf_batch(dataclasses.replace(job, n=1, sample_offset=i, strict=True))[0][0]) for i in range(n)]
print("fast fails  ", per); print("strict fails", pers); print("oracle      ", (g < 0).astype(int).tolist())
np.set_printoptions(precision=17, linewidth=200)
print("X rows", repr(X[:12]))
