"""CPU simulation of tests/test_gpu_x_boundary_and_sampler.py::test_option7_device_walk_sampler_scores_match_libc_walks:
the oracle plays the GPU (test infrastructure; output kept in profiles/r2_sampler_parity_cpu_simulation.log).
libc-walk runs = oracle.run(seed); device-sampler runs = the same epoch loop with counter-based walks."""
import sys, os, time, numpy as np, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
from oracle import oracle as O
import evalscores as E
rp, ci = O.load_mtx(ROOT + '/tests/golden/cora.mtx')
n = len(rp) - 1
labels = E.read_labels(ROOT + '/tests/golden/cora.nodes.labels', n)
dim, B, s, lr, IT = 64, 256, 5, 0.02, 1200
nb = (n + B - 1) // B
lut = O.build_lut()

def run_counter(seed):
    g = O.Rng(seed)
    X = O.init_embeddings(g, 7, n, dim).copy()
    for it in range(IT):
        w = O.walks_counter(seed, it, rp, ci)
        for b in range(nb):
            idx = O.draw_negatives(g, 7, 0, n, B, s, b)
            O.step(7, 0, rp, ci, X, b * B, min(n, (b + 1) * B), idx, s, lr, lut=lut, walks=w, threads=1)
    return X

def scores(X):
    lp = E.link_prediction(rp, ci, X, seeds=(1, 2))
    nc = E.node_classification(X, labels, seeds=tuple(range(5)))
    return lp, nc

res = {"libc": [], "dev": []}
for seed in (1, 2, 3, 4):
    t = time.time()
    X = O.run(7, 0, rp, ci, dim, IT, B, s, lr, seed=seed, threads=2)["X"]
    res["libc"].append(scores(X))
    Xc = run_counter(seed)
    res["dev"].append(scores(Xc))
    print(seed, round(time.time() - t, 1), res["libc"][-1][0], res["dev"][-1][0], flush=True)
json.dump(res, open('/tmp/sim_sampler.json', 'w'))
TOL = 0.005
def agree(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    d = abs(a.mean() - b.mean())
    print(what, 'diff %.4f se %.4f bound %.4f %s   libc %s dev %s' % (d, se, TOL + 3 * se, 'OK' if d <= TOL + 3 * se else 'FAIL', np.round(a, 4), np.round(b, 4)))
for k in ("accuracy", "f1_macro", "f1_micro", "auc"):
    agree([r[0][k] for r in res["libc"]], [r[0][k] for r in res["dev"]], ("lp", k))
for tf in res["libc"][0][1]:
    for k in ("f1_macro", "f1_micro"):
        agree([r[1][tf][k] for r in res["libc"]], [r[1][tf][k] for r in res["dev"]], ("nc", tf, k))
