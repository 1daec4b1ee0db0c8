"""Downstream scores of an embedding: seeded restatement of the reference's evaluation protocols
(/root/reference/performancescores/runlinkpredict.py:51-140 and runnodeclassclust.py:162-190,
254-308).  The reference scripts draw unseeded shuffles, so one run of them is only good to
+-0.03; here every random choice comes from an explicit seed and the scores
are averaged over several splits, which makes +-0.005 comparisons between two embeddings of the
same graph meaningful.  The maths per split is the reference's:

  link prediction    positives = every edge (u < v); per vertex twice as many distinct random
                     non-neighbours as it has positives; Hadamard features x_u * x_v; shuffle;
                     first half trains a LogisticRegression, second half is scored
                     (accuracy, F1-macro, F1-micro as the script prints, plus ROC-AUC of the
                     decision function, which north_star asks for)
  node classification  One-vs-Rest LogisticRegression(random_state=0) on a shuffled train
                     fraction (5..25 %), top-k label prediction (k = number of true labels),
                     F1-macro / F1-micro on the rest.

CLI:  python tools/evalscores.py graph.mtx emb.embd [labels]
"""
import sys
import numpy as np


def read_embd(path):
    """.embd text (sample/algorithms.h:118-136): 'N D' then 'id v1 .. vD ' with 1-based ids."""
    with open(path) as f:
        n, d = [int(t) for t in f.readline().split()[:2]]
        X = np.zeros((n, d), np.float64)
        for line in f:
            tok = line.split()
            if tok:
                X[int(tok[0]) - 1] = [float(t) for t in tok[1:1 + d]]
    return X


def read_labels(path, n):
    """'node label' per line, node 1-based (runnodeclassclust.py:173-190).  Returns list of label lists."""
    labels = [[] for _ in range(n)]
    with open(path) as f:
        for line in f:
            tok = line.split()
            if len(tok) >= 2:
                labels[int(tok[0]) - 1].append(int(tok[1]))
    return labels


def link_prediction_data(rowptr, colids, X, rng):
    n = len(rowptr) - 1
    pu, pv, nu, nv = [], [], [], []
    for u in range(n):
        nb = np.unique(colids[rowptr[u]:rowptr[u + 1]])
        pos = nb[nb > u]
        pu.extend([u] * len(pos))
        pv.extend(pos.tolist())
        want = 2 * len(pos)
        if len(nb) > n // 2:
            want = (n - len(nb)) // 2
        if want == 0:
            continue
        taken, banned = [], set(nb.tolist())
        while len(taken) < want:
            c = int(rng.integers(0, n))
            if c not in banned:
                banned.add(c)
                taken.append(c)
        nu.extend([u] * want)
        nv.extend(taken)
    a = np.array(pu + nu, np.int64)
    b = np.array(pv + nv, np.int64)
    y = np.concatenate([np.ones(len(pu), np.int64), np.zeros(len(nu), np.int64)])
    F = X[a] * X[b]
    perm = rng.permutation(len(y))
    return F[perm], y[perm]


def link_prediction(rowptr, colids, X, seeds=(1, 2, 3)):
    from sklearn.linear_model import LogisticRegression
    from sklearn.metrics import accuracy_score, f1_score, roc_auc_score
    X = np.asarray(X, np.float64)
    out = {"accuracy": [], "f1_macro": [], "f1_micro": [], "auc": []}
    for seed in seeds:
        rng = np.random.default_rng(seed)
        F, y = link_prediction_data(rowptr, colids, X, rng)
        cv = len(y) // 2
        m = LogisticRegression().fit(F[:cv], y[:cv])
        pred = m.predict(F[cv:])
        out["accuracy"].append(accuracy_score(pred, y[cv:]))
        out["f1_macro"].append(f1_score(pred, y[cv:], average="macro", labels=np.unique(pred)))
        out["f1_micro"].append(f1_score(pred, y[cv:], average="micro", labels=np.unique(pred)))
        out["auc"].append(roc_auc_score(y[cv:], m.decision_function(F[cv:])))
    return {k: float(np.mean(v)) for k, v in out.items()}


def node_classification(X, labels, fractions=(0.05, 0.10, 0.15, 0.20, 0.25), seeds=tuple(range(20))):
    from sklearn.linear_model import LogisticRegression
    from sklearn.metrics import f1_score
    from sklearn.multiclass import OneVsRestClassifier
    from sklearn.preprocessing import MultiLabelBinarizer
    keep = [i for i, l in enumerate(labels) if l]
    Xd = np.asarray(X, np.float64)[keep]
    Yd = [labels[i] for i in keep]
    nlab = len({c for l in Yd for c in l})
    mlb = MultiLabelBinarizer(classes=list(range(nlab)))
    Yb = mlb.fit_transform(Yd)
    res = {}
    for tf in fractions:
        ma, mi = [], []
        for seed in seeds:
            idx = np.random.default_rng(1000 * seed + int(round(tf * 100))).permutation(len(Yd))
            cv = int(len(Yd) * tf)
            tr, te = idx[:cv], idx[cv:]
            clf = OneVsRestClassifier(LogisticRegression(random_state=0)).fit(Xd[tr], Yb[tr])
            ps = np.asarray(clf.predict_proba(Xd[te]))
            pred = np.zeros_like(Yb[te])
            for r, i in enumerate(te):
                k = len(Yd[i])
                pred[r, clf.classes_[np.argsort(ps[r])[-k:]]] = 1
            ma.append(f1_score(pred, Yb[te], average="macro"))
            mi.append(f1_score(pred, Yb[te], average="micro"))
        res[tf] = {"f1_macro": float(np.mean(ma)), "f1_micro": float(np.mean(mi))}
    return res


def main(argv):
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from force2vec_b200 import host
    rp, ci = host.load_mtx(argv[1])
    X = read_embd(argv[2])
    print("link prediction (Hadamard, 50/50):", link_prediction(rp, ci, X))
    if len(argv) > 3:
        for tf, r in node_classification(X, read_labels(argv[3], len(rp) - 1)).items():
            print("node classification %.2f:" % tf, r)


if __name__ == "__main__":
    main(sys.argv)
