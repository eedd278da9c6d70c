import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from oracle import oracle as O
import msm_we_b200.clustering_ops as co
from msm_we_b200 import ops
orig_assign = ops.assign_stratified
orig_update = ops.minibatch_update
state = {}
def dbg_assign(X_all, bins, flag, centers, csq, offs, max_k, **kw):
    labels = orig_assign(X_all, bins, flag, centers, csq, offs, max_k, **kw)
    X = X_all.cpu().numpy(); b = bins.cpu().numpy(); c = centers.cpu().numpy(); o = offs.cpu().numpy()
    lab = labels.cpu().numpy()
    for bb in np.unique(b):
        sel = np.where(b == bb)[0]
        ref, m = O.kmeans_assign(X[sel], c[o[bb]:o[bb+1]], return_margin=True)
        bad = np.where(ref + o[bb] != lab[sel])[0]
        if len(bad):
            print("LABEL MISMATCH bin-slot", bb, "n", len(bad), "margins", m[bad][:5], "ref", ref[bad][:5], "got", (lab[sel]-o[bb])[bad][:5])
    state['X']=X; state['lab']=lab; state['c']=c.copy()
    return labels
def dbg_update(X_all, w_all, labels, centers, counts):
    c0 = centers.cpu().numpy().copy(); n0 = counts.cpu().numpy().copy()
    orig_update(X_all, w_all, labels, centers, counts)
    O.minibatch_update(X_all.cpu().numpy(), w_all.cpu().numpy(), c0, n0, labels.cpu().numpy())
    d = np.abs(c0 - centers.cpu().numpy()).max(axis=1)
    bad = np.where(d > 0)[0]
    if len(bad):
        print("UPDATE MISMATCH clusters", bad[:10], d[bad][:10], "counts diff", np.abs(n0-counts.cpu().numpy())[bad][:10], "n0", n0[bad][:10])
        lab = labels.cpu().numpy(); w = w_all.cpu().numpy()
        for k in bad[:3]:
            idx = np.where(lab == k)[0]
            print("  cluster", k, "members", len(idx), "w", w[idx][:8], "wsum", w[idx].sum())
ops.assign_stratified = dbg_assign
ops.minibatch_update = dbg_update
import test_model_gpu as T
cfg, model, mapper, its, _, basis, target = T._build("tiny", True)
model.launch_ray_discretization = lambda *a, **k: None
model.cluster_coordinates(cfg.k_per_bin, stratified=True, use_ray=True, user_bin_mapper=mapper, random_state=1337)
print("done")
