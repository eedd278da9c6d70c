"""Seeded synthetic WE data of the shapes BASELINE.json names.

A WE run is imitated just closely enough for the hot path to see realistic inputs:
  * 1-D progress coordinate on [0, nbins): rectilinear WE bins with float32 boundaries 0,1,...,nbins-1,+inf;
    basis = (0, 0.5), target = (nbins-0.5, +big) lie inside the first / last WE bin;
  * the parent of every segment is a random segment of the previous iteration: pcoord0 and the parent
    features ARE that segment's pcoord1 and child features (so a parent's label must equal the label
    its own iteration gave it as a child -- a free consistency property for tests);
  * child pcoord = parent pcoord + N(0, 0.7) reflected into the bin space;
  * features = one of the K "true" micro-state means of the WE bin the pcoord falls in + N(0, 1);
  * weights = exp(N(0, 3)) normalised to 1 per iteration (many decades, like real WE);
  * cluster centres = true means + N(0, 0.3): fixed, seeded, the same for oracle and GPU.
Everything is float64, C-order.  Generation uses numpy's PCG64 on the host (small configs, oracle
parity) or torch's Philox on the device (benchmark configs); the two streams are different data sets.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class WEConfig:
    name: str
    n_iters: int
    n_segs: int
    dim: int
    n_bins: int
    k_per_bin: int
    seed: int = 20261018

    @property
    def n_clusters(self):
        return self.n_bins * self.k_per_bin

    @property
    def frames(self):
        return self.n_iters * self.n_segs


CONFIGS = {
    # BASELINE.json configs[1..4]; configs[0] (bundled NTL9) is a fixture replay, see tests
    "cfg2": WEConfig("cfg2", 200, 1000, 64, 30, 20, 20261018),
    "cfg3": WEConfig("cfg3", 1000, 4000, 3000, 50, 50, 20261019),
    "cfg5": WEConfig("cfg5", 5000, 8000, 256, 100, 100, 20261021),
    # configs[3] (flux-matrix stress) has no feature data: see generate_cfg4_device
    # reduced variants that keep the per-frame shape (D, bins, K) but fit quick runs
    "cfg3s": WEConfig("cfg3s", 50, 4000, 3000, 50, 50, 20261019),
    "cfg5s": WEConfig("cfg5s", 250, 8000, 256, 100, 100, 20261021),
    "tiny": WEConfig("tiny", 12, 150, 13, 12, 25, 7),
}


def boundaries(cfg: WEConfig):
    b = np.arange(cfg.n_bins + 1, dtype=np.float32)
    b[-1] = np.inf
    return [b]


def region_bounds(cfg: WEConfig):
    basis = np.array([[0.0, 0.5]])
    target = np.array([[cfg.n_bins - 0.5, 1.0e6]])
    return basis, target


def true_means(cfg: WEConfig, rng):
    """[n_bins, K, D]: a per-bin offset plus K micro-state means around it."""
    bin_mean = rng.normal(0.0, 3.0, size=(cfg.n_bins, 1, cfg.dim))
    return bin_mean + rng.normal(0.0, 2.0, size=(cfg.n_bins, cfg.k_per_bin, cfg.dim))


def make_centers(cfg: WEConfig, seed_offset=0):
    rng = np.random.default_rng(cfg.seed + 1000 + seed_offset)
    means = true_means(cfg, rng)
    centers = means + rng.normal(0.0, 0.3, size=means.shape)
    return means, [np.ascontiguousarray(centers[b]) for b in range(cfg.n_bins)]


def generate_host(cfg: WEConfig, means=None):
    """Host arrays for small configs: list over iterations of dicts with pcoord0, pcoord1, weights,
    parent (features), child (features)."""
    rng = np.random.default_rng(cfg.seed)
    if means is None:
        means, _ = make_centers(cfg)
    hi = cfg.n_bins - 1e-3
    its = []
    prev_pc = rng.uniform(0.0, hi, size=cfg.n_segs)
    prev_x = _features(cfg, means, prev_pc, rng)
    for _ in range(cfg.n_iters):
        parent = rng.integers(0, cfg.n_segs, size=cfg.n_segs)
        pc0 = prev_pc[parent]
        xp = prev_x[parent]
        pc1 = _reflect(pc0 + rng.normal(0.0, 0.7, size=cfg.n_segs), hi)
        xc = _features(cfg, means, pc1, rng)
        w = np.exp(rng.normal(0.0, 3.0, size=cfg.n_segs))
        w /= w.sum()
        its.append({"pcoord0": pc0[:, None].copy(), "pcoord1": pc1[:, None].copy(), "weights": w, "parent": xp.copy(),
                    "child": xc})
        prev_pc, prev_x = pc1, xc
    return its


def _reflect(x, hi):
    x = np.abs(x)
    x = np.where(x > hi, 2 * hi - x, x)
    return np.clip(x, 0.0, hi)


def _features(cfg, means, pc, rng):
    b = np.minimum(pc.astype(np.int64), cfg.n_bins - 1)
    k = rng.integers(0, cfg.k_per_bin, size=pc.shape[0])
    return means[b, k] + rng.normal(0.0, 1.0, size=(pc.shape[0], cfg.dim))


def to_iteration_source(its):
    from msm_we_b200._hamsm._data import ArrayIterationSource, IterationRecord

    src = ArrayIterationSource()
    for i, it in enumerate(its, start=1):
        src.add(i, IterationRecord(it["pcoord0"], it["pcoord1"], it["weights"], it["parent"], it["child"]))
    return src


def generate_device(cfg: WEConfig, device, means=None, seed_offset=0, iters=None):
    """Device-resident stacked batch for the benchmark: returns a dict of CUDA tensors
    ``X [2N, D]`` (parents then children), ``pcoord [2N, 1]``, ``weights [N]``, ``iter_offsets [I+1]``.
    Generated with torch's device RNG (no PCIe traffic)."""
    import torch

    n_iters = cfg.n_iters if iters is None else iters
    S, D, B, K = cfg.n_segs, cfg.dim, cfg.n_bins, cfg.k_per_bin
    N = n_iters * S
    g = torch.Generator(device=device)
    g.manual_seed(cfg.seed + 77 + seed_offset)
    if means is None:
        means, _ = make_centers(cfg)
    means_d = torch.from_numpy(np.ascontiguousarray(means)).to(device).reshape(B * K, D)
    hi = B - 1e-3
    X = torch.empty((2 * N, D), dtype=torch.float64, device=device)
    pc = torch.empty((2 * N, 1), dtype=torch.float64, device=device)
    w = torch.empty(N, dtype=torch.float64, device=device)

    def feats(p, out):
        b = p.to(torch.int64).clamp_(max=B - 1)
        k = torch.randint(0, K, (p.numel(),), generator=g, device=device)
        torch.index_select(means_d, 0, b * K + k, out=out)
        out.add_(torch.randn(out.shape, generator=g, device=device, dtype=torch.float64))

    prev_pc = torch.rand(S, generator=g, device=device, dtype=torch.float64) * hi
    prev_x = torch.empty((S, D), dtype=torch.float64, device=device)
    feats(prev_pc, prev_x)
    for it in range(n_iters):
        sl_p = slice(it * S, (it + 1) * S)
        sl_c = slice(N + it * S, N + (it + 1) * S)
        parent = torch.randint(0, S, (S,), generator=g, device=device)
        pc0 = prev_pc[parent]
        X[sl_p] = prev_x[parent]
        pc1 = (pc0 + 0.7 * torch.randn(S, generator=g, device=device, dtype=torch.float64)).abs_()
        pc1 = torch.where(pc1 > hi, 2 * hi - pc1, pc1).clamp_(0.0, hi)
        feats(pc1, X[sl_c])
        ww = torch.exp(3.0 * torch.randn(S, generator=g, device=device, dtype=torch.float64))
        w[sl_p] = ww / ww.sum()
        pc[sl_p, 0] = pc0
        pc[sl_c, 0] = pc1
        prev_pc, prev_x = pc1, X[sl_c]
    offs = torch.arange(0, N + 1, S, dtype=torch.int64, device=device)
    return {"X": X, "pcoord": pc, "weights": w, "iter_offsets": offs, "n": N, "means": means}


def generate_cfg4_device(device, n_transitions, n_clusters=20000, segs_per_iter=8000, seed_offset=0):
    """BASELINE config 4 (flux-matrix stress): ``n_transitions`` weighted transitions between
    ``n_clusters`` clusters x 2 history colours.  Start labels are Zipf-ish over the clusters (a few hot states, a long
    tail), the end label is the start label plus a two-sided geometric offset (locality: nnz << M^2), colours are
    Bernoulli(0.5) with the basis / target rule (a transition INTO the first / last cluster takes colour 0 / 1),
    weights exp(N(0, 3)) normalised per iteration of ``segs_per_iter`` transitions.  Device-resident CUDA tensors."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(20261020 + seed_offset)
    N = int(n_transitions)
    u = torch.rand(N, generator=g, device=device, dtype=torch.float64)
    start = (float(n_clusters) ** u - 1.0).to(torch.int64).clamp_(0, n_clusters - 1)          # log-uniform ~ Zipf(1)
    perm = torch.randperm(n_clusters, generator=g, device=device)                               # hot states scattered
    start = perm[start]
    mag = torch.empty(N, device=device, dtype=torch.float64).geometric_(0.15, generator=g).to(torch.int64) - 1
    sign = torch.randint(0, 2, (N,), generator=g, device=device) * 2 - 1
    end = (start + sign * mag).clamp_(0, n_clusters - 1)
    col0 = torch.randint(0, 2, (N,), generator=g, device=device, dtype=torch.uint8)
    col1 = col0.clone()
    col1[end == 0] = 0
    col1[end == n_clusters - 1] = 1
    w = torch.exp(3.0 * torch.randn(N, generator=g, device=device, dtype=torch.float64))
    n_it = (N + segs_per_iter - 1) // segs_per_iter
    offs = torch.clamp(torch.arange(0, n_it + 1, dtype=torch.int64, device=device) * segs_per_iter, max=N)
    seg = torch.div(torch.arange(N, device=device), segs_per_iter, rounding_mode="floor")
    tot = torch.zeros(n_it, dtype=torch.float64, device=device).index_add_(0, seg, w)
    w = w / tot[seg]
    return {"start": start.contiguous(), "end": end.contiguous(), "col0": col0, "col1": col1, "w": w, "iter_offsets": offs,
            "n": N, "n_clusters": n_clusters}
