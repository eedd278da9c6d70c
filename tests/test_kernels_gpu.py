"""GPU parity tests of the individual kernels against the CPU oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from msm_we_b200 import ops
    return ops


def dev():
    return torch.device("cuda:0")


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        x = x.to(dtype)
    return x.to(dev())


# ------------------------------------------------------------------ sort
@pytest.mark.parametrize("n,bits", [(0, 10), (1, 10), (777, 5), (5000, 19), (100000, 31), (300001, 40)])
def test_radix_sort_stable(n, bits):
    ops = _ops()
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << min(bits, 20), size=n, dtype=np.int64)  # many duplicates
    if bits > 20 and n:
        keys = keys * rng.integers(1, 1 << (bits - 20), size=n, dtype=np.int64)
        keys &= (1 << bits) - 1
    vals = np.arange(n, dtype=np.int32)
    k, v = t(keys), t(vals)
    ops.sort_pairs_(k, v, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k.cpu().numpy(), keys[order])
    assert np.array_equal(v.cpu().numpy(), vals[order])


# ------------------------------------------------------------------ K0
def test_bin_flags_rectilinear_matches_oracle():
    ops = _ops()
    rng = np.random.default_rng(1)
    bounds = [np.array([0, .2, .25, .3, .35, .4, .45, .5, .55, .6, .65, .7, np.inf], dtype=np.float32)]
    pc = rng.uniform(0, 1.2, size=(5000, 1))
    pc[:12, 0] = bounds[0][:12].astype(np.float64)           # exactly on boundaries
    pc[12:24, 0] = np.nextafter(bounds[0][:12].astype(np.float64), -1)  # just below (float32 rounding decides)
    basis = np.array([[0.0, 0.15]]); target = np.array([[0.7, 100.0]])
    remap = np.arange(12, dtype=np.int32); remap[11] = 0
    mapper = ops.MapperSpec.rectilinear(bounds, dev())
    b, f = ops.bin_flags(t(pc), mapper, basis, target, we_remap=t(remap))
    om = O.RectilinearBinMapperOracle(bounds)
    ref_bin = remap[om.assign(pc)]
    assert np.array_equal(b.cpu().numpy(), ref_bin)
    ref_flag = O.is_we_region(pc, basis).astype(np.uint8) | (O.is_we_region(pc, target).astype(np.uint8) << 1)
    assert np.array_equal(f.cpu().numpy(), ref_flag)


def test_bin_flags_2d_and_out_of_range():
    ops = _ops()
    rng = np.random.default_rng(2)
    bounds = [np.linspace(0, 1, 6), np.linspace(-1, 1, 4)]
    pc = np.stack([rng.uniform(0, 0.999, 3000), rng.uniform(-1, 0.999, 3000)], axis=1)
    mapper = ops.MapperSpec.rectilinear(bounds, dev())
    basis = np.array([[0, .1], [-1, 0]]); target = np.array([[.9, 2], [0, 2]])
    b, f = ops.bin_flags(t(pc), mapper, basis, target)
    om = O.RectilinearBinMapperOracle(bounds)
    assert np.array_equal(b.cpu().numpy(), om.assign(pc))
    ref_flag = O.is_we_region(pc, basis).astype(np.uint8) | (O.is_we_region(pc, target).astype(np.uint8) << 1)
    assert np.array_equal(f.cpu().numpy(), ref_flag)
    # outside the bin space -> the reference (westpa) raises ValueError
    pc[7, 0] = 1.5
    errs = ops.DeviceErrors(dev())
    ops.bin_flags(t(pc), mapper, basis, target, errors=errs)
    with pytest.raises(ValueError):
        errs.check()


def test_bin_flags_voronoi():
    ops = _ops()
    rng = np.random.default_rng(3)
    centers = rng.normal(size=(17, 2))
    pc = rng.normal(size=(4000, 2))
    mapper = ops.MapperSpec.voronoi(centers, dev())
    b, _ = ops.bin_flags(t(pc), mapper, np.array([[9, 10], [9, 10]]), np.array([[-10, -9], [-10, -9]]))
    assert np.array_equal(b.cpu().numpy(), O.VoronoiBinMapperOracle(centers).assign(pc))


# ------------------------------------------------------------------ K1
def _strat_case(rng, N, D, nbins, K, ragged=False):
    ks = rng.integers(1, K + 1, size=nbins) if ragged else np.full(nbins, K)
    offs = np.concatenate([[0], np.cumsum(ks)]).astype(np.int64)
    centers = rng.normal(size=(offs[-1], D)) * 2.0
    bins = rng.integers(0, nbins, size=N).astype(np.int32)
    X = centers[offs[bins] + rng.integers(0, 1 << 30, size=N) % ks[bins]] + rng.normal(size=(N, D))
    flags = rng.choice([0, 0, 0, 0, 1, 2, 3], size=N).astype(np.uint8)
    return X, bins, flags, centers, offs, ks


def _ref_labels(X, bins, flags, centers, offs):
    T = offs[-1]
    out = np.zeros(len(X), dtype=np.int64)
    margins = np.full(len(X), np.inf)
    for b in np.unique(bins):
        sel = np.where((bins == b) & (flags == 0))[0]
        if len(sel) == 0:
            continue
        lab, amb = O.kmeans_assign_tiebreak(X[sel], centers[offs[b]:offs[b + 1]], return_ambiguous=True)
        out[sel] = lab + offs[b]
        margins[sel] = np.where(amb, 0.0, 1.0)
    out[(flags & 1) != 0] = T
    out[(flags & 2) != 0] = T + 1   # target tested first
    return out, margins


@pytest.mark.parametrize("N,D,nbins,K,ragged", [
    (1, 3, 1, 1, False),
    (1000, 13, 12, 25, False),      # the bundled NTL9 shape
    (5000, 64, 30, 20, False),      # BASELINE cfg2 shape
    (3000, 7, 5, 100, True),        # odd D -> 8-byte copy path, ragged K
    (2000, 33, 3, 130, False),      # K > 128: two centre blocks
    (700, 300, 4, 50, True),        # many k-chunks
])
def test_assign_matches_oracle(N, D, nbins, K, ragged):
    ops = _ops()
    rng = np.random.default_rng(N * 7 + D)
    X, bins, flags, centers, offs, ks = _strat_case(rng, N, D, nbins, K, ragged)
    c = t(centers)
    lab, local = ops.assign_stratified(t(X), t(bins), t(flags), c, ops.centers_sqnorm(c), t(offs), int(ks.max()), want_local=True)
    ref, margins = _ref_labels(X, bins, flags, centers, offs)
    got = lab.cpu().numpy()
    safe = margins > 1e-11          # rows with a score on the very edge of the tie band (none expected)
    assert safe.all()
    assert np.array_equal(got[safe], ref[safe])
    free = flags == 0
    assert np.array_equal(local.cpu().numpy()[free & safe], (ref - offs[bins])[free & safe])


@pytest.mark.parametrize("N,D,nbins,K,ragged", [
    (1, 3, 1, 1, False),
    (1000, 13, 12, 25, False),      # NTL9 shape; odd D -> scalar loader path
    (5000, 64, 30, 20, False),      # BASELINE cfg2 shape
    (3000, 7, 5, 100, True),
    (2000, 34, 3, 130, False),      # UMMA N = 144
    (1500, 24, 2, 300, False),      # K > 256: two centre blocks, points re-streamed
    (700, 300, 4, 50, True),        # ten k-chunks
    (4000, 256, 6, 100, False),     # BASELINE cfg5 shape
])
def test_assign_tcgen05_path_gives_identical_labels(N, D, nbins, K, ragged):
    """precision path 1 (tcgen05 split-TF32 candidate pass + fp64 re-check) must label exactly like the
    fp64 path: the fast pass only decides what its error bound allows, the rest is re-evaluated in fp64."""
    from msm_we_b200 import _lib
    ops = _ops()
    rng = np.random.default_rng(N * 11 + D)
    X, bins, flags, centers, offs, ks = _strat_case(rng, N, D, nbins, K, ragged)
    X += 5.0                                    # a common offset: the kernel centres by the bin mean
    centers = centers + 5.0
    c = t(centers)
    csq = ops.centers_sqnorm(c)
    lab_tc, loc_tc = ops.assign_stratified(t(X), t(bins), t(flags), c, csq, t(offs), int(ks.max()),
                                           path=_lib.ASSIGN_TF32X3, want_local=True)
    lab_64 = ops.assign_stratified(t(X), t(bins), t(flags), c, csq, t(offs), int(ks.max()), path=_lib.ASSIGN_FP64)
    assert np.array_equal(lab_tc.cpu().numpy(), lab_64.cpu().numpy())
    ref, margins = _ref_labels(X, bins, flags, centers, offs)
    assert np.array_equal(lab_tc.cpu().numpy()[margins > 0], ref[margins > 0])


@pytest.mark.parametrize("path_name", ["FP64", "TF32X3"])
@pytest.mark.parametrize("N,D,nbins,K", [(5000, 64, 30, 20), (6000, 256, 6, 100), (3000, 7, 5, 100)])
def test_assign_reusing_the_buckets_of_the_previous_call(path_name, N, D, nbins, K):
    """Lloyd iterations label the same points against new centres: with MWE_ASSIGN_REUSE_BUCKETS the second call skips
    the bucketing by WE bin (private workspace + the same label buffer) and must label like a fresh call."""
    from msm_we_b200 import _lib
    ops = _ops()
    path = getattr(_lib, "ASSIGN_" + path_name)
    rng = np.random.default_rng(N + D)
    X, bins, flags, centers, offs, ks = _strat_case(rng, N, D, nbins, K, False)
    Xd, bd, fd, od = t(X), t(bins), t(flags), t(offs)
    ws = ops.assign_workspace(Xd, nbins, K, path)
    labels = torch.empty(N, dtype=torch.int64, device=Xd.device)
    for it in range(3):
        c = t(centers + 0.3 * it * rng.normal(size=centers.shape))
        csq = ops.centers_sqnorm(c)
        ops.assign_stratified(Xd, bd, fd, c, csq, od, K, path=path, label_out=labels, workspace=ws, reuse_buckets=it > 0)
        fresh = ops.assign_stratified(Xd, bd, fd, c, csq, od, K, path=path)
        assert np.array_equal(labels.cpu().numpy(), fresh.cpu().numpy()), it


def test_assign_tcgen05_scores_within_error_bound():
    """The rigorous part of the fast path is its error bound: dump the scores the tensor cores produced and
    compare them with fp64 (centred by the bin mean, as the kernel does)."""
    import ctypes
    from msm_we_b200 import _lib
    ops = _ops()
    rng = np.random.default_rng(99)
    worst = 0.0
    for N, D, K in [(600, 64, 20), (500, 300, 50), (400, 13, 25)]:
        centers = rng.normal(size=(K, D)) * 3 + 10.0
        X = centers[rng.integers(0, K, N)] + rng.normal(size=(N, D))
        bins = np.zeros(N, dtype=np.int32); flags = np.zeros(N, dtype=np.uint8)
        offs = np.array([0, K], dtype=np.int64)
        ncols = _lib.lib.mwe_debug_tc_columns(K)
        dbg = torch.full((N, ncols), float("nan"), dtype=torch.float32, device=dev())
        _lib.check(_lib.lib.mwe_debug_set_tc_scores(dbg.data_ptr()), "dbg")
        try:
            c = t(centers)
            ops.assign_stratified(t(X), t(bins), t(flags), c, ops.centers_sqnorm(c), t(offs), K, path=_lib.ASSIGN_TF32X3)
            torch.cuda.synchronize()
        finally:
            _lib.lib.mwe_debug_set_tc_scores(None)
        got = dbg.cpu().numpy()[:, :K].astype(np.float64)
        m = centers.mean(axis=0)
        Xc, Cc = X - m, centers - m
        ref = (Cc * Cc).sum(axis=1)[None, :] - 2.0 * Xc @ Cc.T
        cmax = np.sqrt((Cc * Cc).sum(axis=1).max()); xn = np.sqrt((Xc * Xc).sum(axis=1))
        coef = 2.0 ** -20 + (3 * ((D + 7) // 8) + 10) * 2.0 ** -22
        bound = coef * cmax * (2 * xn + cmax)
        ratio = (np.abs(got - ref) / bound[:, None]).max()
        assert np.isfinite(got).all()
        worst = max(worst, ratio)
    assert worst < 1.0, worst          # measured ratio is reported in profiles/; the bound must never be exceeded


@pytest.mark.parametrize("N,D,nbins,K,ragged", [
    (300000, 64, 30, 20, False),    # BASELINE cfg2 shape, several tiles per CTA
    (4000, 6, 3, 9, True),          # D not a multiple of 8: zero-padded k steps
    (4000, 2, 7, 64, True),         # smallest even D, widest accumulator block
    (3000, 66, 4, 16, False),       # rows of 528 bytes: general copy loop
    (5000, 130, 5, 8, True),        # long rows, few buffers
    (6000, 32, 400, 5, True),       # hundreds of tiny bins: centre buffers recycled with groups still in flight
    (257, 64, 2, 1, False),         # one centre per bin
])
def test_assign_resident_kernel_matches_streaming_kernel_and_oracle(N, D, nbins, K, ragged, monkeypatch):
    """The resident-centre producer/consumer kernel (assign_res.cu) is what the fp64 path runs for K <= 64 and
    16-byte aligned rows; the streaming kernel (assign.cu) is its fallback.  Same labels from both, and from
    the oracle."""
    ops = _ops()
    rng = np.random.default_rng(N + 31 * D)
    X, bins, flags, centers, offs, ks = _strat_case(rng, N, D, nbins, K, ragged)
    c = t(centers)
    args = (t(X), t(bins), t(flags), c, ops.centers_sqnorm(c), t(offs), int(ks.max()))
    from msm_we_b200 import _lib
    monkeypatch.delenv("MWE_ASSIGN_RESIDENT", raising=False)
    lab_res = ops.assign_stratified(*args, path=_lib.ASSIGN_FP64).cpu().numpy()
    monkeypatch.setenv("MWE_ASSIGN_RESIDENT", "0")
    lab_str = ops.assign_stratified(*args, path=_lib.ASSIGN_FP64).cpu().numpy()
    assert np.array_equal(lab_res, lab_str)
    if N <= 6000:
        ref, margins = _ref_labels(X, bins, flags, centers, offs)
        safe = margins > 1e-11
        assert np.array_equal(lab_res[safe], ref[safe])


def test_assign_ties_pick_lowest_index():
    ops = _ops()
    rng = np.random.default_rng(5)
    D, K = 16, 24
    centers = rng.integers(-3, 4, size=(K, D)).astype(np.float64)
    centers[7] = centers[3]; centers[20] = centers[3]; centers[11] = centers[2]   # exact duplicates
    X = np.concatenate([centers + 0.0, centers[rng.integers(0, K, 500)] + rng.integers(-1, 2, size=(500, D))])
    bins = np.zeros(len(X), dtype=np.int32); flags = np.zeros(len(X), dtype=np.uint8)
    offs = np.array([0, K], dtype=np.int64)
    c = t(centers)
    lab = ops.assign_stratified(t(X), t(bins), t(flags), c, ops.centers_sqnorm(c), t(offs), K)
    # small integers: every product and sum is exact in fp64, so the oracle's argmin is exact too
    assert np.array_equal(lab.cpu().numpy(), O.kmeans_assign(X, centers))
    assert np.array_equal(lab.cpu().numpy()[:60], O.kmeans_assign_exact(X[:60], centers))


def test_assign_strided_rows_and_empty_bin_error():
    ops = _ops()
    rng = np.random.default_rng(6)
    N, D = 500, 20
    big = t(rng.normal(size=(N, 2 * D)))
    X = big[:, :D]                                 # row stride 2D
    centers = rng.normal(size=(30, D)); offs = np.array([0, 10, 10, 30], dtype=np.int64)   # bin 1 has no centres
    bins = rng.choice([0, 2], size=N).astype(np.int32); flags = np.zeros(N, dtype=np.uint8)
    c = t(centers)
    lab = ops.assign_stratified(X, t(bins), t(flags), c, ops.centers_sqnorm(c), t(offs), 20)
    ref, m = _ref_labels(X.cpu().numpy(), bins, flags, centers, offs)
    assert np.array_equal(lab.cpu().numpy()[m > 1e-11], ref[m > 1e-11])
    bins[3] = 1
    errs = ops.DeviceErrors(dev())
    ops.assign_stratified(X, t(bins), t(flags), c, ops.centers_sqnorm(c), t(offs), 20, errors=errs)
    with pytest.raises(AssertionError):
        errs.check()


# ------------------------------------------------------------------ K2
@pytest.mark.parametrize("N,D,K,weighted", [(400, 13, 25, False), (3000, 64, 40, True), (50, 5, 60, True)])
def test_minibatch_update_bit_exact(N, D, K, weighted):
    ops = _ops()
    rng = np.random.default_rng(N)
    X = rng.normal(size=(N, D)); centers = rng.normal(size=(K, D)); counts = rng.integers(0, 50, size=K).astype(np.float64)
    w = np.exp(rng.normal(0, 3, size=N)) if weighted else np.ones(N)
    labels = rng.integers(0, K, size=N).astype(np.int64)
    labels[::17] = K + 1                               # basis/target-like labels are skipped
    ref_c, ref_n = centers.copy(), counts.copy()
    valid = labels < K
    O.minibatch_update(X[valid], w[valid], ref_c, ref_n, labels[valid])
    c, n = t(centers), t(counts)
    ops.minibatch_update(t(X), t(w) if weighted else None, t(labels), c, n)
    assert np.array_equal(c.cpu().numpy(), ref_c)      # same order, same rounding: bit-exact
    assert np.array_equal(n.cpu().numpy(), ref_n)


def test_lloyd_accumulate_and_finalize():
    ops = _ops()
    rng = np.random.default_rng(11)
    N, D, K = 5000, 32, 30
    X = rng.normal(size=(N, D)); centers = rng.normal(size=(K, D)); w = rng.uniform(0.1, 2, size=N)
    labels, new, wsum = O.lloyd_iter(X, w, centers)
    sum_wx, sum_w = ops.centroid_accumulate(t(X), t(w), t(labels.astype(np.int64)), K)
    assert np.array_equal(sum_w.cpu().numpy(), wsum)
    c = t(centers)
    ops.lloyd_finalize(sum_wx, sum_w, c)
    assert np.array_equal(c.cpu().numpy(), new)


@pytest.mark.parametrize("D", [256, 250, 257, 1100, 130])
def test_accumulate_skewed_cluster_sizes_bit_exact(D):
    """Clusters of >= 2048 members are split over several CTAs by feature columns (64-column blocks up to D = 768,
    256-column blocks above) and scheduled first: every (cluster, feature) sum must still be the sequential
    sample-order sum, bit for bit, next to small and empty clusters."""
    ops = _ops()
    rng = np.random.default_rng(1234 + D)
    N, K = 9000, 9
    labels = rng.choice([1, 4, 6, 7, 8], size=N, p=[0.55, 0.30, 0.1, 0.04, 0.01]).astype(np.int64)   # 0, 2, 3, 5 empty
    X = rng.normal(size=(N, D)) * 10.0 ** rng.integers(-3, 4, size=(N, 1))
    w = np.exp(rng.normal(0, 2, size=N))
    sums = np.zeros((K, D)); wsum = np.zeros(K)
    for k in range(K):
        rows = np.flatnonzero(labels == k)
        acc = np.zeros(D); a = 0.0
        for i in rows:
            acc = acc + X[i] * w[i]
            a += w[i]
        sums[k], wsum[k] = acc, a
    for weights in (w, None):
        sum_wx, sum_w = ops.centroid_accumulate(t(X), None if weights is None else t(weights), t(labels), K)
        if weights is None:
            ref = np.zeros((K, D))
            for k in range(K):
                acc = np.zeros(D)
                for i in np.flatnonzero(labels == k):
                    acc = acc + X[i]
                ref[k] = acc
            assert np.array_equal(sum_wx.cpu().numpy(), ref)
            assert np.array_equal(sum_w.cpu().numpy(), np.bincount(labels, minlength=K).astype(float))
        else:
            assert np.array_equal(sum_wx.cpu().numpy(), sums)
            assert np.array_equal(sum_w.cpu().numpy(), wsum)


@pytest.mark.parametrize("n_labels,N", [(40, 60000), (2000, 60000), (7, 0)])
def test_group_by_label_and_label_stats(n_labels, N):
    """Stable group-by-label + NaN-skipping per-label count / sum / min / max (get_cluster_centers, relocation): both
    launch shapes (a CTA per label for few labels, a warp per label otherwise), empty labels, NaN values, a column view."""
    ops = _ops()
    rng = np.random.default_rng(n_labels + N)
    labels = rng.integers(0, n_labels, size=N).astype(np.int64)
    labels[labels == 3] = 4                                   # label 3 stays empty
    vals = rng.normal(size=(N, 3)) * 10.0 ** rng.integers(-2, 3, size=(N, 1))
    vals[rng.random(N) < 0.01, 1] = np.nan
    members, seg = ops.group_by_label(t(labels), n_labels)
    mh, sh = members.cpu().numpy().astype(np.int64), seg.cpu().numpy()
    assert sh[0] == 0 and sh[n_labels] == N
    for k in range(n_labels):
        assert np.array_equal(mh[sh[k]:sh[k + 1]], np.flatnonzero(labels == k))     # input order inside a label
    V = t(vals)
    count, ssum, vmin, vmax = ops.label_stats(V[:, 1], members, seg, n_labels)
    count, ssum, vmin, vmax = (x.cpu().numpy() for x in (count, ssum, vmin, vmax))
    for k in range(n_labels):
        x = vals[labels == k, 1]
        x = x[~np.isnan(x)]
        assert count[k] == len(x)
        if len(x):
            assert vmin[k] == x.min() and vmax[k] == x.max()
            assert abs(ssum[k] - x.sum()) <= 1e-12 * np.abs(x).sum()
        else:
            assert ssum[k] == 0.0 and vmin[k] == np.inf and vmax[k] == -np.inf


@pytest.mark.parametrize("k", [1, 3, 8])
def test_segment_topk_is_the_head_of_a_stable_descending_sort(k):
    ops = _ops()
    rng = np.random.default_rng(77 + k)
    n_labels, N = 12, 50000
    labels = rng.integers(0, n_labels, size=N).astype(np.int64)
    labels[labels == 5] = 6                                   # group 5 empty
    labels[:2] = 11; labels[labels == 11] = 10; labels[:2] = 11   # group 11 has exactly two members
    v = np.round(rng.normal(size=N), 2)                       # many exactly equal values: ties are resolved by member order
    v[rng.random(N) < 0.01] = np.nan
    members, seg = ops.group_by_label(t(labels), n_labels)
    ids = np.array([0, 5, 11, 3, 7], dtype=np.int32)
    pos, val = ops.segment_topk(t(v), members, seg, t(ids), k)
    pos, val = pos.cpu().numpy(), val.cpu().numpy()
    for i, g in enumerate(ids):
        mem = np.flatnonzero(labels == g)
        mem = mem[~np.isnan(v[mem])]
        want = mem[np.lexsort((mem, -v[mem]))][:k]
        assert np.array_equal(pos[i, :len(want)], want), g
        assert np.array_equal(val[i, :len(want)], v[want])
        assert (pos[i, len(want):] == -1).all()


# ------------------------------------------------------------------ K3
def _flux_case(rng, n, iters, segs):
    per = []
    for _ in range(iters):
        S = int(rng.integers(max(1, segs // 2), segs + 1))
        pairs = rng.integers(0, n, size=(S, 2)).astype(np.int64)
        p0 = rng.uniform(0, 10, size=(S, 1)); p1 = rng.uniform(0, 10, size=(S, 1))
        w = np.exp(rng.normal(0, 3, size=S)); w /= w.sum()
        per.append((pairs, p0, p1, w))
    return per


def test_flux_dense_bit_exact_with_serial_reference():
    ops = _ops()
    rng = np.random.default_rng(21)
    n = 40
    basis = np.array([[0.0, 1.0]]); target = np.array([[9.0, 10.0]])
    per = _flux_case(rng, n, iters=25, segs=200)
    ref = O.flux_matrix(n, per, basis, target)
    pairs = np.concatenate([p[0] for p in per]); p0 = np.concatenate([p[1] for p in per]); p1 = np.concatenate([p[2] for p in per])
    w = np.concatenate([p[3] for p in per])
    offs = np.concatenate([[0], np.cumsum([len(p[3]) for p in per])]).astype(np.int64)
    mapper = ops.MapperSpec.rectilinear([np.array([0, 5, 10.0])], dev())
    _, f0 = ops.bin_flags(t(p0), mapper, basis, target)
    _, f1 = ops.bin_flags(t(p1), mapper, basis, target)
    dense = ops.flux_accumulate(t(pairs[:, 0].copy()), t(pairs[:, 1].copy()), t(w), n, flag0=f0, flag1=f1,
                                iter_offsets=t(offs))
    ops.divide_(dense, float(len(per)))
    got = dense.cpu().numpy()
    assert np.array_equal(got, ref)                    # same association as the serial reference: bit-exact
    # counts (unit weights) are exact integers
    cnt = ops.flux_accumulate(t(pairs[:, 0].copy()), t(pairs[:, 1].copy()), None, n, flag0=f0, flag1=f1).cpu().numpy()
    ref_cnt = sum(O.iter_flux_matrix(n, p[0], p[1], p[2], np.ones(len(p[3])), basis, target) for p in per)
    assert np.array_equal(cnt, ref_cnt)


def test_flux_coo_and_label_range_error():
    ops = _ops()
    rng = np.random.default_rng(22)
    n, N = 500, 20000
    s = rng.integers(0, n + 2, size=N).astype(np.int64); e = rng.integers(0, n + 2, size=N).astype(np.int64)
    w = rng.uniform(size=N)
    dense, (r, c, v, nnz) = ops.flux_accumulate(t(s), t(e), t(w), n, want_coo=True,
                                                dense=torch.zeros(n + 2, n + 2, dtype=torch.float64, device=dev()))
    k = int(nnz.item())
    from scipy.sparse import coo_matrix
    ref = np.asarray(coo_matrix((w, (s, e)), shape=(n + 2, n + 2)).todense())
    assert np.array_equal(dense.cpu().numpy(), ref)
    rr, cc, vv = r[:k].cpu().numpy(), c[:k].cpu().numpy(), v[:k].cpu().numpy()
    assert k == np.count_nonzero(ref) and np.all(np.diff(rr * (n + 2) + cc) > 0)
    assert np.array_equal(ref[rr, cc], vv)
    s[5] = n + 2
    errs = ops.DeviceErrors(dev())
    ops.flux_accumulate(t(s), t(e), t(w), n, errors=errs)
    with pytest.raises(ValueError):
        errs.check()


def test_flux_colour_kat(golden_dir):
    """The reference's own known-answer test for the coloured scatter (tests/test_non_markov_model.py:8-26)."""
    ops = _ops()
    kat = np.load(f"{golden_dir}/colour_kat.npz")
    np.random.seed(int(kat["seed"]))
    traj = np.random.randint(0, 3, int(kat["n"]))
    s0, s1, c0, c1 = O.colour_transitions([traj], [0], [2], int(kat["lag"]))
    # n_clusters + 2 == 3 states -> n_clusters = 1
    dense = ops.flux_accumulate(t(s0), t(s1), None, 1, col0=t(c0), col1=t(c1), C=2).cpu().numpy()
    assert np.array_equal(dense, O.colour_counts([traj], 3, [0], [2], int(kat["lag"])))
    assert np.allclose(O.normalize_markov_matrix(dense), kat["nmm_tmatrix"])


# ------------------------------------------------------------------ projection
@pytest.mark.parametrize("N,D_in,d_out,use_mean", [
    (1, 3, 1, True), (1000, 80, 13, True),          # the bundled NTL9 shape: 80 features -> 13 components
    (777, 33, 9, False),                            # odd D_in: 8-byte copy path
    (5000, 300, 64, True), (300, 70, 100, True),    # two column blocks
])
def test_projection_matches_numpy(N, D_in, d_out, use_mean):
    """(X - mean) @ components.T: the centred coordinates are rounded like numpy's, only the summation order of the
    matmul differs from BLAS -> 1e-12 relative to the magnitude of the terms."""
    ops = _ops()
    rng = np.random.default_rng(N + D_in)
    X = rng.normal(size=(N, D_in)) * 3 + 10
    W = rng.normal(size=(d_out, D_in))
    mean = X.mean(axis=0) if use_mean else None
    ref = O.linear_transform(X, W, mean)
    got = ops.project(t(X), t(W), None if mean is None else t(mean)).cpu().numpy()
    scale = np.abs(X - (0 if mean is None else mean)) @ np.abs(W).T
    assert np.all(np.abs(got - ref) <= 1e-12 * scale + 1e-300)
    # strided input view
    big = t(np.concatenate([X, X], axis=1))
    got2 = ops.project(big[:, :D_in], t(W), None if mean is None else t(mean)).cpu().numpy()
    assert np.array_equal(got, got2)
