"""CPU: the oracle against the reference's own golden vectors / known answers and against the installed
scikit-learn + scipy (the third-party packages that hold the reference's arithmetic)."""
import numpy as np
import pytest

from oracle import oracle as O


def test_colour_scatter_known_answer(golden_dir):
    """reference: tests/test_non_markov_model.py:8-26 (seed 192348, randint(0,3,100000), lag 100, A=[0], B=[2])."""
    kat = np.load(f"{golden_dir}/colour_kat.npz")
    np.random.seed(int(kat["seed"]))
    traj = np.random.randint(0, 3, int(kat["n"]))
    counts = O.colour_counts([traj], 3, [0], [2], int(kat["lag"]), sliding_window=True)
    assert np.allclose(O.normalize_markov_matrix(counts), kat["nmm_tmatrix"])
    s0, s1, c0, c1 = O.colour_transitions([traj], [0], [2], int(kat["lag"]))
    rebuilt = np.zeros((6, 6))
    np.add.at(rebuilt, (2 * s0 + c0, 2 * s1 + c1), 1.0)
    assert np.array_equal(rebuilt, counts) and counts.sum() == len(s0) == 99758


def test_ntl9_pair_dtrajs_reproduce_golden_flux_pattern(golden_dir):
    """clustered.obj pair_dtrajs (labels 275/276 = predict's basis/target) scattered with the reference's
    build_flux_matrix conventions give exactly the non-zero pattern of fluxmatrix_raw.npy; iteration 1 is
    skipped by get_fluxMatrix's range(first_iter + 1, maxIter)."""
    g = np.load(f"{golden_dir}/ntl9_clustered.npz")
    n = int(g["n_clusters"])
    assert n == 300 and tuple(g["flux_raw_shape"]) == (302, 302)
    offs = np.concatenate([[0], np.cumsum(g["pair_lens"])])
    total = np.zeros((n + 2, n + 2))
    for it in range(1, len(g["pair_lens"])):     # index 0 is WE iteration 1
        s = g["pair_parent"][offs[it]:offs[it + 1]].copy(); e = g["pair_child"][offs[it]:offs[it + 1]].copy()
        pairs = np.stack([s, e], axis=1)
        sb = np.where(s == 275); eb = np.where(e == 275); et = np.where(e == 276)
        # parents in the target are NOT relabelled by the reference; none occur after recycling
        assert not np.any(s == 276)
        total += np.asarray(O.build_flux_matrix(n, pairs, sb, eb, et, np.ones(len(s))).todense())
    got = set(zip(*np.nonzero(total)))
    ref = set(zip(g["flux_raw_nz_i"].tolist(), g["flux_raw_nz_j"].tolist()))
    assert got == ref and len(ref) == 4575
    # including iteration 1 would add cells that are not in the golden matrix
    s = g["pair_parent"][:offs[1]]; e = g["pair_child"][:offs[1]]
    extra = set(zip(np.where(s == 275, n, s).tolist(), np.where(e == 275, n, e).tolist())) - ref
    assert len(extra) > 0


def test_ntl9_bin_mapper_and_centres_fixture(golden_dir):
    g = np.load(f"{golden_dir}/ntl9_clustered.npz")
    bounds = [g["boundaries"]]
    om = O.RectilinearBinMapperOracle(bounds)
    assert om.nbins == 12 and g["fitted"].sum() == 11 and int(g["we_remap"][11]) == 0
    assert all(g[f"centers_{b}"].shape == (25, 13) for b in range(11))
    # float32 boundary semantics: lower <= x < upper after casting x to float32
    x = np.array([0.0, 0.19999999, 0.2, 0.7, 1e9])
    assert om.assign(x[:, None]).tolist() == [0, 0, 1, 11, 11]
    with pytest.raises(ValueError):
        om.assign(np.array([[-1e-3]]))


def test_downstream_chain_from_golden_fluxmatrix(golden_dir):
    d = np.load(f"{golden_dir}/ntl9_downstream.npz")
    T = O.transition_matrix(d["fluxmatrix"], d["indBasis"], d["indTargets"])
    # row sums are pairwise in numpy; the golden file was written by an older numpy, so allow 1 ulp
    assert np.allclose(T, d["tmatrix"], rtol=1e-14, atol=0)
    pss = O.steady_state(T)
    assert np.allclose(pss, d["pSS"], rtol=1e-6, atol=1e-12)
    assert abs(pss.sum() - 1.0) < 1e-12 and d["JtargetSS"] > 0


@pytest.mark.parametrize("N,D,K", [(500, 13, 25), (2000, 64, 20), (300, 7, 100)])
def test_assignment_oracle_matches_sklearn_predict(N, D, K):
    rng = np.random.default_rng(N + K)
    centers = rng.normal(size=(K, D)) * 2
    X = centers[rng.integers(0, K, N)] + rng.normal(size=(N, D))
    model = O.make_fitted_minibatch(centers)
    ref = model.predict(X)
    assert np.array_equal(O.kmeans_assign(X, centers), ref)
    lab, amb = O.kmeans_assign_tiebreak(X, centers, return_ambiguous=True)
    assert not amb.any() and np.array_equal(lab, ref)
    # per-segment calls (the literal reference loop) agree with the batched call
    assert np.array_equal(np.array([model.predict([x])[0] for x in X[:50]]), ref[:50])


def test_tiebreak_rule_on_duplicates_and_near_duplicates():
    rng = np.random.default_rng(1)
    centers = rng.normal(size=(10, 6))
    centers[7] = centers[2]
    centers[9] = centers[2] * (1 + 2e-16)              # ulp-level near-duplicate
    X = centers[[2, 2, 5]] + 1e-3 * rng.normal(size=(3, 6))
    lab = O.kmeans_assign_tiebreak(X, centers)
    assert lab.tolist() == [2, 2, 5]
    assert np.array_equal(O.kmeans_assign_exact(X[2:], centers), [5])


@pytest.mark.parametrize("weighted", [False, True])
def test_minibatch_restatement_matches_sklearn_partial_fit(weighted):
    from sklearn.cluster import MiniBatchKMeans

    rng = np.random.default_rng(5)
    K, D = 12, 9
    init = rng.normal(size=(K, D))
    sk = MiniBatchKMeans(n_clusters=K, init=init.copy(), n_init=1, reassignment_ratio=0.0)
    centers, counts = init.copy(), np.zeros(K)
    for _ in range(4):
        X = rng.normal(size=(200, D)) + rng.integers(0, 3, size=(200, 1))
        w = np.exp(rng.normal(0, 2, size=200)) if weighted else None
        sk.partial_fit(X, sample_weight=w)
        O.mini_batch_step(X, np.ones(200) if w is None else w, centers, counts, None, False)
        assert np.array_equal(centers, sk.cluster_centers_)      # same order, same rounding
        assert np.array_equal(counts, sk._counts)


def test_lloyd_restatement_matches_sklearn_kmeans_single_iteration():
    from sklearn.cluster import KMeans

    rng = np.random.default_rng(6)
    K, D, N = 8, 5, 600
    X = rng.normal(size=(N, D)) + 3 * rng.integers(0, 3, size=(N, 1))
    init = X[:K].copy()
    w = rng.uniform(0.5, 2.0, size=N)
    labels, new, wsum = O.lloyd_iter(X, w, init)
    lab2, new2, wsum2 = O.lloyd_iter_fast(X, w, init)
    assert np.array_equal(labels, lab2) and np.allclose(new, new2, rtol=1e-13) and np.allclose(wsum, wsum2, rtol=1e-13)
    km = KMeans(n_clusters=K, init=init, n_init=1, max_iter=1, algorithm="lloyd", tol=0.0).fit(X, sample_weight=w)
    # KMeans centres the data before fitting, so agreement is to rounding, not bitwise
    assert np.allclose(km.cluster_centers_, new, rtol=1e-10, atol=1e-12)


def test_flux_matrix_conventions():
    n = 4
    pairs = np.array([[0, 1], [1, 1], [2, 3], [3, 0], [0, 1]])
    p0 = np.array([[5.0], [0.2], [5.0], [5.0], [5.0]])     # parent 1 in basis
    p1 = np.array([[5.0], [9.5], [0.3], [5.0], [5.0]])     # child 1 in target, child 2 in basis
    w = np.array([0.1, 0.2, 0.3, 0.4, 0.5])
    F = O.iter_flux_matrix(n, pairs, p0, p1, w, np.array([[0.0, 1.0]]), np.array([[9.0, 10.0]]))
    assert F.shape == (6, 6)
    assert F[0, 1] == pytest.approx(0.6) and F[n, n + 1] == 0.2 and F[2, n] == 0.3 and F[3, 0] == 0.4
    # a child in both regions ends up in the basis (override order of _fluxmatrix.py:135-137)
    F2 = O.iter_flux_matrix(n, pairs[:1], p0[:1], np.array([[0.5]]), w[:1], np.array([[0.0, 1.0]]), np.array([[0.2, 1.0]]))
    assert F2[0, n] == 0.1 and F2[0, n + 1] == 0.0


def test_linear_transform_restatement_matches_sklearn_pca_transform():
    """oracle.linear_transform restates the reference's coordinates.transform (sklearn < 1.1: (X - mean_) @
    components_.T).  The installed sklearn evaluates X @ components_.T - mean_ @ components_.T instead: same map,
    different rounding, so the two agree to 1e-12 of the magnitude of the terms."""
    from sklearn.decomposition import IncrementalPCA

    rng = np.random.default_rng(11)
    X = rng.normal(size=(400, 30)) * 2 + 5
    pca = IncrementalPCA(n_components=7).fit(X)
    ours = O.linear_transform(X, pca.components_, pca.mean_)
    theirs = pca.transform(X)
    scale = (np.abs(X) + np.abs(pca.mean_)) @ np.abs(pca.components_).T
    assert np.all(np.abs(ours - theirs) <= 1e-12 * scale)
