"""B200 parity of the device k-means initialisation (reference index/models/layers.py:69-82 =
``sklearn.cluster.KMeans(n_clusters, max_iter).fit``).  The checker is scikit-learn itself - the third-party routine the
reference calls (pinned 1.9.0 in this image) - run on the host with the same numpy seed.  Floating point: bar 1e-5 relative
on the centres and the inertia when the Lloyd label sequences coincide (same n_iter), which the seeded cases below do."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lcrec_b200 import ops
    from lcrec_b200.models import layers as L
    DEV = torch.device("cuda:0")


def blobs(n, d, k, seed, spread=0.25):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((k, d)).astype(np.float32)
    x = centres[rng.integers(0, k, size=n)] + spread * rng.standard_normal((n, d)).astype(np.float32)
    return (x + 3.0).astype(np.float32)              # off-centre: the centring matters


@pytest.mark.parametrize("n,d,k,iters,seed", [(1024, 32, 256, 100, 0), (1024, 32, 256, 10, 1), (2000, 16, 64, 100, 2),
                                              (5000, 32, 256, 100, 3), (600, 100, 20, 100, 4), (1024, 256, 300, 50, 5)])
def test_device_kmeans_matches_sklearn(n, d, k, iters, seed):
    from sklearn.cluster import KMeans
    x = blobs(n, d, max(k // 2, 2), seed)
    np.random.seed(2024 + seed)
    ref = KMeans(n_clusters=k, max_iter=iters).fit(x)
    np.random.seed(2024 + seed)
    old = L.KMEANS_BACKEND
    L.KMEANS_BACKEND = "device"
    try:
        got = L.kmeans(torch.from_numpy(x).to(DEV), k, iters)
    finally:
        L.KMEANS_BACKEND = old
    assert got.shape == (k, d) and got.dtype == torch.float32 and got.is_cuda
    c = got.cpu().numpy()
    scale = np.abs(ref.cluster_centers_).max()
    close = np.abs(c - ref.cluster_centers_).max(axis=1) <= 1e-5 * scale
    assert close.mean() >= 0.99, (close.mean(), np.abs(c - ref.cluster_centers_).max())
    lab = np.argmin(((x[:, None, :] - c[None]) ** 2).sum(-1), axis=1) if n * k * d < 3e7 else None
    if lab is not None:
        inertia = ((x - c[lab]) ** 2).sum()
        assert abs(inertia - ref.inertia_) <= 1e-4 * ref.inertia_


def test_device_lloyd_details():
    """labels / inertia / n_iter of the C entry point against sklearn started from the SAME explicit seeds (no RNG), and
    the empty-cluster relocation: seeds far away from the data own no point, yet every final cluster is populated."""
    from sklearn.cluster import KMeans
    x = blobs(1500, 24, 40, 11)
    rng = np.random.default_rng(3)
    seeds = x[rng.choice(len(x), 80, replace=False)].copy()
    ref = KMeans(n_clusters=80, init=seeds, n_init=1, max_iter=100).fit(x)
    xc, mean, mean_var = ops.kmeans_center(torch.from_numpy(x).to(DEV))
    np.testing.assert_allclose(mean.cpu().numpy(), x.astype(np.float64).mean(axis=0), rtol=2e-7)   # fp64 sums, one rounding
    np.testing.assert_allclose(mean_var, x.astype(np.float64).var(axis=0).mean(), rtol=1e-6)
    fit = ops.kmeans_lloyd(xc, torch.from_numpy(seeds).to(DEV) - mean, 100, 1e-4 * mean_var, add_mean=mean)
    assert fit["n_iter"] == ref.n_iter_
    assert (fit["labels"].cpu().numpy() != ref.labels_).mean() <= 0.002
    np.testing.assert_allclose(fit["inertia"], ref.inertia_, rtol=1e-4)
    np.testing.assert_allclose(fit["centers"].cpu().numpy(), ref.cluster_centers_, rtol=0, atol=1e-5 * np.abs(x).max())
    # relocation
    seeds2 = seeds.copy()
    seeds2[5] += 100.0
    seeds2[17] -= 100.0
    fit2 = ops.kmeans_lloyd(xc, torch.from_numpy(seeds2).to(DEV) - mean, 100, 1e-4 * mean_var, add_mean=mean)
    counts = np.bincount(fit2["labels"].cpu().numpy(), minlength=80)
    assert (counts > 0).all() and torch.isfinite(fit2["centers"]).all()
    ref2 = KMeans(n_clusters=80, init=seeds2, n_init=1, max_iter=100).fit(x)
    assert abs(fit2["inertia"] - ref2.inertia_) <= 0.05 * ref2.inertia_      # same objective; which point seeds which empty
    #                                                                          cluster is unspecified in sklearn (argpartition)


def test_vq_init_emb_uses_device_kmeans():
    """VectorQuantizer.init_emb (vq.py:40-49) through the device backend: codebook = centres, initted flips."""
    from lcrec_b200.models import VectorQuantizer
    x = torch.from_numpy(blobs(512, 16, 16, 21)).to(DEV)
    vq = VectorQuantizer(32, 16, kmeans_init=True, kmeans_iters=20).to(DEV).train()
    old = L.KMEANS_BACKEND
    L.KMEANS_BACKEND = "device"
    try:
        np.random.seed(1)
        vq(x, use_sk=False)
    finally:
        L.KMEANS_BACKEND = old
    assert vq.initted and torch.isfinite(vq.embedding.weight).all() and float(vq.embedding.weight.detach().abs().sum()) > 0
    d = torch.cdist(x, vq.embedding.weight.detach())
    assert float(d.min(dim=1).values.mean()) < 2.0
