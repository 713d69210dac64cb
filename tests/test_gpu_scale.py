"""Full-size behaviour of the engine (BASELINE.json configs[2]: 4096 x 3 s) checked through
size-independent properties: the oracle cannot score thousands of pairs in test time, but copies
of a pair must score identically wherever they sit in the batch, a batch must equal its single-pair
calls, and the chunking (> 4096 pairs per call, > 1024 pairs per SIIB sub-chunk) must be invisible."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from nele_gan_b200.engine import Engine
    return Engine(0)


def test_bench_size_batch_copies_agree_and_match_single_calls(eng):
    from nele_gan_b200.synth import make_batch
    n, uniq = 4096, 64
    refs, degs = make_batch(n, 48000, seed=777_000, unique=uniq)
    r = eng.score_batch(refs, degs, mapped=False, no_dither=True)
    assert r.ok.all() and r.siib_nullspace_dropped.all()      # 48000 = 240 hops: periodic tiling, flagged
    s = r.scores.reshape(n // uniq, uniq, 3)
    spread = np.abs(s - s[0]).max(axis=0)                 # copies of the same pair, 64 places each
    assert spread[:, 0].max() < 1e-6 * np.abs(s[0][:, 0]).max()   # SIIB
    assert spread[:, 1].max() < 1e-9                               # HASPI (double atomics reorder sums)
    assert spread[:, 2].max() == 0.0                               # ESTOI is deterministic
    one = eng.score_batch(refs[:4], degs[:4], mapped=False, no_dither=True)
    assert np.allclose(one.scores, r.scores[:4], rtol=1e-6, atol=1e-9)
    # all three metrics react to the SNR the synthetic pairs were built with: not a constant output
    assert np.ptp(r.scores[:uniq, 0]) > 10 and np.ptp(r.scores[:uniq, 1]) > 1 and np.ptp(r.scores[:uniq, 2]) > 0.1


def test_more_pairs_than_one_chunk(eng):
    from nele_gan_b200.synth import make_batch
    n = 4096 + 300                                         # forces a second chunk and SIIB sub-chunks
    lens = np.where(np.arange(n) % 3 == 0, 16000, 17333)
    refs, degs = make_batch(n, lens, seed=888_000, unique=48)
    r = eng.score_batch(refs, degs, mapped=True, no_dither=True)
    assert r.scores.shape == (n, 3) and np.isfinite(r.scores).all()
    idx = [0, 1, 47, 4095, 4096, 4097, n - 1]
    one = eng.score_batch([refs[i] for i in idx], [degs[i] for i in idx], mapped=True, no_dither=True)
    assert np.allclose(one.scores, r.scores[idx], rtol=1e-6, atol=1e-9)
    # copies (unique = 48, period 48 with the same length pattern every 3) agree across the chunk boundary
    assert np.allclose(r.scores[48 * 3: 48 * 6], r.scores[48 * 3 + 144 * 28: 48 * 6 + 144 * 28], rtol=1e-6, atol=1e-9)


def test_permutation_invariance_ragged(eng):
    from nele_gan_b200.synth import make_pair
    rng = np.random.default_rng(11)
    lens = rng.integers(16000 * 2, 16000 * 5, size=24)
    pairs = [make_pair(300 + i, int(L))[:2] for i, L in enumerate(lens)]
    a = eng.score_batch([p[0] for p in pairs], [p[1] for p in pairs], mapped=False, no_dither=True)
    perm = rng.permutation(len(pairs))
    b = eng.score_batch([pairs[i][0] for i in perm], [pairs[i][1] for i in perm], mapped=False, no_dither=True)
    assert np.allclose(b.scores, a.scores[perm], rtol=1e-6, atol=1e-9)
    assert np.array_equal(b.status, a.status[perm])


def test_argument_errors(eng):
    from nele_gan_b200.engine import NeleError
    x = np.zeros(1000, np.float32)
    with pytest.raises(ValueError):
        eng.score_batch([x], [x, x])
    with pytest.raises(ValueError):
        eng.score_batch([x[:0]], [x[:0]])
    with pytest.raises(NeleError, match="bad metric mask"):
        eng.score_packed(x, x, np.array([0]), np.array([1000]), metrics=0)
    with pytest.raises(NeleError, match="fs"):
        eng.score_packed(x, x, np.array([0]), np.array([1000]), fs=0)
    r = eng.score_batch([], [])
    assert r.scores.shape == (0, 3)


def test_prefetch_gives_identical_results_and_mismatches_are_ignored(eng):
    """nele_prefetch only moves the upload earlier: same scores with, without, and with a
    prefetch that does not match the following call."""
    from nele_gan_b200.engine import pack
    from nele_gan_b200.synth import make_batch
    refs, degs = make_batch(6, [30001, 24000, 33536, 28111, 16000, 40007])
    fr, offs, lens = pack(refs)
    fd, _, _ = pack(degs)
    base = eng.score_packed(fr, fd, offs, lens, mapped=False, no_dither=True)
    eng.prefetch(fr, fd, offs, lens)
    eng.prefetch(fr, fd, offs, lens)                                   # two calls ahead
    a = eng.score_packed(fr, fd, offs, lens, mapped=False, no_dither=True)
    b = eng.score_packed(fr, fd, offs, lens, mapped=False, no_dither=True)
    same = lambda u, v: np.allclose(u, v, rtol=1e-7, atol=0, equal_nan=True)   # run-to-run noise of the FP32 eigen-solver is ~1e-9
    assert same(a.scores, base.scores) and same(b.scores, base.scores)
    fr2, offs2, lens2 = pack(refs[:3])
    fd2, _, _ = pack(degs[:3])
    eng.prefetch(fr, fd, offs, lens)                                   # stale: a different batch follows
    c = eng.score_packed(fr2, fd2, offs2, lens2, mapped=False, no_dither=True)
    assert same(c.scores, base.scores[:3])
    d = eng.score_packed(fr, fd, offs, lens, mapped=False, no_dither=True)
    assert same(d.scores, base.scores)
