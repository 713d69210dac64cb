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


def _sharded_worker(rank, world, port, ngpu, q):
    """One rank of score_sharded on a real engine: NCCL when every rank has its own GPU, else both ranks share
    GPU 0 and the records travel over gloo (the data path is the same: partition -> score -> gather -> un-permute)."""
    import os
    import torch
    import torch.distributed as dist
    from nele_gan_b200 import shard
    from nele_gan_b200.engine import Engine
    from nele_gan_b200.synth import make_pair
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    own_gpu = ngpu >= world
    dev = rank if own_gpu else 0
    torch.cuda.set_device(dev)
    if own_gpu:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = [32001, 47999, 56789, 40411, 35555, 61003, 33536]
    pairs = [make_pair(400 + i, L)[:2] for i, L in enumerate(lens)]
    refs, degs = [p[0] for p in pairs], [p[1] for p in pairs]
    eng = Engine(dev)
    full = shard.score_sharded(lambda a, b: eng.score_batch(a, b, mapped=False, no_dither=True), refs, degs,
                               device=torch.device("cuda", dev) if own_gpu else None)
    if rank == 0:
        want = shard.pack_records(eng.score_batch(refs, degs, mapped=False, no_dither=True))
        q.put((full.tolist(), want.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_score_sharded_two_ranks_on_real_engines_restores_input_order():
    """BASELINE configs[4]'s data path at world size 2: the length-sorted deal, one engine per rank, the gather and
    the un-permute must give exactly what one engine gives for the whole ragged batch, in input order."""
    import socket
    import torch
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ngpu = torch.cuda.device_count()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, ngpu, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, want = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    full, want = np.asarray(full), np.asarray(want)
    assert full.shape == want.shape == (7, 14)
    assert np.array_equal(full[:, 13], want[:, 13])                            # status words
    assert np.allclose(full[:, :13], want[:, :13], rtol=1e-6, atol=1e-9)       # batch composition moves FP32 sums by ~1e-9


def test_pcm16_host_inputs_are_bit_identical_to_the_float_path(eng):
    """nele_score_batch_pcm16: clean / enhanced / noise as the int16 samples of the reference's WAV files, deg formed
    on the device -- the same scores as the float call on librosa-style conversions, with and without a prefetch."""
    from nele_gan_b200.engine import pack
    from nele_gan_b200.synth import make_pair
    lens = [33536, 34048, 40111, 16001]
    c16, e16, n16 = [], [], []
    for i, L in enumerate(lens):
        x, y, _ = make_pair(500 + i, L)
        c = np.clip(np.round(x * 32768 * 8), -32768, 32767).astype(np.int16)         # RMS 0.03 * 8: uses the 16-bit range
        nz = np.clip(np.round((y - x) * 32768 * 8), -32768, 32767).astype(np.int16)
        en = np.clip(np.round(0.8 * x * 32768 * 8), -32768, 32767).astype(np.int16)
        c16.append(c), e16.append(en), n16.append(nz)
    refs = [c.astype(np.float32) / 32768.0 for c in c16]
    degs = [e.astype(np.float32) / 32768.0 + n.astype(np.float32) / 32768.0 for e, n in zip(e16, n16)]
    want = eng.score_batch(refs, degs, mapped=False, no_dither=True)
    padded = (np.array(lens, dtype=np.int64) + 7) // 8 * 8
    offs = np.concatenate(([0], np.cumsum(padded)[:-1])).astype(np.int64)
    flat = [np.zeros(int(padded.sum()), np.int16) for _ in range(3)]
    for k, arrs in enumerate((c16, e16, n16)):
        for a, o in zip(arrs, offs):
            flat[k][o:o + len(a)] = a
    got = eng.score_packed_pcm16(flat[0], flat[1], flat[2], offs, np.array(lens, np.int32), mapped=False, no_dither=True)
    same = lambda u, v: np.allclose(u, v, rtol=1e-7, atol=0, equal_nan=True)     # run-to-run noise of FP32 atomics ~1e-9
    assert same(got.scores, want.scores) and np.array_equal(got.status, want.status)
    eng.prefetch_pcm16(flat[0], flat[1], flat[2], offs, np.array(lens, np.int32))
    again = eng.score_packed_pcm16(flat[0], flat[1], flat[2], offs, np.array(lens, np.int32), mapped=False, no_dither=True)
    assert same(again.scores, want.scores)
