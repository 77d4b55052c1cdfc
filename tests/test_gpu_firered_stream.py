"""FireRed Stream-VAD and AED on the device: the cache-carrying graph, the streaming segmenter and the
three-event AED section against the UNMODIFIED reference script's record (tests/golden/firered_script.npz,
tests/golden/firered.npz) and against the oracle on ragged multi-stream inputs."""
import os

import numpy as np
import pytest
import torch

import vadx
from vadx import audio_io, firered_vad, postprocess as PP, synth, weights as W
from oracle import postproc as OP
from oracle.firered import FireRedOracle

pytestmark = pytest.mark.gpu
TOL = 1e-3  # BASELINE.json: max abs err of frame probabilities
BOUND = 3e-4  # what this path is ASSERTED to: tcgen05 three-product bf16 split, measured <= 1.9e-4; TOL stays the contract and the decision-margin test


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "firered.npz"))


@pytest.fixture(scope="module")
def gold_script(golden_dir):
    return np.load(os.path.join(golden_dir, "firered_script.npz"))


@pytest.fixture(scope="module")
def stream_session(cuda):
    cfg = W.FireRedConfig(N2=0, S2=0, streaming=True)
    return vadx.FireRedStreamSession(W.firered_random_init(cfg, 5), cfg)


def _pairs(ts):
    return np.array(ts, np.float64).reshape(-1, 2)


# ---------------------------------------------------------------- streaming segmenter (bit-exact)
@pytest.mark.parametrize("i", range(6))
def test_stream_post_whole_and_chunked(cuda, gold_script, i):
    g = gold_script
    p = g[f"sp{i}_probs"]
    ws, thr, pad, msp, mxs, msi, split = g[f"sp{i}_params"]
    prm = (int(ws), thr, int(pad), int(msp), int(mxs), int(msi))
    whole = PP.StreamVadPostprocessor(*prm).process_batch(p)
    assert np.array_equal(_pairs(whole), g[f"sp{i}_whole"])
    if int(split):
        pp = PP.StreamVadPostprocessor(*prm)
        calls = [pp.process_batch(p[j:j + int(split)]) for j in range(0, len(p), int(split))]
        assert [len(c) for c in calls] == g[f"sp{i}_chunked_counts"].tolist()
        assert np.array_equal(_pairs([t for c in calls for t in c]), g[f"sp{i}_chunked"])
        pp.reset()
        assert np.array_equal(_pairs(pp.process_batch(p)), g[f"sp{i}_whole"])


def test_stream_post_many_ragged_streams(cuda):
    """257 streams in lock-step, 14-frame calls, ragged ends; each must equal its own sequential oracle."""
    rs = np.random.RandomState(3)
    S, n_calls, T = 257, 40, 14
    gate = (np.sin(np.arange(n_calls * T)[None, :] / rs.uniform(10, 50, size=(S, 1)) + rs.uniform(0, 6, size=(S, 1))) > 0)
    p = np.clip(0.1 + 0.8 * gate + rs.normal(0, 0.08, size=gate.shape), 0, 1).astype(np.float32)
    n_total = rs.randint(0, n_calls * T + 1, size=S)
    n_total[:3] = (0, 1, n_calls * T)
    pp = PP.StreamVadPostprocessor(5, 0.4, 5, 8, 60, 20, n_streams=S, device=cuda)
    d = torch.from_numpy(p).to(cuda)
    for k in range(n_calls):
        nf = np.clip(n_total - k * T, 0, T).astype(np.int32)
        pp.feed(d[:, k * T:(k + 1) * T], torch.from_numpy(nf).to(cuda))
    got = pp.timestamps()
    n_seg = 0
    for s in range(S):
        ref = OP.StreamPost(5, 0.4, 5, 8, 60, 20).feed(p[s, :n_total[s]])
        assert got[s] == ref, s
        n_seg += len(ref)
    assert n_seg > S


def test_stream_post_rejects_bad_inputs(cuda):
    pp = PP.StreamVadPostprocessor(5, 0.4, 5, 8, 2000, 20, n_streams=2, device=cuda)
    with pytest.raises(ValueError):
        pp.feed(torch.zeros((3, 14), device=cuda))
    with pytest.raises(ValueError):
        pp.feed(torch.zeros((2, 14), device=cuda, dtype=torch.float64))
    with pytest.raises(ValueError):
        PP.StreamVadPostprocessor(65, 0.4, 5, 8, 2000, 20, device=cuda).process_batch(np.zeros(4, np.float32))


# ---------------------------------------------------------------- cache-carrying graph
def test_stream_graph_cache_carry_ort_surface(gold, stream_session):
    """Four 2560-sample calls through the ORT-shaped run(), caches as numpy in/out (golden: reference module)."""
    sess = stream_session
    cfg = sess.cfg
    names_in = [i.name for i in sess.get_inputs()]
    names_out = [o.name for o in sess.get_outputs()]
    assert names_in == ["audio", "caches_in"] and names_out == ["probs", "caches_out"]
    assert sess._inputs_meta[1].shape == [cfg.R, 1, cfg.P, 19]
    chunks = synth.synth_streams(6, 16000, seed=1234)
    caches = np.zeros((cfg.R, 1, cfg.P, 19), np.float32)
    for i in range(4):
        a = chunks[2][i * 2560:(i + 1) * 2560].reshape(1, 1, -1)
        p, caches = sess.run(names_out, {"audio": a, "caches_in": caches})
        assert p.shape == (1, 1, 14)
        assert np.abs(p[0, 0] - gold["stream_probs"][i]).max() <= BOUND
    # caches are un-normalised activations (|v| ~ 10): bound them relative to their scale
    ref_c = gold["stream_caches_last"]
    assert np.abs(caches[:, 0, ::16, :] - ref_c).max() <= BOUND * max(1.0, float(np.abs(ref_c).max()))
    with pytest.raises(ValueError):
        sess.run(names_out, {"audio": a})
    with pytest.raises(ValueError):
        sess.run(names_out, {"audio": a, "caches_in": caches[:, :, :, :5]})
    with pytest.raises(ValueError):
        sess.run(["nope"], {"audio": a, "caches_in": caches})


def test_stream_graph_fp32_path_matches_tighter(gold, cuda):
    cfg = W.FireRedConfig(N2=0, S2=0, streaming=True)
    sess = vadx.FireRedStreamSession(W.firered_random_init(cfg, 5), cfg, tensor_cores=False)
    chunks = synth.synth_streams(6, 16000, seed=1234)
    c = sess.new_caches(1, cuda)
    for i in range(4):
        a = torch.from_numpy(chunks[2][i * 2560:(i + 1) * 2560].copy()).view(1, -1).to(cuda)
        p, c = sess.run_batch(a, c)
        assert np.abs(p.cpu().numpy()[0, 0] - gold["stream_probs"][i]).max() <= 5e-5


def test_vad_sample_stream_section(gold_script, golden_dir, stream_session, measured):
    audio = np.load(os.path.join(golden_dir, "vad_sample_16k.npz"))["audio"]
    r = firered_vad.run_stream_vad(audio, stream_session)
    ref_p = gold_script["stream_probs"]
    assert r.probs.shape == ref_p.shape == (489,)
    err = np.abs(r.probs - ref_p).max()
    measured("firered_stream: stream-VAD vad_sample.wav: max abs err", err, BOUND)
    ref_c = gold_script["stream_caches_last"]
    assert np.abs(r.caches.cpu().numpy()[:, 0, ::16, :] - ref_c).max() <= BOUND * max(1.0, float(np.abs(ref_c).max()))
    # the segmenter on the device probabilities equals the oracle on those same probabilities ...
    assert r.timestamps == OP.StreamPost(5, 0.4, 5, 8, 2000, 20).feed(r.probs)
    # ... and the script's record wherever no smoothed probability sits within TOL of the threshold
    sm = OP.smooth_probs(ref_p, 5)
    ring = np.array([ref_p[max(0, t - 4):t + 1].mean() for t in range(len(ref_p))])
    if min(np.abs(sm - 0.4).min(), np.abs(ring - 0.4).min()) > TOL:
        assert np.array_equal(_pairs(r.timestamps), gold_script["stream_timestamps"])


def test_stream_many_ragged_streams_against_oracle(cuda, stream_session):
    """24 streams of different lengths in lock-step; each stream's frames and segments must equal the
    reference loop on that stream alone (restated by the oracle)."""
    S, c = 24, 2560
    lengths = [c * 9 - 331 * s for s in range(S)]
    lengths[-1] = 300                      # shorter than one frame: no output at all
    lengths[-2] = c * 3 + 250              # last chunk zero-padded to one frame
    raw = synth.synth_streams(S, c * 9, seed=31)
    padded = np.zeros((S, c * 9), np.int16)
    for s in range(S):
        padded[s, :lengths[s]] = raw[s, :lengths[s]]
    probs, post, _, plan = firered_vad.run_stream_vad_streams(stream_session, torch.from_numpy(padded).to(cuda), lengths)
    got_ts = post.timestamps()
    probs = probs.cpu().numpy().reshape(S, -1, 14)
    cfg = stream_session.cfg
    orc = FireRedOracle(W.firered_random_init(cfg, 5), cfg)
    worst, n_seg = 0.0, 0
    for s in range(0, S, 3):
        caches = torch.zeros(cfg.R, 1, cfg.P, 19)
        ref = []
        for pos in range(0, lengths[s], c):
            ch = raw[s, pos:min(pos + c, lengths[s])]
            if len(ch) < 400:
                ch = np.pad(ch, (0, 400 - len(ch)))
            p, caches = orc.forward(torch.from_numpy(np.ascontiguousarray(ch)).view(1, 1, -1), caches)
            ref.append(p.numpy()[0, 0])
        ref = np.concatenate(ref)[:firered_vad.valid_frame_count(lengths[s])] if ref else np.zeros(0, np.float32)
        mine = np.concatenate([probs[s, k, :plan[k, s]] for k in range(plan.shape[0])])
        assert mine.shape == ref.shape, s
        if len(ref):
            worst = max(worst, float(np.abs(mine - ref).max()))
        assert got_ts[s] == OP.StreamPost(5, 0.4, 5, 8, 2000, 20).feed(mine), s
        n_seg += len(got_ts[s])
    print(f"ragged lock-step streams: max abs err {worst:.2e}, segments {n_seg}")
    assert worst <= BOUND and n_seg > 0
    assert got_ts[-1] == []


# ---------------------------------------------------------------- AED section
def test_vad_sample_vad_and_aed_sections(cuda, gold_script, golden_dir):
    """RUN_VAD then RUN_AED of the script: noise tails from one generator, three event tracks."""
    audio = np.load(os.path.join(golden_dir, "vad_sample_16k.npz"))["audio"]
    rng = np.random.RandomState(1234)
    cfg_v, cfg_a = W.FireRedConfig(), W.FireRedConfig(odim=3)
    vad = vadx.FireRedSession(W.firered_random_init(cfg_v, 0), cfg_v, chunk_len=16000)
    aed = vadx.FireRedSession(W.firered_random_init(cfg_a, 2), cfg_a, chunk_len=16000)
    rv = firered_vad.run_vad(audio, vad, rng=rng)
    ra = firered_vad.run_aed(audio, aed, rng=rng)
    g = gold_script
    err_v, err_a = np.abs(rv.probs - g["vad_probs"]).max(), np.abs(ra.probs - g["aed_probs"]).max()
    print(f"script sections: VAD max abs err {err_v:.2e}, AED {err_a:.2e}")
    assert err_v <= BOUND and err_a <= BOUND
    if np.abs(OP.smooth_probs(g["vad_probs"], 5) - np.float32(0.4)).min() > TOL:
        assert np.array_equal(_pairs(rv.timestamps), g["vad_timestamps"])
    for e, (ev, thr) in enumerate((("speech", 0.4), ("singing", 0.5), ("music", 0.5))):
        # device state machine on device probabilities == oracle on the same probabilities
        d = OP.frame_decisions(ra.probs[e], 5, thr, 20, 2000, 20, 5, 0)
        assert ra.event2timestamps[ev] == OP.segments_from_decisions(d, 0.01, 0.025, len(audio) / 16000, True)
        if np.abs(OP.smooth_probs(g["aed_probs"][e], 5) - np.float32(thr)).min() > TOL:
            assert np.array_equal(_pairs(ra.event2timestamps[ev]), g[f"aed_{ev}_timestamps"]), ev
        if np.abs(g["aed_probs"][e] - np.float32(thr)).min() > TOL:
            assert ra.event2ratio[ev] == float(g[f"aed_{ev}_ratio"])


def test_stream_cuda_graph_replay_is_bit_identical(cuda, stream_session):
    """Eager lock-step run vs the two captured steps replayed; twice, so the second call reuses the cached graphs."""
    S, c, n_calls = 40, 2560, 7
    lengths = [c * n_calls - 97 * s for s in range(S)]
    raw = synth.synth_streams(S, c * n_calls, seed=8)
    padded = np.zeros((S, c * n_calls), np.int16)
    for s in range(S):
        padded[s, :lengths[s]] = raw[s, :lengths[s]]
    d = torch.from_numpy(padded).to(cuda)
    p0, post0, c0, _ = firered_vad.run_stream_vad_streams(stream_session, d, lengths)
    ts0 = post0.timestamps()
    for _ in range(2):
        p1, post1, c1, _ = firered_vad.run_stream_vad_streams(stream_session, d, lengths, graph=True)
        assert torch.equal(p0, p1) and torch.equal(c0, c1)
        assert post1.timestamps() == ts0


def test_aed_many_streams_against_oracle(cuda):
    """run_aed_streams: S streams x 3 chunks, three event tracks, each post-processed on the device with its own threshold."""
    cfg = W.FireRedConfig(odim=3)
    wts = W.firered_random_init(cfg, 2)
    sess = vadx.FireRedSession(wts, cfg, chunk_len=16000)
    S, n_chunks = 20, 3
    audio = synth.synth_streams(S, n_chunks * 16000, seed=61)
    lengths = [n_chunks * 16000 - 211 * s for s in range(S)]
    d = torch.from_numpy(audio.reshape(S, n_chunks, 16000)).to(cuda)
    probs, per_event, n_valid = firered_vad.run_aed_streams(sess, d, lengths)
    ref = FireRedOracle(wts, cfg).forward(audio.reshape(S * n_chunks, 16000)).numpy().reshape(S, n_chunks, 3, 98)
    ref = ref.transpose(0, 2, 1, 3).reshape(S, 3, n_chunks * 98)
    assert np.abs(probs.cpu().numpy() - ref).max() <= BOUND
    nv = n_valid.cpu().numpy()
    p_host = probs.cpu().numpy()
    for e, (post, dec, cnt, seg) in enumerate(per_event):
        dec = dec.cpu().numpy()
        for s in range(0, S, 4):
            want = OP.frame_decisions(p_host[s, e, :nv[s]], 5, post.prob_threshold, 20, 2000, 20, 5, 0)
            assert np.array_equal(dec[s, :nv[s]], want), (e, s)
