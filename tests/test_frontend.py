"""CPU tier: the config-4 front-end (deformation/frontend.py) -- feature windows, mel, deltas and the network's shapes --
against straightforward per-window restatements of the reference's loops (sliding_window.py:324-377,
spectrogram.py:66-104, get_features.py:195-214)."""
import numpy as np
import torch

from deformation import frontend as FE


def _window_reference(signal, n_frames):
    """fetch_audio_features' loop (sliding_window.py:345-366) for frames 0..n_frames-1."""
    out = []
    for idx in range(n_frames):
        m = int(np.floor(np.float32(float(idx * FE.SAMPLE_RATE) / float(FE.FPS))))
        e = m + FE.SLIDING // 2
        s = e - FE.SLIDING
        part = signal[max(0, s): min(len(signal), e)]
        if len(part) == 0:
            part = np.zeros(FE.SLIDING, np.float32)
        elif s < 0:
            part = np.pad(part, [[-s, 0]], "constant")
        elif e > len(signal):
            part = np.pad(part, [[0, e - len(signal)]], "constant")
        assert len(part) == FE.SLIDING
        out.append(part)
    return np.stack(out)


def test_features_match_per_window_loop():
    from scipy.signal import savgol_filter
    sig = FE.band_limited_noise(2, seconds=1.0, seed=3)
    assert sig.shape == (2, 8000) and float(sig.abs().max()) <= 1.0
    n_frames = 60
    feats = FE.MelFeatures()(sig, n_frames)
    assert feats.shape == (2, n_frames, 64, 128, 3) and feats.dtype == torch.float32
    mel_fb = FE.mel_filters()
    assert mel_fb.shape == (128, 257) and (mel_fb >= 0).all() and (mel_fb.sum(1) > 0).all()
    ham = np.hamming(FE.WIN).astype(np.float32)
    for u in (0, 1):
        wins = _window_reference(sig[u].numpy(), n_frames)
        for fi in (0, 1, 17, 34, 59):                        # the first ones are zero padded on the left
            w = wins[fi]
            w = np.append(w[0], w[1:] - FE.PREEMPH * w[:-1])                      # misc.py:17
            spec = torch.stft(torch.from_numpy(w.astype(np.float32)), n_fft=FE.WIN, hop_length=FE.HOP, win_length=FE.WIN,
                              window=torch.from_numpy(ham), center=False, normalized=False, onesided=True,
                              return_complex=True)                                # spectrogram.py:83-94
            power = (spec.real ** 2 + spec.imag ** 2).numpy()                     # [257, 64]
            db = 10.0 * np.log10(np.maximum(mel_fb @ power, np.finfo(np.float32).eps))
            feat = np.clip((db - FE.REF_DB + FE.TOP_DB) / FE.TOP_DB, 0.0, 1.0)    # [128, 64] = [feat, time]
            d1 = savgol_filter(feat, 9, polyorder=1, deriv=1, axis=-1, mode="interp")
            d2 = savgol_filter(feat, 9, polyorder=2, deriv=2, axis=-1, mode="interp")
            want = np.stack((feat, d1, d2), axis=0).transpose(2, 1, 0)            # C,F,T -> T,F,C
            got = feats[u, fi].numpy()
            assert np.abs(got[..., 0] - want[..., 0]).max() < 2e-4               # log of fp32 spectra
            assert np.abs(got[..., 1:] - want[..., 1:]).max() < 2e-4


def test_network_shapes_and_condition():
    net = FE.build_network(seed=1, device="cpu")
    x = torch.rand(3, 64, 128, 3)
    with torch.no_grad():
        s0, r0 = net(x, torch.zeros(3, dtype=torch.long))
        s1, r1 = net(x, torch.full((3,), 5, dtype=torch.long))
        s2, _ = net(x, torch.zeros(3, dtype=torch.long))
    assert s0.shape == (3, 85) and r0.shape == (3, 180)
    assert torch.equal(s0, s2)                                # eval mode: deterministic (dropout off)
    assert not torch.allclose(s0, s1)                         # the speaker one-hot reaches the output branches
    n_params = sum(p.numel() for p in net.parameters())
    assert 5_000_000 < n_params < 9_000_000                   # 8192x256 projection + BiLSTMs dominate
