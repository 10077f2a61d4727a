"""The step in FRONT of the dgrad -> mesh path, restated in plain PyTorch for config 4 of BASELINE.json
(audio -> mel -> random-init temporal-attention network -> PCA coefficients -> K1..K5).

Not the product and not accelerated here ("the network front-end stays in PyTorch and is not the target",
BASELINE.json north_star): it exists so that the end-to-end workload can be generated, timed separately and handed to
``Reconstructor.decode_and_get_mesh`` on the device.  ``import saber`` / ``import speech_anime`` do not work in this
image (no librosa, colorama, ...: SURVEY 8c), so the pieces on the way are restated from their sources:

* windows            speech_anime/datasets/sliding_window.py:324-377 (``fetch_audio_features``): one window of
                     hop*(frames-1)+win samples per animation frame, centred on the frame, zero padded at the ends
* mel                saber/data/audio/features/spectrogram.py:66-104 (``torch.stft`` without centring, power, mel filter
                     bank, 10*log10, (db - ref_db + top_db) / top_db clipped to [0,1]); pre-emphasis misc.py:8-19;
                     parameters speech_anime/config/data/voca-dgrad.py:4-28 (8 kHz, win 512, hop 64, 128 mels, 50-3600 Hz,
                     Hamming window, ref 20 dB, top 80 dB, pre-emphasis 0.65)
* mel filter bank    ``librosa.filters.mel`` (librosa is absent; third-party, version unpinned by requirements.txt): its
                     published definition -- Slaney mel scale, triangular filters, area ("slaney") normalisation
* delta, delta-delta speech_anime/datasets/get_features.py:195-214 -> ``librosa.feature.delta(order=1|2)`` =
                     ``scipy.signal.savgol_filter(width 9, polyorder=order, deriv=order, mode="interp")`` over time,
                     restated as one [frames, frames] operator per order built with scipy (identical arithmetic)
* network            speech_anime/config/model/dgrad.py:56-100 built the way speech_anime/layers/__init__.py:37-129,
                     layers/freq_lstm.py, layers/attentions.py:41-125 and saber/nn/layers/{conv2d,linear,extend}.py
                     build it; composition speech_anime/model/model.py:18-47; one-hot speaker condition
                     modules/speaker.py:21-26.  Inference form: dropout off, batch norm in eval mode, and weight
                     normalisation (a reparametrisation: w = g v/|v|) folded into plain weights.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE, WIN, HOP, N_MELS, FMIN, FMAX = 8000, 512, 64, 128, 50.0, 3600.0
REF_DB, TOP_DB, PREEMPH, FPS, WINDOW_FRAMES = 20.0, 80.0, 0.65, 60.0, 64
SLIDING = HOP * (WINDOW_FRAMES - 1) + WIN                       # 4544 samples per animation frame


# ----------------------------------------------------------------------------------------------------- features
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mel = f / (200.0 / 3)
    log_t = f >= 1000.0
    return np.where(log_t, 15.0 + np.log(np.maximum(f, 1e-9) / 1000.0) / (np.log(6.4) / 27.0), mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), m * (200.0 / 3))


def mel_filters(sr=SAMPLE_RATE, n_fft=WIN, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """[n_mels, n_fft/2+1] float32, Slaney scale and normalisation (what misc.py:110-117 asks librosa for)."""
    fft_f = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower, upper = -ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def delta_operator(order, frames=WINDOW_FRAMES, width=9):
    """D with delta(x)[s] = sum_t D[s, t] x[t] over the time axis: savgol_filter applied to the columns of the identity
    (the filter is linear, edge handling included, so this is the same arithmetic)."""
    from scipy.signal import savgol_filter
    return savgol_filter(np.eye(frames), width, polyorder=order, deriv=order, axis=0, mode="interp").astype(np.float32)


class MelFeatures(torch.nn.Module):
    """signal [U, samples] in [-1, 1] -> audio_feat [U, n_frames, 64, 128, 3] (N,T,F,C as the model takes it,
    api.py:158: ``torch.rand(1, 64, 128, 3)``) for the animation frames 0..n_frames-1 at 60 fps."""

    def __init__(self):
        super().__init__()
        self.register_buffer("mel", torch.from_numpy(mel_filters()))
        self.register_buffer("window", torch.from_numpy(np.hamming(WIN).astype(np.float32)))
        self.register_buffer("d1", torch.from_numpy(delta_operator(1)))
        self.register_buffer("d2", torch.from_numpy(delta_operator(2)))

    @staticmethod
    def window_starts(n_frames):
        """sliding_window.py:349-351: m = floor(frame * sr / fps) (float32 like frame_to_sample), e = m + S//2, s = e - S."""
        m = np.floor((np.arange(n_frames, dtype=np.float32) * np.float32(SAMPLE_RATE) / np.float32(FPS)).astype(np.float32))
        return (m.astype(np.int64) + SLIDING // 2) - SLIDING

    def forward(self, signal, n_frames):
        U, L = signal.shape
        starts = torch.from_numpy(self.window_starts(n_frames)).to(signal.device)
        pad = SLIDING                                            # zero padding on both sides (sliding_window.py:356-363)
        sig = F.pad(signal, (pad, pad))
        idx = (starts + pad)[:, None] + torch.arange(SLIDING, device=signal.device)[None, :]     # [n_frames, SLIDING]
        idx = idx.clamp_(0, L + 2 * pad - 1)
        win = sig[:, idx]                                        # [U, n_frames, SLIDING]
        # pre-emphasis per window (misc.py:8-19): first sample kept
        win = torch.cat((win[..., :1], win[..., 1:] - PREEMPH * win[..., :-1]), dim=-1)
        fr = win.unfold(-1, WIN, HOP)                            # [U, n_frames, 64, WIN] (center=False)
        spec = torch.fft.rfft(fr * self.window, dim=-1)
        power = spec.real.square() + spec.imag.square()          # [U, n, 64, 257]
        mel = torch.matmul(power, self.mel.t())                  # [U, n, 64, 128]
        db = 10.0 * torch.log10(torch.clamp(mel, min=torch.finfo(torch.float32).eps))
        feat = torch.clamp((db - REF_DB + TOP_DB) / TOP_DB, 0.0, 1.0)
        # deltas over time (axis 2 here): x @ D.T in the reference's [feat, time] orientation
        d1 = torch.einsum("untf,st->unsf", feat, self.d1)
        d2 = torch.einsum("untf,st->unsf", feat, self.d2)
        return torch.stack((feat, d1, d2), dim=-1)               # [U, n, 64 (T), 128 (F), 3 (C)]


def band_limited_noise(n_utterances, seconds=4.0, seed=0, device="cpu", lo=80.0, hi=3400.0):
    """Synthetic speech stand-in: white noise band-passed in the FFT domain, syllable-rate amplitude
    envelope, peak-normalised into [-1, 1] like the reference asserts (sliding_window.py:330)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = int(seconds * SAMPLE_RATE)
    x = torch.randn(n_utterances, n, generator=g)
    X = torch.fft.rfft(x, dim=-1)
    f = torch.fft.rfftfreq(n, 1.0 / SAMPLE_RATE)
    X[:, (f < lo) | (f > hi)] = 0
    x = torch.fft.irfft(X, n=n, dim=-1)
    t = torch.arange(n) / SAMPLE_RATE
    rate = 2.0 + 3.0 * torch.rand(n_utterances, 1, generator=g)
    env = 0.55 + 0.45 * torch.sin(2 * math.pi * rate * t[None, :] + 6.28 * torch.rand(n_utterances, 1, generator=g))
    x = x * env
    return (x / x.abs().amax(dim=-1, keepdim=True).clamp_min(1e-9) * 0.9).to(device)


# ------------------------------------------------------------------------------------------------------ network
def _lrelu(x):
    return F.leaky_relu(x, 0.2)


class _ConvBlock(torch.nn.Module):
    """("conv2d", cin, cout, (k,1), (1,1), "act=lrelu@a:0.2", batch_norm) [+ ("pool2d","max",(2,1))]:
    "same" padding over frequency, activation BEFORE batch norm (extend.py:_ext_post_module, bn_first=False)."""

    def __init__(self, cin, cout, k, pool):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, (k, 1), padding=(k // 2, 0))
        self.bn = torch.nn.BatchNorm2d(cout, momentum=0.01, eps=1e-3)
        self.pool = pool

    def forward(self, x):
        x = self.bn(_lrelu(self.conv(x)))
        return F.max_pool2d(x, (2, 1)) if self.pool else x


class FreqLstm(torch.nn.Module):
    """layers/freq_lstm.py, mode "full": a BiLSTM along the frequency axis of every time step, then one projection."""

    def __init__(self, cin=64, freq=32, hidden=128, out=256):
        super().__init__()
        self.freq, self.out = freq, out
        self.lstm = torch.nn.LSTM(cin, hidden, num_layers=1, batch_first=True, bidirectional=True)
        self.proj = torch.nn.Linear(freq * 2 * hidden, out)

    def forward(self, x):                                         # [B, C, F, T]
        b, c, f, t = x.shape
        y, _ = self.lstm(x.permute(0, 3, 2, 1).reshape(b * t, f, c))
        return self.proj(y.reshape(b * t, -1)).view(b, t, self.out)      # [B, T, out] (the reference returns B,C,1,T and permutes back)


class BahdanauAttention(torch.nn.Module):
    """layers/attentions.py:41-125 with query_radius 2: the three centre frames become one query through a stride-3
    convolution; additive scores over all 64 frames; softmax; context = align @ values."""

    def __init__(self, size=512, units=128, radius=2, scale_score_at_eval=1.0):
        super().__init__()
        self.radius, self.scale = radius, scale_score_at_eval
        q = 2 * radius - 1
        self.conv_query = torch.nn.Conv1d(size, size, q, stride=q, bias=False)
        self.proj_key = torch.nn.Linear(size, units, bias=False)
        self.proj_qry = torch.nn.Linear(size, units, bias=False)
        self.v = torch.nn.Linear(units, 1, bias=False)
        self.b = torch.nn.Parameter(torch.zeros(1, 1, units))

    def forward(self, x):                                         # [B, T, size]
        mid = x.shape[1] // 2                                     # layers/__init__.py:95-100
        query = x[:, mid - (self.radius - 1): mid + self.radius]
        query = self.conv_query(query.transpose(1, 2)).transpose(1, 2)      # [B, 1, size]
        score = self.v(torch.tanh(self.proj_qry(query) + self.proj_key(x) + self.b)).transpose(1, 2)   # [B, 1, T]
        if not self.training:
            score = score * self.scale
        return torch.bmm(torch.softmax(score, dim=-1), x)         # [B, 1, size]


class SpeechToCoefficients(torch.nn.Module):
    """config/model/dgrad.py:56-100 in ``prediction_type = pca_coeffs`` form: audio_feat [N,64,128,3] + speaker id ->
    (coeff_scale [N,85], coeff_rotat [N,180]) -- what ``PcaInversion`` (output_module.py:115-116) would decode; here
    the decode is K1 of the CUDA path."""

    def __init__(self, num_speakers=8, k_scale=85, k_rotat=180):
        super().__init__()
        self.num_speakers = num_speakers
        self.c1, self.c2, self.c3 = _ConvBlock(3, 32, 3, True), _ConvBlock(32, 64, 3, True), _ConvBlock(64, 64, 1, False)
        self.freq_lstm = FreqLstm(64, 32, 128, 256)
        self.lstm = torch.nn.LSTM(256, 256, num_layers=2, bias=False, batch_first=True, dropout=0.1, bidirectional=True)
        self.attn = BahdanauAttention(512, 128, 2, 1.0)
        cond = 512 + num_speakers
        self.fc = torch.nn.Linear(cond, 512)
        self.scale = torch.nn.ModuleList([torch.nn.Linear(cond, 512), torch.nn.Linear(512, 256), torch.nn.Linear(256, k_scale)])
        self.rotat = torch.nn.ModuleList([torch.nn.Linear(cond, 512), torch.nn.Linear(512, 256), torch.nn.Linear(256, k_rotat)])

    def forward(self, audio_feat, speaker_id):
        cond = F.one_hot(speaker_id, self.num_speakers).to(audio_feat.dtype)      # modules/speaker.py:21-26
        x = audio_feat.permute(0, 3, 2, 1)                        # N,T,F,C -> N,C,F,T
        x = self.c3(self.c2(self.c1(x)))                          # [N, 64, 32, 64]
        x = self.freq_lstm(x)                                     # [N, 64, 256]
        x, _ = self.lstm(x)                                       # [N, 64, 512]
        z = self.attn(x)[:, 0]                                    # [N, 512]   (L = 1)
        h = _lrelu(self.fc(torch.cat((z, cond), dim=-1)))         # "cat_condition=2" (layers/__init__.py:62-77)

        def branch(layers):
            y = _lrelu(layers[0](torch.cat((h, cond), dim=-1)))
            return layers[2](torch.tanh(layers[1](y)))

        return branch(self.scale), branch(self.rotat)


def build_network(seed=0, device="cuda", coeff_gain=1.0):
    """Random-init network in eval mode (no checkpoints are shipped: README.md:22,43-46).  ``coeff_gain`` scales the last
    layers so that the coefficients have roughly unit spread like config 2's."""
    torch.manual_seed(seed)
    net = SpeechToCoefficients()
    with torch.no_grad():
        for last in (net.scale[2], net.rotat[2]):
            last.weight.mul_(coeff_gain)
    return net.to(device).eval()
