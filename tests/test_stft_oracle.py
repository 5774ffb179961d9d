"""Known-answer tests of the (builder-defined, parity-unpinned) IQ -> spectrogram specification."""
import numpy as np

from oracle import stft_ref


def test_tone_peaks_at_its_bin_and_reads_0_dbfs():
    n, k = 1 << 14, 100                       # e^{+j 2 pi k n / 1024}: bin k, image row k + 512 after fftshift
    iq = np.exp(2j * np.pi * k * np.arange(n) / 1024)
    p = stft_ref.stft_power(iq)
    assert p.shape == (1024, 1 + (n - 1024) // 256)
    assert np.all(p.argmax(0) == 512 + k)
    v = stft_ref.normalise_db(p, 1024, -100.0, 0.0)
    assert np.allclose(v[512 + k], 1.0)       # unit-amplitude tone == 0 dBFS == top of the range
    neg = stft_ref.stft_power(np.conj(iq))
    assert np.all(neg.argmax(0) == 512 - k)


def test_impulse_is_flat():
    iq = np.zeros(4096, dtype=np.complex128)
    iq[700] = 1.0                              # inside frame 0 only partly; frame 0 covers [0,1024)
    p = stft_ref.stft_power(iq)[:, 0]
    w = stft_ref.hann_periodic(1024)[700] ** 2
    assert np.allclose(p, w)


def test_parseval_per_frame():
    rng = np.random.default_rng(1)
    iq = rng.normal(size=8192) + 1j * rng.normal(size=8192)
    p = stft_ref.stft_power(iq)
    w = stft_ref.hann_periodic(1024)
    for t in range(p.shape[1]):
        seg = iq[t * 256: t * 256 + 1024] * w
        assert np.isclose(p[:, t].sum(), 1024 * np.sum(np.abs(seg) ** 2))


def test_letterbox_layout_north_star_geometry():
    rng = np.random.default_rng(2)
    iq = (rng.normal(size=(1, 1 << 20)) + 1j * rng.normal(size=(1, 1 << 20))) * 0.05
    img = stft_ref.iq_to_letterbox(iq)
    assert img.shape == (1, 3, 640, 640)
    assert np.all(img[0, :, :240] == 114 / 255) and np.all(img[0, :, 400:] == 114 / 255)   # 160-row band
    assert np.array_equal(img[0, 0], img[0, 1]) and np.array_equal(img[0, 0], img[0, 2])
    assert 0.0 <= img[0, 0, 240:400].min() and img[0, 0, 240:400].max() <= 1.0


def test_stft_power_matches_torch_and_scipy():
    """The oracle's STFT convention (periodic Hann, hop 256, center=False, fftshift along frequency) against two independent
    library implementations on the same samples: `torch.stft` and `scipy.signal.ShortTimeFFT`-free `scipy.fft` framing."""
    import scipy.fft
    import torch

    rng = np.random.default_rng(5)
    iq = (rng.standard_normal(9000) + 1j * rng.standard_normal(9000)).astype(np.complex64)
    p = stft_ref.stft_power(iq)                                        # [1024, T]
    T = 1 + (iq.size - 1024) // 256
    assert p.shape == (1024, T)
    w = torch.hann_window(1024, periodic=True, dtype=torch.float64)
    X = torch.stft(torch.from_numpy(iq.astype(np.complex128)), n_fft=1024, hop_length=256, win_length=1024, window=w,
                   center=False, onesided=False, return_complex=True)  # [1024, T], bins 0..1023
    pt = torch.fft.fftshift(X.abs() ** 2, dim=0).numpy()
    assert np.allclose(p, pt, rtol=1e-9, atol=1e-9 * pt.max())
    frames = np.stack([iq[t * 256: t * 256 + 1024].astype(np.complex128) * w.numpy() for t in range(T)], 1)
    ps = np.abs(scipy.fft.fftshift(scipy.fft.fft(frames, axis=0), axes=0)) ** 2
    assert np.allclose(p, ps, rtol=1e-9, atol=1e-9 * ps.max())
