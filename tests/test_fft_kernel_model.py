"""CPU model of the index arithmetic of csrc/fir_fft.cu (4096 = 16 x 16 x 16 overlap-save frames): the three radix-16
passes with their inter-pass twiddles, the digit-reversed position of every spectrum bin, the per-thread table layout
the plan builder writes, and the frame bookkeeping -- restated in numpy and pinned against numpy.fft / np.convolve.
It documents (and guards) the layout conventions the kernel and the host-side table builder must agree on; the kernel
itself is checked against the oracle in tests/test_round2_features.py (-m gpu)."""
import numpy as np

N = 4096


def dft16(v, inverse=False):
    """16-point DFT along the last axis, natural order in and out (the kernel's fft16 leaves output k at register
    4 (k & 3) + (k >> 2); that permutation is register naming, not data movement)"""
    return np.fft.ifft(v, axis=-1) * 16 if inverse else np.fft.fft(v, axis=-1)


def spectrum_table(taps):
    """what fft_build_tables stores: H[k] / 4096 for k = k0 + 16 k1 + 256 k2 at [k2 >> 1][k0 * 16 + k1][k2 & 1]"""
    H = np.fft.fft(np.concatenate([taps, np.zeros(N - len(taps))])) / N
    tab = np.zeros((8, 256, 2), dtype=np.complex128)
    for k in range(N):
        k0, k1, k2 = k & 15, (k >> 4) & 15, k >> 8
        tab[k2 >> 1, k0 * 16 + k1, k2 & 1] = H[k]
    return tab


def frame_model(x, tab):
    """one frame exactly as the kernel walks it; x: 4096 complex samples, returns the 4096-point circular convolution"""
    W = lambda e, n: np.exp(-2j * np.pi * e / n)
    t = np.arange(256)
    sm = np.zeros(N, dtype=np.complex128)
    # forward pass 1 (over n2, stride 256): thread t holds x[t + 256 j]; output k0 times W_4096^(k0 t) -> sm[k0 * 256 + t]
    v = dft16(x.reshape(16, 256).T)                                    # [t][k0]
    for k0 in range(16):
        sm[k0 * 256 + t] = v[:, k0] * W(k0 * t, N)
    # forward pass 2 (over n1, stride 16) inside block k0: thread (k0, n0); output k1 times W_256^(k1 n0), in place
    blk = sm.reshape(16, 16, 16)                                       # [k0][n1][n0]
    v = dft16(np.transpose(blk, (0, 2, 1)))                            # [k0][n0][k1]
    n0 = np.arange(16)
    for k1 in range(16):
        blk[:, k1, :] = v[:, :, k1] * W(k1 * n0, 256)[None, :]
    # forward pass 3 (over n0): thread t = k0 * 16 + k1 holds sm[t * 16 + n0]; output k2 is bin k0 + 16 k1 + 256 k2
    spec = dft16(sm.reshape(256, 16))                                  # [t][k2]
    for k2 in range(16):
        spec[:, k2] *= tab[k2 >> 1, :, k2 & 1]                         # the thread's own 16 table values
    # inverse pass 1 (over k2) on the same registers; output n0 times conj(W_256^(k1 n0)), k1 = t & 15
    u = dft16(spec, inverse=True)                                      # [t][n0]
    k1 = t & 15
    for n in range(16):
        sm.reshape(256, 16)[:, n] = u[:, n] * np.conj(W(k1 * n, 256))
    # inverse pass 2 (over k1) inside block k0: thread (k0, n0) reads sm[k0 * 256 + k1 * 16 + n0], writes n1 back
    blk = sm.reshape(16, 16, 16)                                       # [k0][k1][n0]
    v = dft16(np.transpose(blk, (0, 2, 1)), inverse=True)              # [k0][n0][n1]
    for n1 in range(16):
        blk[:, n1, :] = v[:, :, n1]
    # inverse pass 3 (over k0, stride 256): the twiddle conj(W_4096^(k0 t)) applied on the input side, t = 16 n1 + n0
    w = np.stack([sm[k0 * 256 + t] * np.conj(W(k0 * t, N)) for k0 in range(16)], axis=1)     # [t][k0]
    y = dft16(w, inverse=True)                                         # [t][n2] -> position n2 * 256 + t
    return y.T.reshape(N)


def test_frame_model_is_a_circular_convolution():
    rng = np.random.default_rng(0)
    for K in (2, 33, 257, 1024, 2049):
        taps = rng.standard_normal(K)
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        y = frame_model(x, spectrum_table(taps))
        ref = np.fft.ifft(np.fft.fft(x) * np.fft.fft(np.concatenate([taps, np.zeros(N - K)])))
        assert np.abs(y - ref).max() <= 1e-11 * np.abs(ref).max()


def test_overlap_save_bookkeeping():
    """frames of 4096 inputs starting K-1 before their first output, 4096-(K-1) valid outputs each, zeros (or the
    history) in front of the stream: the concatenated valid parts are lfilter(b, 1, x)"""
    rng = np.random.default_rng(1)
    for K, n in ((257, 10000), (1024, 7000), (2049, 5000)):
        taps = rng.standard_normal(K)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        hist = rng.standard_normal(K - 1) + 1j * rng.standard_normal(K - 1)
        tab = spectrum_table(taps)
        valid = N - (K - 1)
        xe = np.concatenate([hist, x, np.zeros(N)])
        y = np.zeros(n, dtype=np.complex128)
        for f in range((n + valid - 1) // valid):
            out0 = f * valid
            frame = xe[out0:out0 + N]                                  # g0 = out0 - (K-1) in stream coordinates
            yf = frame_model(frame, tab)[K - 1:]
            m = min(valid, n - out0)
            y[out0:out0 + m] = yf[:m]
        ref = np.convolve(np.concatenate([hist, x]), taps)[K - 1:K - 1 + n]
        assert np.abs(y - ref).max() <= 1e-10 * np.abs(ref).max()


def test_two_real_frames_ride_one_complex_transform():
    """float32 streams: the taps are real, so frame 2f as the real part and frame 2f+1 as the imaginary part come out
    of the inverse transform separated again"""
    rng = np.random.default_rng(2)
    taps = rng.standard_normal(700)
    a, b = rng.standard_normal(N), rng.standard_normal(N)
    y = frame_model(a + 1j * b, spectrum_table(taps))
    Hf = np.fft.fft(np.concatenate([taps, np.zeros(N - 700)]))
    assert np.abs(y.real - np.fft.ifft(np.fft.fft(a) * Hf).real).max() <= 1e-10 * np.abs(y).max()
    assert np.abs(y.imag - np.fft.ifft(np.fft.fft(b) * Hf).real).max() <= 1e-10 * np.abs(y).max()
