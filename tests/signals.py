"""Deterministic synthetic PCM for the parity tests (float32, interleaved, in [-1, 1]); BASELINE.md recipes."""
import numpy as np


def sine_noise(seconds, sr=44100, channels=2, f_left=440.0, f_right=554.37, amp=0.5, noise=0.05, seed=1234):
    """C1 / C4 recipe."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    chans = [amp * np.sin(2 * np.pi * f_left * t) + noise * rng.standard_normal(n)]
    if channels == 2:
        chans.append(amp * np.sin(2 * np.pi * f_right * t + 0.3) + noise * rng.standard_normal(n))
    x = np.stack(chans, axis=1).astype(np.float32)
    return np.clip(x, -1.0, 1.0).reshape(-1)


def white(seconds, sr=48000, seed=2, amp=0.5):
    """C2(a): mono white noise U(-amp, amp)."""
    n = int(round(seconds * sr))
    return np.random.default_rng(seed).uniform(-amp, amp, n).astype(np.float32)


def pink(seconds, sr=48000, seed=3, peak=0.5):
    """C2(b): mono pink noise (Kellet 3-pole filter of white noise), scaled to `peak`."""
    n = int(round(seconds * sr))
    w = np.random.default_rng(seed).standard_normal(n)
    b0 = b1 = b2 = 0.0
    out = np.empty(n)
    for i in range(n):
        b0 = 0.99765 * b0 + w[i] * 0.0990460
        b1 = 0.96300 * b1 + w[i] * 0.2965164
        b2 = 0.57000 * b2 + w[i] * 1.0526913
        out[i] = b0 + b1 + b2 + w[i] * 0.1848
    out *= peak / np.max(np.abs(out))
    return out.astype(np.float32)


def castanets(seconds, sr=44100, seed=4, period=0.25, stagger=True):
    """C3: noise floor + decaying bursts (forces short / mixed blocks); R = 0.9 L + tiny noise (forces M/S)."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    left = 1e-3 * rng.standard_normal(n)
    env = 0.8 * np.exp(-np.arange(2000) / 130.0)
    pos, k = 0.05, 0
    while True:
        start = int(pos * sr) + (k * 67 % 576 if stagger else 0)
        if start + 2000 >= n:
            break
        left[start:start + 2000] += env * rng.standard_normal(2000)
        pos += period
        k += 1
    right = 0.9 * left + 5e-4 * rng.standard_normal(n)
    x = np.stack([left, right], axis=1).astype(np.float32)
    return np.clip(x, -1.0, 1.0).reshape(-1)


def sine440(frames, sr=44100, channels=2, amp=0.5):
    """The reference tests' own input: sin(2 pi 440 t) * amp on every channel (TST:84-90, 626-635)."""
    n = frames * 1152
    t = np.arange(n, dtype=np.float64) / sr
    x = (np.sin(2 * np.pi * 440.0 * t) * amp).astype(np.float32)
    return np.repeat(x, channels) if channels == 2 else x
