"""Synthetic 8-bit IQ for parity tests and the benchmark (SURVEY.md §8d "Synthetic IQ generator").

x[n] = sigma * (g_I + i g_Q) + sum_m A_m(frame) * exp(2 pi i (f_m / fs) n + phi_m), rounded to nearest-even, clipped to
int8, interleaved I,Q. Tones sit at (bin centre + 0.1 bin) so that no two-bin argmax tie can occur; each tone is keyed
on/off per frame. The first `quiet_frames` frames are noise only so the learned threshold is noise.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np


@dataclass
class Tone:
    bin_offset: float  # position in FFT bins relative to DC (negative = below centre), e.g. 1234.1
    amplitude: float = 40.0  # LSB
    on_frames: Sequence[Tuple[int, int]] = field(default_factory=list)  # [start, stop) frame intervals; empty = always on
    phase: float = 0.0
    fm_dev_bins: float = 0.0  # optional sinusoidal FM deviation (bins) — narrow-band FM-like carrier
    fm_rate_cycles_per_frame: float = 3.3


def seed_for(config_no: int, band: int = 0) -> int:
    return 0xB2000000 + 1000 * config_no + band


def tone_active(t: Tone, frame: int) -> bool:
    if not t.on_frames:
        return True
    return any(a <= frame < b for a, b in t.on_frames)


def make_iq_int8(
    n_fft: int,
    n_frames: int,
    tones: List[Tone],
    *,
    sigma: float = 8.0,
    seed: int = 0,
    quiet_frames: int = 0,
    stride: int | None = None,
    chunk_frames: int = 256,
) -> np.ndarray:
    """Returns int8 array of shape [n_frames * stride * 2] (stride defaults to n_fft)."""
    stride = stride or n_fft
    rng = np.random.default_rng(seed)
    out = np.empty(n_frames * stride * 2, dtype=np.int8)
    n = np.arange(stride, dtype=np.float64)
    for f0 in range(0, n_frames, chunk_frames):
        f1 = min(n_frames, f0 + chunk_frames)
        nf = f1 - f0
        x = rng.standard_normal((nf, stride, 2)) * sigma
        z = x[..., 0] + 1j * x[..., 1]
        for t in tones:
            for fi in range(f0, f1):
                if fi < quiet_frames or not tone_active(t, fi):
                    continue
                nn = n + float(fi) * stride  # continuous phase across frames
                ph = 2.0 * np.pi * (t.bin_offset / n_fft) * nn + t.phase
                if t.fm_dev_bins:
                    beta = t.fm_dev_bins / max(t.fm_rate_cycles_per_frame / n_fft, 1e-12) / n_fft
                    ph = ph + beta * np.sin(2.0 * np.pi * t.fm_rate_cycles_per_frame * nn / n_fft)
                z[fi - f0] += t.amplitude * np.exp(1j * ph)
        inter = np.empty((nf, stride, 2), dtype=np.float64)
        inter[..., 0] = z.real
        inter[..., 1] = z.imag
        q = np.clip(np.rint(inter), -128, 127).astype(np.int8)
        out[f0 * stride * 2 : f1 * stride * 2] = q.reshape(-1)
    return out


def standard_scene(n_fft: int, n_frames: int, learn_frames: int) -> List[Tone]:
    """Three keyed narrow-band-FM carriers used by the parity tests: starts, stops, timeouts and overlaps all occur.
    (A bare tone barely moves the 21-bin boxcar of dB values; FM spreads it over ~12 bins like a real NFM channel.)"""
    span = max(n_frames - learn_frames, 1)
    a = learn_frames
    return [
        Tone(bin_offset=0.31 * n_fft / 2 + 0.1, amplitude=60.0, on_frames=[(a + int(0.10 * span), a + int(0.55 * span))], fm_dev_bins=6.0),
        Tone(bin_offset=-0.62 * n_fft / 2 + 0.1, amplitude=60.0, on_frames=[(a + int(0.25 * span), a + int(0.40 * span)), (a + int(0.70 * span), a + int(0.95 * span))], fm_dev_bins=6.0),
        Tone(bin_offset=0.055 * n_fft / 2 + 0.1, amplitude=40.0, on_frames=[(a + int(0.05 * span), a + int(0.30 * span))], phase=1.0, fm_dev_bins=5.0),
    ]


def frame_period_ms(n_fft: int, sample_rate_hz: int, decimator: int = 1) -> float:
    return n_fft * decimator * 1000.0 / sample_rate_hz


def make_iq_int8_torch(n_fft: int, n_frames: int, tones: List[Tone], *, sigma: float = 8.0, seed: int = 0, quiet_frames: int = 0, device="cuda"):
    """GPU generator for benchmark-sized batches (same model, torch RNG). Returns a torch.int8 tensor [n_frames*n_fft*2]."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed & 0x7FFFFFFF)
    out = torch.empty(n_frames * n_fft * 2, dtype=torch.int8, device=device)
    n = torch.arange(n_fft, dtype=torch.float64, device=device)
    chunk = 256
    for f0 in range(0, n_frames, chunk):
        f1 = min(n_frames, f0 + chunk)
        nf = f1 - f0
        re = torch.randn((nf, n_fft), generator=g, device=device, dtype=torch.float32) * sigma
        im = torch.randn((nf, n_fft), generator=g, device=device, dtype=torch.float32) * sigma
        frames = torch.arange(f0, f1, device=device, dtype=torch.float64)[:, None]
        for t in tones:
            mask = torch.tensor([(fi >= quiet_frames) and tone_active(t, fi) for fi in range(f0, f1)], device=device)
            if not bool(mask.any()):
                continue
            nn = n[None, :] + frames * n_fft
            ph = 2.0 * np.pi * (t.bin_offset / n_fft) * nn + t.phase
            if t.fm_dev_bins:
                ph = ph + (t.fm_dev_bins / t.fm_rate_cycles_per_frame) * torch.sin(2.0 * np.pi * t.fm_rate_cycles_per_frame * nn / n_fft)
            ph = torch.remainder(ph, 2.0 * np.pi).to(torch.float32)
            amp = (mask.to(torch.float32) * t.amplitude)[:, None]
            re += amp * torch.cos(ph)
            im += amp * torch.sin(ph)
        q = torch.stack((re, im), dim=-1).round().clamp(-128, 127).to(torch.int8)
        out[f0 * n_fft * 2 : f1 * n_fft * 2] = q.reshape(-1)
    return out
