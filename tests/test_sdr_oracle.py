"""oracle/sdr_oracle.py (BSS-Eval SDR for one source) cross-validated: the normal-equation projection against an explicit
least-squares solve on the delay matrix, and closed-form cases."""
import numpy as np

from oracle import sdr_oracle as so
from voicesplit_b200 import synth


def _delay_matrix(ref, flen):
    n = len(ref)
    A = np.zeros((n + flen - 1, flen))
    for k in range(flen):
        A[k:k + n, k] = ref
    return A


def test_projection_equals_explicit_least_squares():
    rng = np.random.Generator(np.random.PCG64(0))
    ref = synth.make_reference_audio(1, 3000, 1)[0].astype(np.float64)
    est = 0.7 * ref + 0.02 * rng.standard_normal(3000)
    flen = 64
    A = _delay_matrix(ref, flen)
    est_pad = np.concatenate((est, np.zeros(flen - 1)))
    c_ls = np.linalg.lstsq(A, est_pad, rcond=None)[0]
    s = A @ c_ls
    want = 10 * np.log10((s ** 2).sum() / ((est_pad - s) ** 2).sum())
    assert abs(so.sdr(ref, est, flen) - want) < 1e-6


def test_correlations_are_plain_sums():
    rng = np.random.Generator(np.random.PCG64(1))
    ref, est = rng.standard_normal(500), rng.standard_normal(500)
    r, d = so.correlations(ref, est, 16)
    for k in range(16):
        assert abs(r[k] - (ref[:500 - k] * ref[k:]).sum()) < 1e-9
        assert abs(d[k] - (ref[:500 - k] * est[k:]).sum()) < 1e-9


def test_closed_form_cases():
    rng = np.random.Generator(np.random.PCG64(2))
    ref = synth.make_reference_audio(1, 16000, 4)[0].astype(np.float64)
    ref[-8:] = 0.0                                          # so that the delayed copy below is not truncated by the frame
    # a filtered (delayed, scaled) copy of the reference lies in the span: SDR is huge
    est = np.zeros_like(ref)
    est[5:] = 0.4 * ref[:-5]
    assert so.sdr(ref, est) > 80
    # white noise at a known level: SDR ~ SNR (noise is almost orthogonal to the 512-dimensional span)
    noise = rng.standard_normal(16000)
    noise *= np.sqrt((ref ** 2).sum() / (noise ** 2).sum()) * 10 ** (-10 / 20)      # 10 dB below the reference
    assert abs(so.sdr(ref, ref + noise) - 10.0) < 0.4
    # scale invariance in the estimate's gain
    assert abs(so.sdr(ref, ref + noise) - so.sdr(ref, 3 * (ref + noise))) < 1e-9
