"""TEST INFRASTRUCTURE - CPU restatement (float64 numpy) of the SDR the reference's evaluation reports
(/root/reference/utils/generic_utils.py:511: `bss_eval_sources(clean_wav, est_wav, False)[0][0]`, mean over the test
set at :529-533, driven by /root/reference/test.py:71).

bss_eval_sources lives in mir_eval (requirements.txt, unpinned, ABSENT here).  Its published algorithm (BSS Eval v3,
Vincent et al. 2006; mir_eval.separation._bss_decomp_mtifilt / _project / _bss_source_crit) for ONE source and
compute_permutation=False, filter length 512, is restated below:
    s_filt  = orthogonal projection of the (zero-padded) estimate onto the span of the reference delayed by 0..511 samples
              (normal equations G C = D with G the Toeplitz autocorrelation matrix of the reference, D the cross-correlation)
    e_artif = estimate - s_filt          (no interference term with a single source)
    SDR     = 10 log10(|s_filt|^2 / |e_artif|^2)
PARITY UNPINNED against mir_eval itself (it cannot be imported); tests/test_sdr_oracle.py cross-validates the projection
against an explicit least-squares solve on the delay matrix and against closed-form cases.
Only tests/, smoke() and bench tools may import this module."""
import numpy as np

FLEN = 512


def correlations(ref, est, flen=FLEN):
    """r[k] = sum_n ref[n] ref[n+k],  d[k] = sum_n ref[n-k] est[n]   (k = 0..flen-1, signals zero outside their support)."""
    ref = np.asarray(ref, np.float64)
    est = np.asarray(est, np.float64)
    n = len(ref)
    nfft = 1 << int(np.ceil(np.log2(n + flen - 1)))
    sf, sef = np.fft.rfft(ref, nfft), np.fft.rfft(est, nfft)
    r = np.fft.irfft(sf * np.conj(sf), nfft)[:flen]
    c = np.fft.irfft(sf * np.conj(sef), nfft)                 # c[m] = sum_n ref[n+m] est[n]
    d = np.concatenate(([c[0]], c[-1:-flen:-1]))              # d[k] = c[-k]
    return r, d


def projection_filter(ref, est, flen=FLEN):
    r, d = correlations(ref, est, flen)
    idx = np.abs(np.arange(flen)[:, None] - np.arange(flen)[None, :])
    G = r[idx]
    try:
        return np.linalg.solve(G, d)
    except np.linalg.LinAlgError:
        return np.linalg.lstsq(G, d, rcond=None)[0]


def sdr(ref, est, flen=FLEN):
    """bss_eval_sources(ref[None], est[None], compute_permutation=False)[0][0] for one source."""
    ref = np.asarray(ref, np.float64)
    est = np.asarray(est, np.float64)
    if ref.shape != est.shape:
        raise ValueError("reference and estimate must have the same length")
    c = projection_filter(ref, est, flen)
    s_filt = np.convolve(c, ref)                              # length n + flen - 1
    e_artif = -s_filt
    e_artif[:len(est)] += est
    return 10 * np.log10((s_filt ** 2).sum() / (e_artif ** 2).sum())
