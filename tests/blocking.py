"""Error analysis used by the 3-sigma statistical tests: Flyvbjerg-Petersen blocking,
restated from the reference's own python/hsmcblocking.py:6-50 (`blocking_std`) without its
matplotlib dependency.  test_stat_cpu.py checks it against the reference's function when
/root/reference is present."""
import numpy as np


def blocking_std(data):
    """sigma[k, 0] = standard error after k blocking transformations, sigma[k, 1] its error."""
    data = np.asarray(data, dtype=float)
    nn = len(data)
    maxp2 = nn.bit_length() - 1
    nn = 2 ** maxp2
    data = data[:nn]
    sigma = np.zeros((maxp2 - 1, 2))
    for ii in range(maxp2 - 1):
        sigma[ii, 0] = np.std(data) / np.sqrt(nn - 1)
        sigma[ii, 1] = sigma[ii, 0] / np.sqrt(2 * (nn - 1))
        data = np.mean(data.reshape(-1, 2), axis=1)
        nn = len(data)
    return sigma


def std_error(data):
    """Plateau estimate: the largest blocked standard error among levels that still have
    >= 16 blocks (conservative, automatic stand-in for reading the plateau off the plot)."""
    s = blocking_std(data)
    n = 2 ** (len(data).bit_length() - 1)
    keep = [k for k in range(len(s)) if n // (2 ** k) >= 16]
    return float(np.max(s[keep, 0])) if keep else float(s[0, 0])


def agree(a, b, nsigma=3.0):
    """|mean(a) - mean(b)| <= nsigma * sqrt(se_a^2 + se_b^2)."""
    ma, mb = float(np.mean(a)), float(np.mean(b))
    se = np.hypot(std_error(a), std_error(b))
    return abs(ma - mb) <= nsigma * se, ma, mb, se


def family_nsigma(m, nsigma=3.0):
    """Per-comparison threshold that keeps the FAMILY-WISE two-sided false-alarm probability of m independent
    comparisons at the level of one nsigma test (Sidak): 1 - (1 - p)^(1/m) with p = erfc(nsigma / sqrt 2).
    m = 1 gives nsigma back; m = 8 gives 3.59 for nsigma = 3."""
    from scipy.special import erfc, erfcinv
    p = erfc(nsigma / np.sqrt(2.0))
    pk = 1.0 - (1.0 - p) ** (1.0 / m)
    return float(np.sqrt(2.0) * erfcinv(pk))
