"""Running summaries of the prior-scaled coefficients: they define the CG initial guess and the
preconditioner of the un-shrunk coordinates (reference: reg_coef_sampler/reg_coef_posterior_summarizer.py)."""
import numpy as np


class OntheflySummarizer():
    """Online first and second moments with the 1/(1+n) weighting of the reference (:93-124)."""

    def __init__(self, n_param, sd_prior_samplesize=5):
        self.sd_prior_samplesize = sd_prior_samplesize
        self.sd_prior_guess = np.ones(n_param)
        self.n_averaged = 0
        self.stats = {'mean': np.zeros(n_param), 'square': np.ones(n_param)}

    def update_stats(self, theta):
        """mean <- w*theta + (1-w)*mean, square <- w*theta^2 + (1-w)*square (same rounding as the plain
        expressions; written with out= so that no P-length temporaries are allocated per iteration)."""
        w = 1 / (1 + self.n_averaged)
        a = np.multiply(theta, w)
        np.multiply(self.stats['mean'], 1 - w, out=self.stats['mean'])
        np.add(a, self.stats['mean'], out=self.stats['mean'])
        np.multiply(theta, theta, out=a)
        np.multiply(a, w, out=a)
        np.multiply(self.stats['square'], 1 - w, out=self.stats['square'])
        np.add(a, self.stats['square'], out=self.stats['square'])
        self.n_averaged += 1

    def estimate_post_sd(self):
        n = self.n_averaged
        if n <= 1:
            return self.sd_prior_guess
        mean, sq = self.stats['mean'], self.stats['square']
        var_est = n / (n - 1) * (sq - mean ** 2)
        w = (n - 1) / (n - 1 + self.sd_prior_samplesize)
        return np.sqrt(w * var_est + (1 - w) * self.sd_prior_guess ** 2)


class RegressionCoeffficientPosteriorSummarizer():

    def __init__(self, n_coef, n_unshrunk, regularizing_slab_size, pc_summary_method='average'):
        self.n_unshrunk = n_unshrunk
        self.coef_scaled_summarizer = OntheflySummarizer(n_coef)
        self.slab_size = regularizing_slab_size

    def compute_prior_scale(self, gscale, lscale):
        """tau*lambda damped by the slab (reg_coef_sampler.py:194-201). Without a slab the damping factor is exactly 1.
        Not cached: lscale is mutated in place by prior.adjust_scale, and this object is pickled with the chain state."""
        raw = gscale * lscale
        if not np.isinf(self.slab_size):
            raw /= np.sqrt(1 + (raw / self.slab_size) ** 2)
        return raw

    def scale_coef(self, coef, gscale, lscale):
        scaled = coef.copy()
        np.divide(scaled[self.n_unshrunk:], self.compute_prior_scale(gscale, lscale), out=scaled[self.n_unshrunk:])
        return scaled

    def update(self, coef, gscale, lscale):
        self.coef_scaled_summarizer.update_stats(self.scale_coef(coef, gscale, lscale))

    def extrapolate_coef_condmean(self, gscale, lscale):
        guess = self.coef_scaled_summarizer.stats['mean'].copy()
        guess[self.n_unshrunk:] *= self.compute_prior_scale(gscale, lscale)
        return guess

    def estimate_coef_precond_scale_sd(self):
        return self.coef_scaled_summarizer.estimate_post_sd()
