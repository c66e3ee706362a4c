"""Posterior processes: predictions and samples at arbitrary time points on top of a Gauss-Markov posterior,
with the class names, constructor and method signatures of ``markovflow/posterior.py:37-700``
(SURVEY.md §8f-1, §8f-4).

The structured linear algebra underneath runs on the CUDA operators: pairwise marginals (one fused moment
sweep), conditional statistics / prediction with the gather fused in, sampling with the standard normals drawn
inside the sweep, ``log_pdf`` as one block-per-segment reduction.  ``kernel`` is anything implementing the
``SDEKernel`` protocol of :mod:`markovflow_b200.kernels`; ``likelihood`` anything with ``log_prob(f, y)`` /
``predict_mean_and_var(f_mean, f_var)`` (the reference takes gpflow / markovflow likelihood objects);
``mean_function`` a callable ``t -> batch + [N, output_dim]`` (default: zero).
"""
from __future__ import annotations

import abc
from typing import Callable, Optional, Tuple

import torch

from .conditionals import conditional_predict, conditional_statistics, pairwise_marginals
from .gauss_markov import GaussMarkovDistribution
from .interop import as_torch, boundary, framework_of


def _zero_mean(output_dim: int) -> Callable:
    def mean(t):
        t = as_torch(t)
        return torch.zeros(tuple(t.shape) + (output_dim,), dtype=t.dtype, device=t.device)

    return mean


class PosteriorProcess(abc.ABC):
    """Reference ``posterior.py:37-163``."""

    def sample_state(self, new_time_points, sample_shape, *, input_data=None):
        samples, _ = self.sample_state_trajectories(new_time_points, sample_shape, input_data=input_data)
        return samples

    @abc.abstractmethod
    def sample_state_trajectories(self, new_time_points, sample_shape, *, input_data=None): ...

    @abc.abstractmethod
    def sample_f(self, new_time_points, sample_shape, *, input_data=None): ...

    @abc.abstractmethod
    def predict_state(self, new_time_points): ...

    @abc.abstractmethod
    def predict_f(self, new_time_points, full_output_cov: bool = False): ...


class ConditionalProcess(PosteriorProcess):
    """``q(s(.)) = int p(s(.) | s(Z)) q(s(Z)) ds(Z)`` (reference ``posterior.py:166-410``)."""

    def __init__(self, posterior_dist: GaussMarkovDistribution, kernel, conditioning_time_points,
                 mean_function: Optional[Callable] = None) -> None:
        self._fw = framework_of(conditioning_time_points, posterior_dist)
        self.gauss_markov_model = posterior_dist
        self.kernel = kernel
        self.conditioning_time_points = as_torch(conditioning_time_points)
        self.mean_function = mean_function if mean_function is not None else _zero_mean(kernel.output_dim)

    @property
    def _dev(self) -> torch.device:
        return self.conditioning_time_points.device

    @boundary
    def predict_state(self, new_time_points) -> Tuple[torch.Tensor, torch.Tensor]:
        """State marginals at the (sorted) ``new_time_points`` (reference :207-231)."""
        new = as_torch(new_time_points, self._dev).to(self.conditioning_time_points.dtype)
        dist = self.gauss_markov_model
        pw_mu, pw_cov = pairwise_marginals(
            dist, self.kernel.initial_mean(tuple(dist.batch_shape), new),
            self.kernel.initial_covariance(new[..., :1]))
        return conditional_predict(new, self.conditioning_time_points, self.kernel, pw_mu, pw_cov)

    @boundary
    def predict_f(self, new_time_points, full_output_cov: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Function-value marginals at ``new_time_points`` (reference :233-258); far from the conditioning
        points they revert to the prior."""
        new = as_torch(new_time_points, self._dev).to(self.conditioning_time_points.dtype)
        emission = self.kernel.generate_emission_model(new)
        f_mean, f_cov = emission.project_state_marginals_to_f(*self.predict_state(new), full_output_cov=full_output_cov)
        return f_mean + as_torch(self.mean_function(new), new.device), f_cov

    @boundary
    def sample_state_trajectories(self, new_time_points, sample_shape, *, input_data=None,
                                  seed: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Joint state samples at ``new_time_points`` and at the conditioning points (reference :260-377,
        "Doubly Sparse Variational Gaussian Processes", appendix 2): sample the prior jointly on all points,
        sample the posterior at the conditioning points, and correct the prior sample at the new points by the
        conditional mean of the difference, ``s_o = s_p - P (u_p - u_o)``."""
        if isinstance(sample_shape, int):
            sample_shape = (sample_shape,)
        sample_shape = tuple(int(s) for s in sample_shape)
        z = self.conditioning_time_points
        new = as_torch(new_time_points, z.device).to(z.dtype)
        nz = z.shape[-1]
        joint = torch.cat([z, new], dim=-1)
        order = torch.argsort(joint, dim=-1, stable=True)
        sorted_joint = torch.gather(joint, -1, order)
        s0 = None if seed is None else 2 * int(seed)
        s1 = None if seed is None else 2 * int(seed) + 1
        sorted_samples = self.kernel.state_space_model(sorted_joint).sample(sample_shape, seed=s0)
        unsort = torch.argsort(order, dim=-1)
        d = sorted_samples.shape[-1]
        idx = unsort.expand(sample_shape + tuple(unsort.shape))[..., None].expand(sample_shape + tuple(unsort.shape) + (d,))
        joint_samples = torch.gather(sorted_samples, -2, idx)
        prior_z, prior_new = joint_samples[..., :nz, :], joint_samples[..., nz:, :]
        post_z = self.gauss_markov_model.sample(sample_shape, seed=s1)
        delta = prior_z - post_z
        zero = torch.zeros_like(delta[..., :1, :])
        delta_aug = torch.cat([zero, delta, zero], dim=-2)
        ins = torch.searchsorted(z.contiguous(), new.contiguous())
        ins = ins.expand(sample_shape + tuple(ins.shape))[..., None].expand(sample_shape + tuple(ins.shape) + (d,))
        v = torch.cat([torch.gather(delta_aug, -2, ins), torch.gather(delta_aug, -2, ins + 1)], dim=-1)
        proj, _ = conditional_statistics(new, z, self.kernel)
        return prior_new - (proj @ v[..., None])[..., 0], post_z

    @boundary
    def sample_f(self, new_time_points, sample_shape, *, input_data=None, seed: Optional[int] = None):
        """Function-value samples at ``new_time_points`` (reference :379-410)."""
        new = as_torch(new_time_points, self._dev).to(self.conditioning_time_points.dtype)
        states, _ = self.sample_state_trajectories(new, sample_shape, input_data=input_data, seed=seed)
        f = self.kernel.generate_emission_model(new).project_state_to_f(states)
        return f + as_torch(self.mean_function(new), new.device)


class AnalyticPosteriorProcess(ConditionalProcess):
    """Reference ``posterior.py:413-467``: adds ``predict_y`` through the likelihood."""

    def __init__(self, posterior_dist, kernel, conditioning_time_points, likelihood,
                 mean_function: Optional[Callable] = None) -> None:
        super().__init__(posterior_dist, kernel, conditioning_time_points, mean_function)
        self.likelihood = likelihood

    @boundary
    def predict_y(self, new_time_points, full_output_cov: bool = False):
        return self.likelihood.predict_mean_and_var(*self.predict_f(new_time_points, full_output_cov=full_output_cov))


class ImportanceWeightedPosteriorProcess(PosteriorProcess):
    """Posterior inferred by importance-weighted variational inference (reference ``posterior.py:470-700``):
    samples are drawn from the proposal process and weighted by ``w = p(Y | s) p(u) / q(u)``."""

    def __init__(self, num_importance_samples: int, proposal_dist: GaussMarkovDistribution, kernel,
                 conditioning_time_points, likelihood, mean_function: Optional[Callable] = None) -> None:
        self.proposal_process = ConditionalProcess(proposal_dist, kernel, conditioning_time_points, mean_function)
        self._fw = self.proposal_process._fw
        self.num_importance_samples = int(num_importance_samples)
        self.likelihood = likelihood

    @property
    def _dev(self) -> torch.device:
        return self.proposal_process._dev

    def _log_qu_density(self, samples_u, stop_gradient: bool = False):
        q = self.proposal_process.gauss_markov_model
        if stop_gradient:
            q = q.create_non_trainable_copy()
        return q.log_pdf(samples_u)

    @boundary
    def log_importance_weights(self, samples_s, samples_u, input_data, stop_gradient: bool = False):
        """``log w = log p(Y | s) + log p(u) - log q(u)``, shape ``sample_shape`` (reference :533-590)."""
        pp = self.proposal_process
        times, obs = as_torch(input_data[0], pp._dev), as_torch(input_data[1], pp._dev)
        dist_p = pp.kernel.state_space_model(pp.conditioning_time_points)
        log_pu = dist_p.log_pdf(samples_u)
        log_qu = self._log_qu_density(samples_u, stop_gradient=stop_gradient)
        f = pp.kernel.generate_emission_model(times).project_state_to_f(as_torch(samples_s, pp._dev))
        f = f + as_torch(pp.mean_function(times), f.device)
        log_lik = torch.sum(self.likelihood.log_prob(f, obs), dim=-1)
        nb = len(tuple(dist_p.batch_shape))
        diff = log_pu - log_qu
        if nb:
            diff = diff.sum(dim=tuple(range(-nb, 0)))
            log_lik = log_lik.sum(dim=tuple(range(-nb, 0))) if log_lik.dim() > diff.dim() else log_lik
        return log_lik + diff

    def _iwvi_samples_and_weights(self, new_time_points, input_data, sample_shape, seed=None):
        pp = self.proposal_process
        times = as_torch(input_data[0], pp._dev)
        new = as_torch(new_time_points, pp._dev).to(times.dtype)
        all_t = torch.cat([times, new], dim=-1)
        s, u = pp.sample_state_trajectories(all_t, sample_shape, seed=seed)
        n_new = new.shape[-1]
        s_new, s_data = s[..., -n_new:, :], s[..., :-n_new, :]
        return s_new, self.log_importance_weights(s_data, u, input_data), u

    @boundary
    def sample_state_trajectories(self, new_time_points, sample_shape, *, input_data=None, seed: Optional[int] = None):
        """One state trajectory per requested sample, resampled from ``num_importance_samples`` proposals by
        their importance weights; the conditioning samples are returned as drawn, ``sample_shape +
        [num_importance_samples] + ...`` (reference :622-671)."""
        if input_data is None:
            raise ValueError("You need to provide `input_data` for doing inference with IW")
        if isinstance(sample_shape, int):
            sample_shape = (sample_shape,)
        sample_shape = tuple(int(s) for s in sample_shape)
        k = self.num_importance_samples
        s_new, log_w, u = self._iwvi_samples_and_weights(new_time_points, input_data, sample_shape + (k,), seed=seed)
        rest = tuple(s_new.shape[len(sample_shape) + 1:])
        flat = s_new.reshape((-1, k) + rest)
        gen = None if seed is None else torch.Generator(device=log_w.device).manual_seed(int(seed))
        pick = torch.multinomial(torch.softmax(log_w.reshape(-1, k), dim=-1), 1, generator=gen)[:, 0]
        out = flat[torch.arange(flat.shape[0], device=flat.device), pick]
        return out.reshape(sample_shape + rest), u

    @boundary
    def sample_f(self, new_time_points, sample_shape, *, input_data=None, seed: Optional[int] = None):
        pp = self.proposal_process
        new = as_torch(new_time_points, pp._dev)
        states, _ = self.sample_state_trajectories(new, sample_shape, input_data=input_data, seed=seed)
        f = pp.kernel.generate_emission_model(new).project_state_to_f(states)
        return f + as_torch(pp.mean_function(new), f.device)

    @boundary
    def expected_value(self, new_time_points, input_data, func: Callable = lambda x: x, seed: Optional[int] = None):
        """Self-normalised importance-weighted expectation ``E[func(f(new_time_points))]`` (reference :592-618)."""
        pp = self.proposal_process
        new = as_torch(new_time_points, pp._dev)
        s_new, log_w, _ = self._iwvi_samples_and_weights(new, input_data, (self.num_importance_samples,), seed=seed)
        f = pp.kernel.generate_emission_model(new).project_state_to_f(s_new)
        f = f + as_torch(pp.mean_function(new), f.device)
        w = torch.softmax(log_w, dim=0)
        return torch.sum(w.reshape((-1,) + (1,) * (f.dim() - 1)) * func(f), dim=0)

    @boundary
    def predict_state(self, new_time_points):
        raise NotImplementedError("the importance-weighted posterior has no closed-form state marginals "
                                  "(reference posterior.py:667-677)")

    @boundary
    def predict_f(self, new_time_points, full_output_cov: bool = False):
        raise NotImplementedError("use expected_value / sample_f (reference posterior.py:679-700)")
