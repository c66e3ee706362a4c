"""Generate the golden fixtures under tests/golden/ from the REFERENCE's own numpy test oracle.

Run in the build container only (needs the read-only reference checkout):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

It imports, unmodified, ``tests/tools/numpy_kalman_filter.py`` (``NumpyKalmanFilter`` /
``NumpyKalmanFilterWithSites``: classical Kalman filter + RTS smoother, the oracle the reference's
``tests/integration/test_kalman_filter.py:105-139`` and ``test_kalman_filter_with_sites.py`` use) and
``tests/tools/generate_random_objects.py``, follows the set-up of those tests (seed 71892305,
``tests/conftest.py:22``; T=8, D=3, m=2; batch shapes (3,), (), (2,1)) and stores inputs + expected
outputs as ``.npz``.  The TensorFlow side of the reference cannot run here, so these vectors pin what
the reference's tests demand of ``KalmanFilter.log_likelihood`` and
``posterior_state_space_model().marginal_means / marginal_covariances``.

The fixtures are committed; nothing at test time reads /root/reference.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.environ.get("MARKOVFLOW_REFERENCE", "/root/reference"))

from tests.tools.generate_random_objects import (  # noqa: E402  (reference helpers)
    generate_random_lower_triangular_matrix,
    generate_random_pos_def_matrix,
)
from tests.tools.numpy_kalman_filter import (  # noqa: E402
    NumpyKalmanFilter,
    NumpyKalmanFilterWithSites,
)

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 71892305  # tests/conftest.py:22


def kalman_case(batch_shape, tag, num_transitions=7, state_dim=3, output_dim=2, a_scale=1.0):
    """Mirrors tests/integration/test_kalman_filter.py:33-102 (``a_scale`` < 1 keeps the longer,
    larger extra case stable; the three reference-sized cases use the unscaled recipe)."""
    a = a_scale * np.random.normal(size=(state_dim, state_dim))
    chol_q = generate_random_lower_triangular_matrix(state_dim)
    h = np.random.normal(size=(output_dim, state_dim))
    r = generate_random_pos_def_matrix(output_dim)
    chol_r = np.linalg.cholesky(r)
    mu0 = np.random.normal(size=state_dim)
    b = np.random.normal(size=state_dim)
    chol_p0 = generate_random_lower_triangular_matrix(state_dim)
    kf = NumpyKalmanFilter(
        num_timesteps=num_transitions + 1,
        transition_matrix=a,
        transition_mean=b,
        transition_noise=chol_q @ chol_q.T,
        observation_matrix=h,
        observation_noise=r,
        initial_state_prior_mean=mu0,
        initial_state_prior_cov=chol_p0 @ chol_p0.T,
    )
    y = kf.generate_trajectories(batch_shape)
    lls, fm, fp, pm, pp = kf.forward_filter(y)
    sm, sp = kf.backward_smoothing_pass(fm, fp, pm, pp)
    np.savez(
        os.path.join(HERE, f"kalman_{tag}.npz"),
        batch_shape=np.array(batch_shape, dtype=np.int64),
        A=a, chol_Q=chol_q, H=h, chol_R=chol_r, mu0=mu0, b=b, chol_P0=chol_p0, y=y,
        log_liks=lls, filter_means=fm, filter_covs=fp, smooth_means=sm, smooth_covs=sp,
    )


def kalman_sites_case(tag, num_transitions=9, state_dim=3):
    """Mirrors tests/integration/test_kalman_filter_with_sites.py (univariate sites, m=1)."""
    t = num_transitions + 1
    a = 0.5 * np.random.normal(size=(state_dim, state_dim))
    chol_q = generate_random_lower_triangular_matrix(state_dim)
    h = np.random.normal(size=(1, state_dim))
    mu0 = np.random.normal(size=state_dim)
    b = np.random.normal(size=state_dim)
    chol_p0 = generate_random_lower_triangular_matrix(state_dim)
    site_means = np.random.normal(size=(t, 1))
    site_vars = np.random.uniform(0.1, 2.0, size=(t, 1, 1))
    kf = NumpyKalmanFilterWithSites(
        num_timesteps=t,
        transition_matrix=a,
        transition_mean=b,
        transition_noise=chol_q @ chol_q.T,
        observation_matrix=h,
        observation_covariances=site_vars,
        observation_means=site_means,
        initial_state_prior_mean=mu0,
        initial_state_prior_cov=chol_p0 @ chol_p0.T,
    )
    lls, fm, fp, pm, pp = kf.forward_filter(site_means)
    sm, sp = kf.backward_smoothing_pass(fm, fp, pm, pp)
    nat2 = -0.5 / site_vars
    nat1 = site_means / site_vars[..., 0]
    np.savez(
        os.path.join(HERE, f"kalman_sites_{tag}.npz"),
        A=a, chol_Q=chol_q, H=h, mu0=mu0, b=b, chol_P0=chol_p0,
        nat1=nat1, nat2=nat2, site_means=site_means, site_vars=site_vars,
        log_liks=lls, smooth_means=sm, smooth_covs=sp,
    )


def main():
    np.random.seed(SEED)
    kalman_case((3,), "b3")
    kalman_case((), "b0")
    kalman_case((2, 1), "b21")
    kalman_case((4,), "d5m1", num_transitions=20, state_dim=5, output_dim=1, a_scale=0.3)
    kalman_sites_case("t10")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
