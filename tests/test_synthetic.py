"""CPU checks of the synthetic input generators the benchmarks and full-size tests share."""
import numpy as np

import oracle
import synthetic


def test_config3_generator_is_the_cpp_mt19937_stream():
    # std::mt19937(1234) / (4321) + uniform_real_distribution<float>; the
    # product-side numpy generator must equal the C++ stream the oracle draws with libstdc++
    lo, hi = (-50.0, -50.0, -3.0), (50.0, 50.0, 10.0)
    for seed in (1234, 4321):
        a = synthetic.mt19937_uniform_box(5000, seed, lo, hi)
        b = oracle.Rng(seed).box_points(5000, lo, hi)
        assert np.array_equal(a, b)
    Q, T = synthetic.knn_config3(100, 50)
    assert Q.shape == (100, 4) and T.shape == (50, 4) and (Q[:, 3] == 1).all()
