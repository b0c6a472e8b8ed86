"""The sampling oracle (oracle/sampler.py) against known answers: the splitmix64 reference output, hand-computed
distributions, tie order, the top-p cut, and the frequency of draws over many seeds."""
import numpy as np

from oracle import sampler as S


def test_uniform_known_answers():
    # splitmix64 from state 0: first output 0xE220A8397B1DCDAF (published test vector of the generator)
    assert S.uniform24(0, 0) == float(0xE220A8397B1DCDAF >> 40) / 2 ** 24
    # the second output of the same stream is what (seed 0, index 1) draws from
    assert S.uniform24(0, 1) == float(0x6E789E6AA1B965F4 >> 40) / 2 ** 24
    u = [S.uniform24(7, i) for i in range(4096)]
    assert 0.0 <= min(u) and max(u) < 1.0 and abs(np.mean(u) - 0.5) < 0.02


def test_candidates_order_and_ties():
    l = np.array([1.0, 3.0, 3.0, -2.0, 3.0, 0.5], dtype=np.float32)
    assert list(S.candidates(l, 4)) == [1, 2, 4, 0]        # equal logits: lowest index first
    assert list(S.candidates(l, 0)) == [1, 2, 4, 0, 5, 3]  # "off" = all (up to MAX_K)
    assert len(S.candidates(np.zeros(5000, np.float32), 0)) == S.MAX_K


def test_distribution_hand_computed():
    l = np.log(np.array([0.5, 0.25, 0.125, 0.125], dtype=np.float64)).astype(np.float32)
    ids, c = S.distribution(l, 1.0, 0, 1.0)
    assert list(ids) == [0, 1, 2, 3]
    np.testing.assert_allclose(c / c[-1], [0.5, 0.75, 0.875, 1.0], rtol=1e-6)
    ids, c = S.distribution(l, 1.0, 0, 0.75)              # the prefix that REACHES 0.75
    assert list(ids) == [0, 1]
    ids, c = S.distribution(l, 1.0, 0, 0.76)
    assert list(ids) == [0, 1, 2]
    ids, c = S.distribution(l, 0.5, 2, 1.0)               # T = 0.5 squares the odds: 4 : 1
    np.testing.assert_allclose(c / c[-1], [0.8, 1.0], rtol=1e-6)


def test_greedy_limit_and_logprob():
    rng = np.random.default_rng(3)
    l = (rng.standard_normal(777) * 4).astype(np.float32)
    i, lp, _ = S.sample(l, 0.0)
    assert i == int(np.argmax(l))
    ref = l.astype(np.float64) - l.max()
    assert abs(lp - (ref[i] - np.log(np.exp(ref).sum()))) < 1e-12
    # a very low temperature draws the argmax whatever the seed
    assert all(S.sample(l, 1e-3, 40, 0.9, seed=s)[0] == i for s in range(50))
    # the reported logprob is the T = 1 log-softmax of the drawn token
    t, lp, _ = S.sample(l, 1.5, 40, 1.0, seed=11)
    assert abs(lp - (ref[t] - np.log(np.exp(ref).sum()))) < 1e-9


def test_draw_frequencies():
    rng = np.random.default_rng(5)
    l = (rng.standard_normal(300) * 2).astype(np.float32)
    ids, c = S.distribution(l, 0.9, 8, 0.95)
    p = np.diff(np.concatenate([[0.0], c])) / c[-1]
    n = 20000
    cnt = np.zeros(len(ids))
    pos = {int(t): j for j, t in enumerate(ids)}
    for s in range(n):
        cnt[pos[S.sample(l, 0.9, 8, 0.95, seed=s, out_index=s % 7)[0]]] += 1
    chi2 = float(((cnt - n * p) ** 2 / (n * p)).sum())
    assert chi2 < 40.0, (chi2, cnt / n, p)                 # dof <= 7: P(chi2 > 40) ~ 1e-6
